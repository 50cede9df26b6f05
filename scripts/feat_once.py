"""One magic_featurize_graph call (B = 64, 256-viewpoint synthetic world) -- the target of the ncu capture in
scripts/gpu_r2_k2.sh."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
print(bench.featurizer_rate(calls=3))
