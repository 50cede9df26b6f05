#!/bin/bash
# nav path + attention v3 + key_skip + GELU: tests, then attention micro (v3 vs v2/v1) and the step benches
OUT=gpurun_out
timeout 900 python -m pytest tests/test_nav_gpu.py tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/q_pytest.log | cut -c1-220
echo "== attn micro v3"; timeout 300 python scripts/graph_micro.py attn_l 2>&1 | grep "attn fwd" > $OUT/q_attn_v3.log; cat $OUT/q_attn_v3.log
echo "== gemm_l"; timeout 300 python scripts/graph_micro.py gemm_l 2>&1 | grep "gelu\|fwd " | head -12
run() { echo "== $*"; timeout 600 env $1 python bench.py --timed-only --steps 30 ${@:2} 2>&1 | grep "timed-only"; }
for wl in magic_s_distill_t768_b64 rxr_stress_distill_b128 magic_l_icod_b32 magic_s_pretrain_b64; do
  run MAGIC_ATTN_FWD=0 --workload $wl
  run MAGIC_ATTN_FWD=2 --workload $wl
done
