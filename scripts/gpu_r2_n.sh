#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_featurizer_gpu.py tests/test_kernels_gpu.py tests/test_stepper_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/n_pytest.log | cut -c1-200
for fs in 2048 0; do
  timeout 900 python bench.py --no-cpu --no-gpu-baseline --sub-workloads none --feature-store $fs > $OUT/n_bench_fs$fs.json 2> $OUT/n_bench_fs$fs.err; echo "bench fs=$fs rc=$?"
done
python - <<'PY'
import json
for fs in (2048, 0):
    d=json.loads(open(f'gpurun_out/n_bench_fs{fs}.json').read().strip().splitlines()[-1])
    print(fs, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['e2e']['h2d_bytes_per_step'], d['profile'].get('sum_kernel_ms'))
PY
