#!/bin/bash
OUT=gpurun_out
for v in v1 v2 v3 v4 v5 v6 v7; do
  timeout 300 python scripts/capture_repro.py $v > $OUT/i_repro_$v.log 2>&1
  grep -v "Warning\|warn" $OUT/i_repro_$v.log | tail -12 | cut -c1-220
done
timeout 900 python -m pytest tests/test_train_loop_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/i_pytest.log 2>&1; tail -15 $OUT/i_pytest.log | cut -c1-250
timeout 900 python bench.py --no-cpu --no-gpu-baseline --sub-workloads none > $OUT/i_bench_default.json 2> $OUT/i_bench_default.err; echo "bench rc=$?"; tail -3 $OUT/i_bench_default.err
timeout 900 python bench.py --no-cpu --no-gpu-baseline --sub-workloads none --workload magic_l_icod_b32 > $OUT/i_bench_icod.json 2> $OUT/i_bench_icod.err; echo "bench rc=$?"
timeout 900 python bench.py --no-cpu --no-gpu-baseline --sub-workloads none --workload rxr_stress_distill_b128 > $OUT/i_bench_rxr.json 2> $OUT/i_bench_rxr.err; echo "bench rc=$?"
