#!/bin/bash
# Round 2, GPU call B: re-check the bf16 parity file, default bench line with the in-graph stopwatch, per-workload
# profiles, bf16 gradient bisect.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_bf16_parity_gpu.py tests/test_stepper_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/b_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/b_pytest.log
tail -5 $OUT/b_pytest.log
timeout 900 python bench.py --no-cpu > $OUT/b_bench_default.json 2> $OUT/b_bench_default.err; echo "bench rc=$?"; tail -5 $OUT/b_bench_default.err
for wl in magic_l_pretrain_b32 rxr_stress_distill_b128 magic_s_pretrain_b64; do
  timeout 600 python bench.py --workload $wl --sub-workloads none --no-cpu --no-gpu-baseline > $OUT/b_bench_$wl.json 2> $OUT/b_bench_$wl.err; echo "$wl rc=$?"
done
timeout 300 python scripts/bf16_bisect.py sap > $OUT/b_bisect_sap.log 2>&1
timeout 300 python scripts/bf16_bisect.py mlm > $OUT/b_bisect_mlm.log 2>&1
ls -la $OUT | tail -12
