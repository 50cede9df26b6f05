#!/bin/bash
# Round 2, GPU call A: full GPU test suite, smoke, the default bench line, per-shape GEMM table, bf16 gradient probe.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/a_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/a_pytest.log
tail -40 $OUT/a_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/a_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/a_smoke.log; tail -8 $OUT/a_smoke.log
timeout 900 python bench.py > $OUT/a_bench_default.json 2> $OUT/a_bench_default.err; echo "bench rc=$?"; tail -5 $OUT/a_bench_default.err
for wl in magic_l_pretrain_b32 magic_l_icod_b32 rxr_stress_distill_b128; do
  timeout 600 python bench.py --workload $wl --sub-workloads none --no-cpu --no-gpu-baseline > $OUT/a_bench_$wl.json 2> $OUT/a_bench_$wl.err; echo "$wl rc=$?"
done
timeout 300 python scripts/pair_check.py > $OUT/a_pair.log 2>&1
timeout 300 python scripts/pair_check.py small > $OUT/a_pair_small.log 2>&1
timeout 300 python scripts/bf16_grad_probe.py sap > $OUT/a_gradprobe_sap.log 2>&1
timeout 300 python scripts/bf16_grad_probe.py mlm > $OUT/a_gradprobe_mlm.log 2>&1
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/a_bench_ref.json 2> $OUT/a_bench_ref.err; echo "ref rc=$?"
ls -la $OUT | tail -30
