#!/bin/bash
set -u
OUT=gpurun_out
run() { echo "== $*"; timeout 600 python bench.py --timed-only --steps 40 "$@" 2>&1 | grep "timed-only"; }
run --pdl 0
run --pdl 1
run --pdl 2
run --pdl 3
run --pdl 0 --workload magic_s_pretrain_b64
run --pdl 1 --workload magic_s_pretrain_b64
run --pdl 0 --workload magic_l_pretrain_b32
run --pdl 1 --workload magic_l_pretrain_b32
run --pdl 0 --workload magic_l_icod_b32
run --pdl 3 --workload magic_l_icod_b32
run --pdl 0 --workload rxr_stress_distill_b128
timeout 900 python bench.py --no-cpu --no-gpu-baseline --sub-workloads none > $OUT/h_bench_default.json 2> $OUT/h_bench_default.err; echo "bench rc=$?"; tail -3 $OUT/h_bench_default.err
timeout 900 python bench.py --no-cpu --no-gpu-baseline --sub-workloads none --workload magic_l_icod_b32 > $OUT/h_bench_icod.json 2> $OUT/h_bench_icod.err; echo "bench rc=$?"
timeout 900 python bench.py --no-cpu --no-gpu-baseline --sub-workloads none --workload rxr_stress_distill_b128 > $OUT/h_bench_rxr.json 2> $OUT/h_bench_rxr.err; echo "bench rc=$?"
