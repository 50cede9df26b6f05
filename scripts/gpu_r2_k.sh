#!/bin/bash
OUT=gpurun_out
B="python bench.py --no-cpu --no-gpu-baseline --sub-workloads none"
for wl in magic_s_distill_t768_b64 magic_l_icod_b32 rxr_stress_distill_b128 magic_s_pretrain_b64; do
  timeout 600 $B --workload $wl > $OUT/k_$wl.json 2> $OUT/k_$wl.err; echo "$wl rc=$?"
done
timeout 600 python -m pytest tests/test_stepper_gpu.py tests/test_train_loop_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
