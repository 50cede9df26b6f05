#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_nav_gpu.py tests/test_bf16_parity_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "attention or attn or rollout or config2 or config4" 2>&1 | tail -3
timeout 300 python scripts/graph_micro.py attn_l 2>&1 | grep "attn fwd" | grep "pbar1" | tee $OUT/a2_attn.log
run() { echo "== $*"; timeout 600 python bench.py --timed-only --steps 40 "$@" 2>&1 | grep "timed-only\|Error" | head -2; }
run --workload magic_s_distill_t768_b64
run --workload rxr_stress_distill_b128
