#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_adamw_golden_gpu.py tests/test_model_gpu.py tests/test_stepper_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "adamw or inactive or task or stepper or icod or graph" > $OUT/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/y_pytest.log | cut -c1-250
run() { echo "== $*"; timeout 600 python bench.py --timed-only --steps 40 "$@" 2>&1 | grep "timed-only\|Error" | head -3; }
run --workload magic_s_distill_t768_b64
run --workload magic_l_icod_b32
