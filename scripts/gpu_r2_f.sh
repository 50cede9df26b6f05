#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_gemm_tc_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/f_pytest.log
run() { echo "== $*"; timeout 600 env $1 python bench.py --timed-only --steps 40 ${@:2} 2>&1 | grep "timed-only"; }
run X=1 --pipeline 1
run X=1 --pipeline 0
run MAGIC_TC_MAX_SMS=132 --pipeline 1
run MAGIC_TC_MAX_SMS=120 --pipeline 1
run MAGIC_TC_MAX_SMS=120 --pipeline 0
run X=1 --pipeline 1 --teacher-sms 120
run X=1 --pipeline 1 --teacher-sms 104
run X=1 --pipeline 1 --teacher-sms 88
run X=1 --pipeline 1 --workload magic_s_pretrain_b64
run X=1 --pipeline 1 --workload magic_l_pretrain_b32
run MAGIC_TC_MAX_SMS=132 --pipeline 1 --workload magic_l_pretrain_b32
run X=1 --pipeline 1 --workload rxr_stress_distill_b128
run X=1 --pipeline 1 --workload rxr_stress_distill_b128 --teacher-sms 120
