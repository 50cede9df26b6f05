#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_nav_gpu.py tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py tests/test_model_gpu.py tests/test_bf16_parity_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/r_pytest.log | cut -c1-220
run() { echo "== $*"; timeout 600 env $1 python bench.py --timed-only --steps 30 ${@:2} 2>&1 | grep "timed-only"; }
for wl in magic_s_distill_t768_b64 rxr_stress_distill_b128 magic_l_icod_b32 magic_s_pretrain_b64 magic_l_pretrain_b32; do
  run X=0 --workload $wl
done
