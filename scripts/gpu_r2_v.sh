#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_nav_gpu.py tests/test_stepper_gpu.py tests/test_bf16_parity_gpu.py tests/test_train_loop_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/v_pytest.log | cut -c1-220
run() { echo "== $*"; timeout 600 python bench.py --timed-only --steps 30 "$@" 2>&1 | grep "timed-only"; }
run --workload magic_s_distill_t768_b64
run --workload rxr_stress_distill_b128
run --workload magic_l_icod_b32
# memcheck of the kernels added this round (graph featuriser, attention key_skip / forward v3, nav path)
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_featurize_graph_gpu.py tests/test_nav_gpu.py "tests/test_kernels_gpu.py::test_attention_key_skip" "tests/test_kernels_gpu.py::test_attention" -m gpu -q -x --timeout 1100 -p no:cacheprovider -k "not 768 and not 16-128" > $OUT/v_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -6 $OUT/v_sanitizer.log | cut -c1-200
# GEMM capture after the one-MUFU GELU epilogue (teacher GEMMs of the 4th step)
TMP=/tmp/ncu_v; mkdir -p $TMP
B="python bench.py --workload magic_s_distill_t768_b64 --steps 1 --warmup 3 --no-cpu --graphs 0 --timed-only"
ncu --set full --clock-control none -k regex:gemm_tc --launch-skip 1060 -c 16 -o $TMP/gemm $B > $OUT/r02_v2_gemm.log 2>&1
ncu -i $TMP/gemm.ncu-rep --page raw --csv > $OUT/r02_v2_gemm_raw.csv 2>/dev/null
ls -la $OUT/r02_v2_gemm_raw.csv
