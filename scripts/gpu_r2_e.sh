#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "og or mrc or ignore" > $OUT/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/e_pytest.log
timeout 600 python scripts/timeline.py --workload magic_s_distill_t768_b64 --steps 6 --lookahead 1 --out $OUT/e_timeline_pl1.csv > $OUT/e_timeline_pl1.log 2>&1; tail -3 $OUT/e_timeline_pl1.log
timeout 600 python scripts/timeline.py --workload magic_s_distill_t768_b64 --steps 6 --lookahead 0 --out $OUT/e_timeline_pl0.csv > $OUT/e_timeline_pl0.log 2>&1; tail -3 $OUT/e_timeline_pl0.log
ls -la $OUT/e_*
