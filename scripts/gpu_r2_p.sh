#!/bin/bash
# full GPU test suite at HEAD + the round-2 profile pass
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/p_pytest.log | cut -c1-200
timeout 900 bash scripts/profile_r02.sh r02_v1 > $OUT/p_profile.log 2>&1; echo "profile rc=$?"; tail -14 $OUT/p_profile.log
timeout 300 python scripts/graph_micro.py attn_l > $OUT/p_attn_l.log 2>&1; cat $OUT/p_attn_l.log
