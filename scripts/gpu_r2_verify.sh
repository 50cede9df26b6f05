#!/bin/bash
# Final verification of the round: the whole GPU suite, then smoke().  Logs land in gpurun_out/.
OUT=gpurun_out
timeout 150 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider > $OUT/verify_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/verify_pytest.log | cut -c1-300
grep -E "^(FAILED|ERROR)" $OUT/verify_pytest.log | head -20
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/verify_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/verify_smoke.log
