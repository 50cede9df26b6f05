#!/bin/bash
# Local wrapper: rebuild the in-tree .so (so the snapshot carries a library that matches the sources), then gpurun.
# usage: scripts/run_gpu.sh <timeout_s> <log> [--gpus N] -- <command>
set -e
cd "$(dirname "$0")/.."
T=$1; LOG=$2; shift 2
python -c "import __graft_entry__ as g; g.build()" >/dev/null
python -m pytest tests/test_abi.py -q -x >/dev/null
/usr/local/graft/bin/gpurun --timeout $T "$@" > $LOG 2>&1 || true
tail -30 $LOG
