#!/bin/bash
run() { echo "== $*"; timeout 600 env $1 python bench.py --timed-only --steps 40 ${@:2} 2>&1 | grep "timed-only\|Error\|error" | head -3; }
for wl in magic_s_distill_t768_b64 rxr_stress_distill_b128; do
  run MAGIC_TEACHER_PRIO=0 --workload $wl
  run MAGIC_TEACHER_PRIO=1 --workload $wl
done
