#!/bin/bash
# Round-2 profile pass on the GPU box (through gpurun): the headline workload (teacher h=768 -> MAGIC-S distillation
# step, configs[2]) under ncu.  (1) launch list of the eager step (gpu__time_duration.sum per launch);
# (2) `--set full` captures of the kernels bench.py's rooflines name -- the teacher's tcgen05 GEMMs, attention forward /
# backward, the MAKD loss kernels -- taken FROM THE STEP, reduced on the box to raw CSVs (the .ncu-rep files stay
# behind: gpurun_out is capped at 64 MiB).  Usage: bash scripts/profile_r02.sh <tag> [workload]
set -u
TAG=${1:-r02_v1}
WL=${2:-magic_s_distill_t768_b64}
OUT=gpurun_out
mkdir -p $OUT
TMP=/tmp/ncu_$TAG
mkdir -p $TMP
B="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu --graphs 0 --timed-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none -k regex:$rx --launch-skip $skip -c $cnt -o $TMP/$name "$@" > $OUT/${TAG}_$name.log 2>&1
  ncu -i $TMP/$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  rm -f $TMP/$name.ncu-rep
}
# the 4th step of the run (skip = launches of that kernel family in 3 steps, measured from the launch list of r01/r02)
cap gemm gemm_tc 1060 40 $B
cap attn_fwd "attn_(mma_fwd|fwd2)" 141 24 $B
cap attn_bwd attn_mma_bwd 120 12 $B
cap makd makd 12 8 $B
ls -la $OUT | tail -12
du -sh $OUT
