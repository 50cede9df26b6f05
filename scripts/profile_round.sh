#!/bin/bash
# Round profile pass (run on the GPU box through gpurun): ncu launch list of the eager step + `--set full` captures of
# the GEMM / MAKD / attention kernels, reduced ON THE BOX to small CSVs (the .ncu-rep files stay behind: gpurun_out
# is capped at 64 MiB).  Usage: bash scripts/profile_round.sh <tag>
set -u
TAG=${1:-r01_v5}
OUT=gpurun_out
mkdir -p $OUT
TMP=/tmp/ncu_$TAG
mkdir -p $TMP
B="python bench.py --steps 1 --warmup 3 --no-cpu --graphs 0 --timed-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none -k regex:$rx --launch-skip $skip -c $cnt -o $TMP/$name "$@" > $OUT/${TAG}_$name.log 2>&1
  ncu -i $TMP/$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  rm -f $TMP/$name.ncu-rep
}
cap gemm_s gemm_tc 500 24 $B
cap gemm_pair gemm_tc 0 8 python scripts/pair_check.py pair_only
cap makd makd 0 8 python scripts/makd_micro.py bf16 once
cap attn attn_mma_fwd 0 6 python scripts/graph_micro.py attn
cap attn_bwd attn_mma_bwd 0 6 python scripts/graph_micro.py attn
ls -la $OUT | tail -20
du -sh $OUT
