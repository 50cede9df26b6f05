#!/bin/bash
# second profile pass: the CTA-pair GEMM at the h=768 shapes, all four MAKD kernels, attention backward
set -u
TAG=${1:-r01_v5}
OUT=gpurun_out
mkdir -p $OUT
TMP=/tmp/ncu_$TAG
mkdir -p $TMP
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none -k regex:$rx --launch-skip $skip -c $cnt -o $TMP/$name "$@" > $OUT/${TAG}_$name.log 2>&1
  ncu -i $TMP/$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  rm -f $TMP/$name.ncu-rep
}
cap gemm_pair gemm_tc 0 8 python scripts/pair_check.py pair_only
cap makd makd 0 8 python scripts/makd_micro.py bf16 once
cap attn_bwd attn_mma_bwd 0 6 python scripts/graph_micro.py attn
du -sh $OUT
