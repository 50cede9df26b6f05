#!/bin/bash
# ncu --set full of the kernels added this round: attention forward v3, segmented AdamW, graph featuriser
OUT=gpurun_out; TMP=/tmp/ncu_k2; mkdir -p $TMP
B="python bench.py --workload magic_s_pretrain_b64 --steps 1 --warmup 3 --no-cpu --graphs 0 --timed-only"
timeout 600 ncu --set full --clock-control none -k regex:"attn_fwd3|adamw_seg" --launch-skip 60 -c 14 -o $TMP/a $B > $OUT/r02_v4_new.log 2>&1
ncu -i $TMP/a.ncu-rep --page raw --csv > $OUT/r02_v4_new_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"featurize" --launch-skip 2 -c 4 -o $TMP/f python scripts/feat_once.py > $OUT/r02_v4_feat.log 2>&1
ncu -i $TMP/f.ncu-rep --page raw --csv > $OUT/r02_v4_feat_raw.csv 2>/dev/null
ls -la $OUT | grep r02_v4
