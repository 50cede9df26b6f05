#!/bin/bash
# Round 2, GPU call C (2 GPUs): overlapped gradient exchange -- correctness vs the plain path, then throughput.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR scripts/overlap_check.py magic_s_pretrain_b64 > $OUT/c_check_s.log 2>&1; echo "check_s rc=$?"; grep -v Warn $OUT/c_check_s.log | tail -8
timeout 600 $TR scripts/overlap_check.py magic_l_icod_b32 > $OUT/c_check_icod.log 2>&1; echo "check_icod rc=$?"; grep -v Warn $OUT/c_check_icod.log | tail -8
for ov in 1 0; do
  for wl in magic_s_pretrain_b64 magic_l_icod_b32; do
    timeout 400 $TR bench.py --gpus 2 --workload $wl --overlap $ov --sub-workloads none --no-profile > $OUT/c_n2_${wl}_ov$ov.json 2> $OUT/c_n2_${wl}_ov$ov.err; echo "$wl ov=$ov rc=$?"
  done
done
timeout 400 $TR bench.py --gpus 2 --sub-workloads none > $OUT/c_n2_default.json 2> $OUT/c_n2_default.err; echo "default n2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c_n2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
