#!/bin/bash
OUT=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 4 > $OUT/n8_bench.json 2> $OUT/n8_bench.err; echo "bench n8 rc=$?"; tail -2 $OUT/n8_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n8_bench.json').read().strip().splitlines()[-1])
print('N=8', d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config'].get('gradient_exchange'))
for k,v in (d.get('workloads') or {}).items(): print('  ', k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
