#!/bin/bash
# the driver's own lines: default bench (all legs) + the reference arm
OUT=gpurun_out
timeout 1500 python bench.py > $OUT/u_bench.json 2> $OUT/u_bench.err; echo "bench rc=$?"; tail -3 $OUT/u_bench.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/u_bench_ref.json 2> $OUT/u_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/u_bench.json').read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e'])
for r in d['rooflines']: print('  roof', r['kernel'], round(r['frac'],3), round(r['achieved'],1), r['unit'])
print('  nav', d.get('nav_inference')); print('  feat', d.get('featurizer'))
print('  gpu_baseline', d.get('gpu_baseline')); print('  cpu', d.get('cpu_baseline'))
for k,v in (d.get('workloads') or {}).items(): print('  ', k, v.get('value'), v.get('ms_per_step'), v.get('error'))
print(open('gpurun_out/u_bench_ref.json').read()[:600])
PY
