#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "inactive" 2>&1 | tail -2
timeout 1500 python bench.py > $OUT/z_bench.json 2> $OUT/z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z_bench.json').read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])
for r in d['rooflines']: print('  roof', r['kernel'][:60], round(r['frac'],3), round(r['achieved'],1), r['unit'], 'traffic', r.get('traffic'))
print('  gpu_baseline', d.get('gpu_baseline',{}).get('value')); print('  cpu', d.get('cpu_baseline',{}).get('value'))
for k,v in (d.get('workloads') or {}).items(): print('  ', k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
