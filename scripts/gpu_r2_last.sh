#!/bin/bash
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/last_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/last_pytest.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/last_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/last_smoke.log
timeout 1500 python bench.py > $OUT/last_bench.json 2> $OUT/last_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/last_bench.json').read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], d['clocks'])
for k,v in (d.get('workloads') or {}).items(): print('  ', k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
