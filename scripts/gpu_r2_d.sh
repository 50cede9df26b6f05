#!/bin/bash
# Round 2, GPU call D: full GPU suite (teacher pipelining, deterministic grad norm), A/B of the pipelined teacher.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/d_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/d_pytest.log
tail -6 $OUT/d_pytest.log
for pl in 1 0; do
  timeout 600 python bench.py --pipeline $pl --timed-only --steps 40 2> $OUT/d_timed_pl$pl.err; tail -1 $OUT/d_timed_pl$pl.err
done
timeout 600 python bench.py --workload rxr_stress_distill_b128 --pipeline 1 --timed-only --steps 20 2> $OUT/d_timed_rxr_pl1.err; tail -1 $OUT/d_timed_rxr_pl1.err
timeout 900 python bench.py --no-cpu > $OUT/d_bench_default.json 2> $OUT/d_bench_default.err; echo "bench rc=$?"; tail -3 $OUT/d_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d_bench_default.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['profile'])
for k,v in (d.get('workloads') or {}).items(): print(k, {a:b for a,b in v.items() if a in('value','ms_per_step','error')})
PY
