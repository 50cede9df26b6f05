#!/bin/bash
# final verification + profile pass of round 2
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/final_pytest.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $OUT/final_smoke.log | cut -c1-200
timeout 1500 python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/final_bench_ref.json 2> $OUT/final_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])
for r in d['rooflines']: print('  roof', r['kernel'][:60], round(r['frac'],3), round(r['achieved'],1), r['unit'], 'traffic', r.get('traffic'))
print('  nav', json.dumps(d.get('nav_inference'))[:600]); print('  feat', d.get('featurizer'))
print('  gpu_baseline', d.get('gpu_baseline',{}).get('value')); print('  cpu', d.get('cpu_baseline',{}).get('value'))
for k,v in (d.get('workloads') or {}).items(): print('  ', k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
# profile pass: launch list of the eager step + --set full of the whole GEMM family of one step, attention, MAKD
TMP=/tmp/ncu_f; mkdir -p $TMP
B="python bench.py --workload magic_s_distill_t768_b64 --steps 1 --warmup 3 --no-cpu --graphs 0 --timed-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r02_v3_launches.csv $B > $OUT/r02_v3_launches.log 2>&1
cap() { local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none -k regex:$rx --launch-skip $skip -c $cnt -o $TMP/$name "$@" > $OUT/r02_v3_$name.log 2>&1
  ncu -i $TMP/$name.ncu-rep --page raw --csv > $OUT/r02_v3_${name}_raw.csv 2>/dev/null; rm -f $TMP/$name.ncu-rep; }
cap gemm_step gemm_tc 1050 350 $B
cap attn "attn_(mma_fwd|fwd2|fwd3|mma_bwd)" 200 67 $B
cap makd makd 12 4 $B
ls -la $OUT | grep r02_v3
