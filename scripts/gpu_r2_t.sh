#!/bin/bash
# 2-GPU check of the driver's scaling launch line (default workload + sub-workloads) and the graph-featuriser tests
OUT=gpurun_out
timeout 300 python -m pytest tests/test_featurize_graph_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 4 > $OUT/t_bench_n2.json 2> $OUT/t_bench_n2.err; echo "bench n2 rc=$?"; tail -3 $OUT/t_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench_n2.json').read().strip().splitlines()[-1])
print('N=2', d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config'].get('gradient_exchange'))
for k,v in (d.get('workloads') or {}).items(): print('  ', k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
