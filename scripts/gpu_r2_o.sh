#!/bin/bash
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR scripts/overlap_check.py magic_l_icod_b32 fp32 2>&1 | grep -v "^$\|Warning\|warn" | tail -8
timeout 600 $TR scripts/overlap_check.py magic_l_icod_b32 bf16 2>&1 | grep -v "^$\|Warning\|warn" | tail -8
for x in fp32 bf16; do
  echo "== icod N=2 exchange=$x"; timeout 600 $TR bench.py --gpus 2 --timed-only --steps 30 --workload magic_l_icod_b32 --exchange $x 2>&1 | grep "timed-only\|Error" | head -2
done
