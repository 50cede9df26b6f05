#!/bin/bash
for p in 1 2; do echo "== MAGIC_TC_PAIR=$p"; MAGIC_TC_PAIR=$p timeout 300 python scripts/graph_micro.py gemm_m 2>&1 | grep "gemm fwd\|dgrad" ; done | tee gpurun_out/w_gemm_m2.log
