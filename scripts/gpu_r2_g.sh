#!/bin/bash
set -u
run() { echo "== $*"; timeout 600 env $1 python bench.py --timed-only --steps 40 ${@:2} 2>&1 | grep "timed-only"; }
run CUDA_DEVICE_MAX_CONNECTIONS=32 --pipeline 1
run CUDA_DEVICE_MAX_CONNECTIONS=32 --pipeline 0
run CUDA_DEVICE_MAX_CONNECTIONS=32 --pipeline 1 --teacher-sms 120
run CUDA_DEVICE_MAX_CONNECTIONS=1 --pipeline 1
run CUDA_DEVICE_MAX_CONNECTIONS=32 --pipeline 1 --workload magic_s_pretrain_b64
run CUDA_DEVICE_MAX_CONNECTIONS=32 --pipeline 1 --workload magic_l_icod_b32
run X=1 --pipeline 1 --branch-streams 0 --side-stream 0
run X=1 --pipeline 0 --branch-streams 0 --side-stream 0
run MAGIC_PDL=0 --pipeline 1
