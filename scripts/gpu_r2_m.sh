#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_oracle_kd_pinned.py tests/test_bf16_parity_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "kl or kd or makd or config1 or config2" 2>&1 | tail -3
for dt in bf16 f32; do echo "== makd_micro $dt"; timeout 300 python scripts/makd_micro.py $dt 2>&1 | grep -v "^$" | tail -8; done | tee gpurun_out/m_makd_micro2.log
