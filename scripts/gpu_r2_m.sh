#!/bin/bash
for dt in bf16 f32; do echo "== makd_micro $dt"; timeout 300 python scripts/makd_micro.py $dt 2>&1 | grep -v "^$" | tail -12; done | tee gpurun_out/m_makd_micro.log
