#!/bin/bash
timeout 600 python -m pytest tests/test_adamw_golden_gpu.py tests/test_stepper_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --timed-only --steps 20 2>&1 | grep "timed-only"
