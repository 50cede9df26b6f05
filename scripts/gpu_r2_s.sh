#!/bin/bash
timeout 600 python -m pytest tests/test_featurize_graph_gpu.py tests/test_featurizer_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1; tail -3 gpurun_out/s_pytest.log; grep -n "^E " gpurun_out/s_pytest.log | head -6 | cut -c1-300
