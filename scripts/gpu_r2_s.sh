#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_featurize_graph_gpu.py tests/test_featurizer_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/s_pytest.log | cut -c1-250
