#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== conv tests (EPI_RED unit + model level)"; V6B200_TEST_BN_RED=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -40 | cut -c1-260
