#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] conv tests incl. EPI_RED"; V6B200_TEST_BN_RED=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -12 | cut -c1-260
echo "== [2] bench, BN backward reduction in the dgrad epilogue"; V6B200_BN_RED=1 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [3] bench, default"; timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
