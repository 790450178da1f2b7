#!/bin/bash
# Round 2, 1-GPU call 10: BatchNorm backward reduction in the data-gradient epilogue (EPI_RED)
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] conv / bn / resnet tests"; timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bn.py tests/test_gpu_resnet_ops.py tests/test_gpu_linear_bwd.py -q -m gpu --timeout 300 2>&1 | tail -25 | cut -c1-300
echo "== [2] bench (reduction fused)"; V6B200_BN_RED=1 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [3] bench (reduction as its own pass, default)"; timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [4] smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
