#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] tests: resnet ops (fused stem BN+pool), norms, conv"; timeout 600 python -m pytest tests/test_gpu_resnet_ops.py tests/test_gpu_kernels.py tests/test_gpu_conv.py tests/test_gpu_bn.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -14 | cut -c1-260
echo "== [2] bench, stem BN + pool fused (default)"; timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [3] bench, separate stem passes"; V6B200_STEM_POOL=separate timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [4] smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
