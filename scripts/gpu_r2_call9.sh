#!/bin/bash
# Round 2, 1-GPU call 9: filter gradients on a second stream (overlap with the BatchNorm backward passes)
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] pytest conv + bn + resnet ops"; timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bn.py tests/test_gpu_resnet_ops.py -q -m gpu --timeout 200 2>&1 | tail -4 | cut -c1-300
echo "== [2] bench, side-stream wgrad (default) + comparator arms"; timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_r2g.json; tail -3 gpurun_out/bench_r2g.err
echo "== [3] bench, in-line wgrad"; V6B200_WGRAD_STREAM=0 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [4] smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
