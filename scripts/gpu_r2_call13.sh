#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] norm tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "norm" --timeout 300 2>&1 | tail -4 | cut -c1-260
echo "== [2] norm bench: split backward (default)"; timeout 200 python scripts/kernel_bench.py --only norm 2>&1 | grep "^{" | cut -c1-330
echo "== [2b] norm bench: one-kernel backward"; V6B200_NORM_BWD=1 timeout 200 python scripts/kernel_bench.py --only norm 2>&1 | grep "^{" | cut -c1-330
echo "== [3] BERT-base, split / one-kernel norm backward"
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
V6B200_NORM_BWD=1 timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
