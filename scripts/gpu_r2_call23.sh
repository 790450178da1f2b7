#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] frozen linear + fused CE tests"; timeout 600 python -m pytest tests/test_gpu_linear_bwd.py tests/test_gpu_kernels.py -q -m gpu --timeout 300 -k "frozen or cross_entropy or llama" 2>&1 | grep -v "^E   \s*+" | tail -40 | cut -c1-220
echo "== [2] Llama-3 8B LoRA with the fused cross-entropy"; timeout 400 python bench.py --model llama3_8b_lora --steps 4 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
