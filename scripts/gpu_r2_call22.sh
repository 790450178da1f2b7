#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] linear / llama tests"; timeout 600 python -m pytest tests/test_gpu_linear_bwd.py -q -m gpu --timeout 300 2>&1 | tail -3 | cut -c1-260
echo "== [2] Llama-3 8B LoRA"; timeout 400 python bench.py --model llama3_8b_lora --steps 4 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
