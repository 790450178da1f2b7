#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] gemm / linear / transformer tests (split-K reverted)"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm2.py tests/test_gpu_linear_bwd.py tests/test_gpu_attention.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -8 | cut -c1-260
echo "== [2] BERT-base / Llama-3 8B LoRA rounds, small GEMMs on the implicit-GEMM forward kernel"
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
timeout 400 python bench.py --model llama3_8b_lora --steps 4 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
echo "== [3] same, everything on csrc/gemm.cu"
V6B200_LINEAR_FWD=gemm timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
V6B200_LINEAR_FWD=gemm timeout 400 python bench.py --model llama3_8b_lora --steps 4 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
