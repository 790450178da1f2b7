#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] attention tests incl. the opt-in variants"; V6B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_attention_bwd.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -14 | cut -c1-220
echo "== [2] BERT-base: V read in place (MN-major B operand) vs transposed copy"
V6B200_ATTN_V=mn timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
echo "== [3] attention kernel bench, V in place"; V6B200_ATTN_V=mn timeout 300 python scripts/kernel_bench.py --only attn 2>&1 | grep "^{" | cut -c1-260
