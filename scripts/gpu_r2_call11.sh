#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] conv tests incl. EPI_RED unit + model level"; V6B200_TEST_BN_RED=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -40 | cut -c1-260
echo "== [2] linear / llama tests"; timeout 600 python -m pytest tests/test_gpu_linear_bwd.py -q -m gpu --timeout 300 2>&1 | tail -4 | cut -c1-260
echo "== [3] Llama-3 8B LoRA: frozen dX on the K-major GEMM (transposed copy) / implicit GEMM / cuBLAS"
for mode in gemm igemm cublas; do
  LB=tc; [ "$mode" = "cublas" ] && LB=cublas
  V6B200_LINEAR_BWD=$LB V6B200_FROZEN_DX=$mode timeout 400 python bench.py --model llama3_8b_lora --steps 4 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
done
