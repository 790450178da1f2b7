#!/bin/bash
# Round 2, GPU call 1 (1 GPU): numerics of the new implicit-GEMM convolution kernels, then acceptance of the round-1
# state and the opt-in items written after the round-1 GPU budget was spent.
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== conv probe (numerics)"; timeout 900 python scripts/conv_probe.py --out gpurun_out/conv_probe_r2a.jsonl 2>&1 | cut -c1-400
echo "== pytest -m gpu (default suite)"; timeout 400 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== opt-in items"; V6B200_EXPERIMENTAL=1 timeout 200 python -m pytest tests -q -m gpu -s -k "experimental or glm_kernel_timing" --timeout 120 2>&1 | tail -15 | cut -c1-300
echo "== attention with V in place (bench)"; V6B200_ATTN_V=mn timeout 200 python scripts/kernel_bench.py --only attn 2>&1 | grep "^{'B'" | cut -c1-260
echo "== federated GLM, fused iteration"; V6B200_GLM_FUSED=1 timeout 100 python scripts/bench_models.py --model glm --impl b200 --rounds 5 --warmup 3 2>/dev/null | cut -c1-330
echo "== federated GLM, default"; timeout 100 python scripts/bench_models.py --model glm --impl b200 --rounds 5 --warmup 3 2>/dev/null | cut -c1-330
