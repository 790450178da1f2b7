#!/bin/bash
# 1-GPU: GEMM-focused check after an epilogue / pipeline change.
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gemm.log
echo "== kernel bench gemm"; timeout 600 python scripts/kernel_bench.py --only gemm > gpurun_out/kernel_bench_gemm.log 2>&1; echo "rc=$?"; grep -E "^\{'M'" gpurun_out/kernel_bench_gemm.log | cut -c1-420
cp gpurun_out/kernel_bench.json gpurun_out/kernel_bench_gemm.json 2>/dev/null
echo "== conv1x1"; timeout 300 python scripts/conv1x1_bench.py gpurun_out/conv1x1_bench.json 2>&1 | tail -14 | cut -c1-330
