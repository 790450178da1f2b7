#!/bin/bash
# 1-GPU: full GPU test suite, kernel bench (gemm 1cta/2cta, attention fwd/bwd vs flash-attn 2, norms...), secondary model benches.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== kernel bench"; timeout 900 python scripts/kernel_bench.py --only gemm,attn,norm,glm > gpurun_out/kernel_bench.log 2>&1; echo "rc=$?"; grep -E "^\{'M'|^\{'B'|^\{'rows'" gpurun_out/kernel_bench.log | cut -c1-420
rm -f gpurun_out/models_1.jsonl
for m in bert_base llama3_8b_lora glm; do for impl in b200 nccl; do
  echo "== model $m $impl"; timeout 600 python scripts/bench_models.py --model $m --impl $impl --rounds 5 --warmup 3 --out gpurun_out/models_1.jsonl 2> gpurun_out/model_${m}_${impl}.err | cut -c1-500; tail -2 gpurun_out/model_${m}_${impl}.err | cut -c1-300
done; done
