#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] gemm tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm2.py tests/test_gpu_linear_bwd.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -12 | cut -c1-260
echo "== [2] gemm timing on the Llama shapes, split-K tail on / off"
python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, ".")
from vantage6_b200.ops import gemm as G
dev = torch.device("cuda", 0)
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for (M, N, K) in [(1024, 14336, 4096), (1024, 4096, 14336), (1024, 4096, 4096), (1024, 1024, 4096), (1024, 128256, 4096), (4096, 3072, 768), (4096, 768, 3072)]:
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16); w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    ms_split = timeit(lambda: G.gemm_bf16(x, w, variant="1cta"))
    os.environ["V6B200_GEMM_SPLITK"] = "0"; G._splitk_ws.clear()
    ms_plain = timeit(lambda: G.gemm_bf16(x, w, variant="1cta"))
    os.environ["V6B200_GEMM_SPLITK"] = "1"
    ms_cublas = timeit(lambda: torch.mm(x, w.t()))
    print(json.dumps({"M": M, "N": N, "K": K, "split_ms": round(ms_split, 4), "plain_ms": round(ms_plain, 4), "cublas_ms": round(ms_cublas, 4),
                      "tflops_split": round(2.0 * M * N * K / ms_split / 1e9, 1), "vs_cublas": round(ms_cublas / ms_split, 3)}), flush=True)
PY
echo "== [3] Llama-3 8B LoRA / BERT-base rounds"
timeout 400 python bench.py --model llama3_8b_lora --steps 4 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
