#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== linear forward: ops/gemm.py (128x256 tiles) vs implicit-GEMM kernel (N tile per problem) vs cuBLAS"
python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
from vantage6_b200.ops import gemm as G, conv as C
dev = torch.device("cuda", 0)
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
shapes = [(1024, 14336, 4096), (1024, 4096, 14336), (1024, 4096, 4096), (1024, 1024, 4096), (4096, 2304, 768), (4096, 768, 768), (4096, 3072, 768),
          (4096, 768, 3072), (614, 768, 768), (64, 1024, 2048), (2048, 4096, 4096), (8192, 8192, 8192)]
for (M, N, K) in shapes:
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16); w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    ref = x.float() @ w.float().t()
    y = C.linear_fprop(x, w)
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    r = {"M": M, "N": N, "K": K, "gemm_ms": round(timeit(lambda: G.gemm_bf16(x, w, variant="1cta")), 4), "igemm_ms": round(timeit(lambda: C.linear_fprop(x, w)), 4),
         "cublas_ms": round(timeit(lambda: torch.mm(x, w.t())), 4), "igemm_rel_err": round(err, 5)}
    if M >= 256 and N >= 256:
        r["gemm2_ms"] = round(timeit(lambda: G.gemm_bf16(x, w, variant="2cta")), 4)
    print(json.dumps(r), flush=True)
PY
