#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== linear backward kernels, same timer: dX = dY.W  (a) implicit-GEMM DGRAD (W in place, MN-major)  (b) implicit-GEMM FPROP on a transposed copy  (c) cuBLAS;  dW += dY^T.X  (d) implicit-GEMM WGRAD  (e) cuBLAS"
python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
from vantage6_b200.ops import conv as C
dev = torch.device("cuda", 0)
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for (M, Kin, Nout) in [(4096, 768, 2304), (4096, 768, 768), (4096, 768, 3072), (4096, 3072, 768), (1024, 4096, 14336), (1024, 14336, 4096), (1024, 4096, 4096), (1024, 4096, 1024)]:
    x = torch.randn(M, Kin, device=dev, dtype=torch.bfloat16); w = torch.randn(Nout, Kin, device=dev, dtype=torch.bfloat16) * 0.05
    dy = torch.randn(M, Nout, device=dev, dtype=torch.bfloat16)
    wt = w.t().contiguous()
    dw = torch.zeros(Nout, Kin, device=dev)
    ref = dy.float() @ w.float()
    e_a = float((C.linear_dgrad(dy, w).float() - ref).abs().max() / ref.abs().max())
    e_b = float((C.linear_fprop(dy, wt).float() - ref).abs().max() / ref.abs().max())
    r = {"M": M, "K_in": Kin, "N_out": Nout,
         "dx_dgrad_ms": round(timeit(lambda: C.linear_dgrad(dy, w)), 4), "dx_fprop_T_ms": round(timeit(lambda: C.linear_fprop(dy, wt)), 4),
         "dx_cublas_ms": round(timeit(lambda: torch.mm(dy, w)), 4), "transpose_ms": round(timeit(lambda: w.t().contiguous()), 4),
         "dw_wgrad_ms": round(timeit(lambda: C.linear_wgrad(dy, x, dw)), 4), "dw_cublas_ms": round(timeit(lambda: torch.mm(dy.t(), x)), 4),
         "err_dgrad": round(e_a, 5), "err_fprop_T": round(e_b, 5)}
    print(json.dumps(r), flush=True)
PY
