"""1x1-convolution micro-benchmark (ResNet-50 shapes, batch 64, NHWC bf16): cuDNN implicit GEMM vs plain GEMMs
(cuBLAS via torch.mm, and the hand-written tcgen05 GEMM for the forward)."""
import json
import sys

import torch

from vantage6_b200.ops import gemm as G


def time_cuda(fn, warmup=5, iters=20):
    """ms per call, CUDA events, back-to-back (operands of these sizes are L2-resident in the training step as well)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    shapes = [(64, 64, 56), (64, 256, 56), (256, 64, 56), (256, 128, 56), (128, 512, 28), (512, 128, 28), (512, 256, 28),
              (256, 1024, 14), (1024, 256, 14), (1024, 512, 14), (512, 2048, 7), (2048, 512, 7)]
    N = 64
    out = []
    for cin, cout, hw in shapes:
        x = torch.randn(N, cin, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = torch.randn(cout, cin, 1, 1, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        y = torch.ops.aten.convolution(x, w, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1)
        dy = torch.randn_like(y)
        x2 = x.permute(0, 2, 3, 1).reshape(-1, cin)
        dy2 = dy.permute(0, 2, 3, 1).reshape(-1, cout)
        w2 = w.view(cout, cin)
        conv = lambda: torch.ops.aten.convolution(x, w, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1)
        dgrad = lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (True, False, False))
        wgrad = lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (False, True, False))
        r = {"cin": cin, "cout": cout, "hw": hw, "rows": x2.shape[0]}
        r["cudnn_fprop_us"] = 1e3 * time_cuda(conv, warmup=5, iters=20)
        r["cublas_fprop_us"] = 1e3 * time_cuda(lambda: torch.mm(x2, w2.t()), warmup=5, iters=20)
        try:
            r["tcgen05_fprop_us"] = 1e3 * time_cuda(lambda: G.gemm_bf16(x2, w2), warmup=5, iters=20)
        except Exception as e:  # noqa: BLE001
            r["tcgen05_fprop_us"] = str(e)[:60]
        r["cudnn_dgrad_us"] = 1e3 * time_cuda(dgrad, warmup=5, iters=20)
        r["cublas_dgrad_us"] = 1e3 * time_cuda(lambda: torch.mm(dy2, w2), warmup=5, iters=20)
        r["cudnn_wgrad_us"] = 1e3 * time_cuda(wgrad, warmup=5, iters=20)
        r["cublas_wgrad_us"] = 1e3 * time_cuda(lambda: torch.mm(dy2.t(), x2), warmup=5, iters=20)
        out.append(r)
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv1x1_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
