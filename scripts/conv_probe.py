"""Numerics + timing probe of the tcgen05 implicit-GEMM convolution kernels (csrc/igemm.cu) against fp32 PyTorch
references.  Every group runs in its own process (a kernel trap in one group must not take the others down) and prints
one JSON line per case to stdout; ``--out`` collects them.

    PYTHONPATH=. python scripts/conv_probe.py --out gpurun_out/conv_probe.jsonl [--groups fprop1x1,fprop3x3,...] [--time]
"""
from __future__ import annotations

import argparse
import json
import subprocess
import sys
import time

GROUPS = ["dgrad_strided", "stem", "fprop1x1", "fprop1x1_im2col", "fprop3x3", "fprop_strided", "fprop_stats", "dgrad1x1", "dgrad3x3", "wgrad1x1", "wgrad3x3",
          "wgrad_strided", "linear"]


def _rel(a, b):
    import torch

    d = (a.float() - b.float()).abs().max().item()
    s = b.float().abs().max().item()
    return d, d / max(s, 1e-6)


def _mk(n, c, h, w, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn((n, c, h, w), device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def _time(fn, iters=20):
    import torch

    for _ in range(3):
        fn()
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    evs = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3


def run_group(group: str, do_time: bool):
    import torch
    import torch.nn.functional as F

    from vantage6_b200.ops import conv as C

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = []

    def emit(**kw):
        kw["group"] = group
        out.append(kw)
        if group != "resnet50":
            print(json.dumps(kw), flush=True)

    def fprop_case(n, cin, cout, h, w, r, stride, pad, force=False, stats=False, seed=0):
        x = _mk(n, cin, h, w, seed)
        wt = (_mk(cout, cin, r, r, seed + 1) * (1.0 / (cin * r * r) ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        bn = None
        if stats:
            bn = dict(gamma=torch.rand(cout, device="cuda") + 0.5, beta=torch.randn(cout, device="cuda"),
                      running_mean=torch.zeros(cout, device="cuda"), running_var=torch.ones(cout, device="cuda"),
                      num_batches_tracked=torch.zeros((), device="cuda", dtype=torch.long),
                      mean=torch.empty(cout, device="cuda"), rstd=torch.empty(cout, device="cuda"),
                      scale_bias=torch.empty(2 * cout, device="cuda"), eps=1e-5, momentum=0.1)
        y = C.conv_fprop(x, wt, stride, pad, bn=bn, force_im2col=force)
        torch.cuda.synchronize()
        ref = F.conv2d(x.float(), wt.float(), stride=stride, padding=pad)
        d, rel = _rel(y, ref)
        rec = dict(case=f"fprop n{n} {cin}->{cout} {h}x{w} k{r} s{stride} p{pad} force={int(force)}", max_abs=d, rel=rel, ok=rel < 2e-2)
        if stats:
            yf = y.float()
            m_ref = yf.mean(dim=(0, 2, 3))
            v_ref = yf.var(dim=(0, 2, 3), unbiased=False)
            rec["mean_err"] = (bn["mean"] - m_ref).abs().max().item()
            rec["rstd_rel"] = ((bn["rstd"] - torch.rsqrt(v_ref + 1e-5)).abs() / torch.rsqrt(v_ref + 1e-5)).max().item()
            sc = bn["gamma"] * torch.rsqrt(v_ref + 1e-5)
            rec["scale_err"] = (bn["scale_bias"][:cout] - sc).abs().max().item()
            rec["bias_err"] = (bn["scale_bias"][cout:] - (bn["beta"] - m_ref * sc)).abs().max().item()
            rec["rm_err"] = (bn["running_mean"] - 0.1 * m_ref).abs().max().item()
            cnt = yf.numel() // cout
            rec["rv_err"] = (bn["running_var"] - (0.9 + 0.1 * v_ref * cnt / (cnt - 1))).abs().max().item()
            rec["nbt"] = int(bn["num_batches_tracked"].item())
            rec["ok"] = bool(rec["ok"] and rec["mean_err"] < 2e-3 and rec["rstd_rel"] < 2e-3 and rec["rm_err"] < 1e-3 and rec["nbt"] == 1)
        if do_time:
            rec["us"] = _time(lambda: C.conv_fprop(x, wt, stride, pad, bn=bn, force_im2col=force))
            xx, ww = x, wt
            rec["cudnn_us"] = _time(lambda: torch.ops.aten.convolution(xx, ww, None, (stride, stride), (pad, pad), (1, 1), False, (0, 0), 1))
            p = (h + 2 * pad - r) // stride + 1
            rec["tflops"] = 2.0 * n * p * p * cout * cin * r * r / rec["us"] * 1e-6
        emit(**rec)

    def dgrad_case(n, cin, cout, h, w, r, pad, force=False, seed=0, stride=1):
        p, q = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - r) // stride + 1
        dy = _mk(n, cout, p, q, seed)
        wt = (_mk(cout, cin, r, r, seed + 1) * (1.0 / (cout * r * r) ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = C.conv_dgrad(dy, wt, (h, w), pad, force_im2col=force, stride=stride)
        torch.cuda.synchronize()
        ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt.float(), dy.float(), stride=stride, padding=pad)
        d, rel = _rel(dx, ref)
        rec = dict(case=f"dgrad n{n} {cin}<-{cout} {h}x{w} k{r} s{stride} p{pad} force={int(force)}", max_abs=d, rel=rel, ok=rel < 2e-2)
        if do_time:
            rec["us"] = _time(lambda: C.conv_dgrad(dy, wt, (h, w), pad, force_im2col=force, stride=stride))
            x0 = _mk(n, cin, h, w, seed + 5)
            rec["cudnn_us"] = _time(lambda: torch.ops.aten.convolution_backward(dy, x0, wt, None, (stride, stride), (pad, pad), (1, 1), False, (0, 0), 1, (True, False, False)))
            rec["tflops"] = 2.0 * n * h * w * cout * cin * r * r / rec["us"] * 1e-6
        emit(**rec)

    def wgrad_case(n, cin, cout, h, w, r, stride, pad, force=False, splits=0, seed=0):
        p, q = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - r) // stride + 1
        x = _mk(n, cin, h, w, seed)
        dy = (_mk(n, cout, p, q, seed + 1) * 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dw = torch.zeros((cout, r, r, cin), device="cuda", dtype=torch.float32)
        C.conv_wgrad(dy, x, dw, (r, r), stride, pad, splits=splits, force_im2col=force)
        torch.cuda.synchronize()
        ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, r, r), dy.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
        d, rel = _rel(dw, ref)
        rec = dict(case=f"wgrad n{n} {cin}->{cout} {h}x{w} k{r} s{stride} p{pad} force={int(force)} splits={splits}", max_abs=d, rel=rel,
                   ok=rel < 1e-2)
        if do_time:
            rec["us"] = _time(lambda: C.conv_wgrad(dy, x, dw, (r, r), stride, pad, splits=splits, force_im2col=force))
            wt = _mk(cout, cin, r, r, seed + 3)
            rec["cudnn_us"] = _time(lambda: torch.ops.aten.convolution_backward(dy, x, wt, None, (stride, stride), (pad, pad), (1, 1), False, (0, 0), 1, (False, True, False)))
            rec["tflops"] = 2.0 * n * p * q * cout * cin * r * r / rec["us"] * 1e-6
        emit(**rec)

    def stem_case(n, hw, cout=64, seed=0):
        from vantage6_b200.ops import pool as PL

        hs = hw // 2 + 3
        xs = _mk(n, 16, hs, hs, seed)
        ws = (_mk(cout, 16, 4, 4, seed + 1) * (1.0 / 256 ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        y = C.stem_fprop(xs, ws)
        torch.cuda.synchronize()
        ref = F.conv2d(xs.float(), ws.float())
        d, rel = _rel(y, ref)
        dy = (_mk(n, cout, hs - 3, hs - 3, seed + 2) * 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dws = torch.zeros((cout, 4, 4, 16), device="cuda")
        C.stem_wgrad(dy, xs, dws)
        torch.cuda.synchronize()
        refw = torch.nn.grad.conv2d_weight(xs.float(), (cout, 16, 4, 4), dy.float()).permute(0, 2, 3, 1)
        d2, rel2 = _rel(dws, refw)
        rec = dict(case=f"stem n{n} {hw}x{hw} -> {cout}", max_abs=d, rel=rel, wgrad_rel=rel2, ok=rel < 2e-2 and rel2 < 1e-2)
        if do_time:
            rec["us"] = _time(lambda: C.stem_fprop(xs, ws))
            rec["cudnn_us"] = _time(lambda: torch.ops.aten.convolution(xs, ws, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1))
            rec["wgrad_us"] = _time(lambda: C.stem_wgrad(dy, xs, dws))
            rec["cudnn_wgrad_us"] = _time(lambda: torch.ops.aten.convolution_backward(dy, xs, ws, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (False, True, False)))
        emit(**rec)

    B = 64 if do_time else 8
    if group == "dgrad_strided":
        dgrad_case(2, 64, 64, 16, 16, 3, 1, stride=2)
        dgrad_case(2, 64, 128, 16, 16, 1, 0, stride=2)
        dgrad_case(B, 128, 128, 56, 56, 3, 1, stride=2)
        dgrad_case(B, 256, 512, 56, 56, 1, 0, stride=2)
        dgrad_case(B, 512, 512, 14, 14, 3, 1, stride=2)
        dgrad_case(B, 1024, 2048, 14, 14, 1, 0, stride=2)
    elif group == "stem":
        stem_case(2, 32)
        stem_case(B, 224)
    elif group == "fprop1x1":
        fprop_case(2, 64, 64, 8, 8, 1, 1, 0)            # one tile exactly (128 pixels)
        fprop_case(B, 64, 256, 56, 56, 1, 1, 0)
        fprop_case(B, 256, 64, 56, 56, 1, 1, 0)
        fprop_case(B, 512, 128, 28, 28, 1, 1, 0)
        fprop_case(B, 1024, 2048, 7, 7, 1, 1, 0)        # partial last row tile
    elif group == "fprop1x1_im2col":
        fprop_case(2, 64, 64, 8, 8, 1, 1, 0, force=True)
        fprop_case(B, 256, 64, 56, 56, 1, 1, 0, force=True)
        fprop_case(3, 128, 128, 7, 7, 1, 1, 0, force=True)
    elif group == "fprop3x3":
        fprop_case(2, 64, 64, 8, 8, 3, 1, 1)
        fprop_case(B, 64, 64, 56, 56, 3, 1, 1)
        fprop_case(B, 128, 128, 28, 28, 3, 1, 1)
        fprop_case(B, 256, 256, 14, 14, 3, 1, 1)
        fprop_case(B, 512, 512, 7, 7, 3, 1, 1)
    elif group == "fprop_strided":
        fprop_case(2, 64, 64, 16, 16, 3, 2, 1)
        fprop_case(B, 128, 128, 56, 56, 3, 2, 1)
        fprop_case(B, 256, 512, 56, 56, 1, 2, 0)
        fprop_case(B, 512, 512, 14, 14, 3, 2, 1)
    elif group == "fprop_stats":
        fprop_case(2, 64, 64, 8, 8, 1, 1, 0, stats=True)
        fprop_case(B, 64, 256, 56, 56, 1, 1, 0, stats=True)
        fprop_case(B, 128, 128, 28, 28, 3, 1, 1, stats=True)
        fprop_case(B, 1024, 2048, 7, 7, 1, 1, 0, stats=True)
        fprop_case(64, 64, 64, 56, 56, 3, 1, 1, stats=True)      # long runs of row tiles per CTA
    elif group == "dgrad1x1":
        dgrad_case(2, 64, 64, 8, 8, 1, 0)
        dgrad_case(B, 64, 256, 56, 56, 1, 0)
        dgrad_case(B, 256, 64, 56, 56, 1, 0)
        dgrad_case(B, 2048, 512, 7, 7, 1, 0)
    elif group == "dgrad3x3":
        dgrad_case(2, 64, 64, 8, 8, 3, 1)
        dgrad_case(B, 64, 64, 56, 56, 3, 1)
        dgrad_case(B, 128, 128, 28, 28, 3, 1)
        dgrad_case(B, 512, 512, 7, 7, 3, 1)
    elif group == "wgrad1x1":
        wgrad_case(2, 64, 64, 8, 8, 1, 1, 0, splits=1)
        wgrad_case(2, 64, 128, 8, 8, 1, 1, 0, splits=1)
        wgrad_case(B, 64, 256, 56, 56, 1, 1, 0)
        wgrad_case(B, 256, 64, 56, 56, 1, 1, 0)
        wgrad_case(B, 1024, 2048, 7, 7, 1, 1, 0)
    elif group == "wgrad3x3":
        wgrad_case(2, 64, 64, 8, 8, 3, 1, 1, splits=1)
        wgrad_case(B, 64, 64, 56, 56, 3, 1, 1)
        wgrad_case(B, 128, 128, 28, 28, 3, 1, 1)
        wgrad_case(B, 512, 512, 7, 7, 3, 1, 1)
    elif group == "wgrad_strided":
        wgrad_case(2, 64, 64, 16, 16, 3, 2, 1, splits=1)
        wgrad_case(B, 128, 128, 56, 56, 3, 2, 1)
        wgrad_case(B, 256, 512, 56, 56, 1, 2, 0)
    elif group == "profile":
        # the launches captured by `ncu --set full -k regex:igemm_kernel` (scripts/gpu_r2_prof_conv.sh), in this order
        fprop_case(64, 64, 256, 56, 56, 1, 1, 0)
        fprop_case(64, 64, 64, 56, 56, 3, 1, 1)
        fprop_case(64, 64, 64, 56, 56, 3, 1, 1, stats=True)
        fprop_case(64, 512, 512, 7, 7, 3, 1, 1)
        dgrad_case(64, 64, 64, 56, 56, 3, 1)
        wgrad_case(64, 64, 64, 56, 56, 3, 1, 1)
        wgrad_case(64, 64, 256, 56, 56, 1, 1, 0)
        wgrad_case(64, 512, 512, 7, 7, 3, 1, 1)
    elif group == "linear":
        # nn.Linear shapes through the same kernels: [tokens, K] x [N, K]^T (1x1 convolution over a 1-pixel-high image)
        fprop_case(1, 768, 2304, 1, 4096, 1, 1, 0)
        dgrad_case(1, 768, 3072, 1, 4096, 1, 0)
        wgrad_case(1, 768, 3072, 1, 4096, 1, 1, 0)
        fprop_case(1, 2048, 1000, 1, 64, 1, 1, 0)       # ResNet fc: N not a multiple of 64
    elif group == "resnet50":
        # every convolution of ResNet-50 (batch 64, 224x224; stem separately) with its multiplicity: the per-step accounting
        # that scripts/conv_layer_table.py turns into "sum of our kernels vs sum of cuDNN's" (profiles/)
        for (cin, cout, hw, r, stride, mult) in R50_LAYERS:
            pad = 1 if r == 3 else 0
            n0 = len(out)
            fprop_case(64, cin, cout, hw, hw, r, stride, pad)
            fprop_case(64, cin, cout, hw, hw, r, stride, pad, stats=True)
            if C.dgrad_supported(cin, cout, r, r, stride, pad, hw, hw):
                dgrad_case(64, cin, cout, hw, hw, r, pad, stride=stride)
            wgrad_case(64, cin, cout, hw, hw, r, stride, pad)
            for rec in out[n0:]:
                rec["mult"] = mult
        stem_case(64, 224)
    if group == "resnet50":
        for kw in out:
            print(json.dumps(kw), flush=True)
    return out


# (Cin, Cout, input H=W, filter, stride, how many times per forward pass)
R50_LAYERS = [
    (64, 64, 56, 1, 1, 1), (64, 64, 56, 3, 1, 3), (64, 256, 56, 1, 1, 4), (256, 64, 56, 1, 1, 2),
    (256, 128, 56, 1, 1, 1), (128, 128, 56, 3, 2, 1), (128, 512, 28, 1, 1, 4), (256, 512, 56, 1, 2, 1), (512, 128, 28, 1, 1, 3), (128, 128, 28, 3, 1, 3),
    (512, 256, 28, 1, 1, 1), (256, 256, 28, 3, 2, 1), (256, 1024, 14, 1, 1, 6), (512, 1024, 28, 1, 2, 1), (1024, 256, 14, 1, 1, 5), (256, 256, 14, 3, 1, 5),
    (1024, 512, 14, 1, 1, 1), (512, 512, 14, 3, 2, 1), (512, 2048, 7, 1, 1, 3), (1024, 2048, 14, 1, 2, 1), (2048, 512, 7, 1, 1, 2), (512, 512, 7, 3, 1, 2),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", default=",".join(GROUPS))
    ap.add_argument("--out", default="")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--child", default="")
    a = ap.parse_args()
    if a.child:
        run_group(a.child, a.time)
        return
    lines = []
    for g in a.groups.split(","):
        t0 = time.time()
        cmd = [sys.executable, __file__, "--child", g] + (["--time"] if a.time else [])
        try:
            pr = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900 if g == "resnet50" else 240)
            rc, so, se = pr.returncode, pr.stdout, pr.stderr
        except subprocess.TimeoutExpired as e:
            rc, so, se = -9, (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "timeout"
        for ln in so.splitlines():
            if ln.startswith("{"):
                lines.append(ln)
                print(ln)
        if rc != 0:
            rec = json.dumps({"group": g, "crashed": rc, "stderr_tail": se[-600:]})
            lines.append(rec)
            print(rec)
        print(f"# group {g}: rc={rc} {time.time() - t0:.1f}s", flush=True)
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")
    recs = [json.loads(x) for x in lines]
    bad = [r for r in recs if not r.get("ok", False)]
    print(f"# {len(recs) - len(bad)}/{len(recs)} ok")


if __name__ == "__main__":
    main()
