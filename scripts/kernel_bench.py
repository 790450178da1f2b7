#!/usr/bin/env python
"""Per-kernel roofline numbers on ONE B200 (CUDA events, >= 3 warm-ups, inputs larger than L2 or
L2 flushed between iterations), reported as achieved fraction of the MEASURED peaks in
MEASURED_PEAKS.json (fallback 6.65 TB/s / 1.59 PFLOP/s).  Writes gpurun_out/kernel_bench.json.

    python scripts/kernel_bench.py [--only gemm,norm,optim,glm,rope,k2]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=20, warm=5, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters            # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="gemm,norm,optim,glm,rope,k2")
    args = ap.parse_args()
    only = set(args.only.split(","))
    from vantage6_b200.ops import gemm as G
    from vantage6_b200.ops import glm as K8
    from vantage6_b200.ops import native, stream_ptr
    from vantage6_b200.ops import norm as N
    from vantage6_b200.ops import optim as O
    from vantage6_b200.ops import rope as R

    native()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    hbm, tflops, src = peaks()
    out = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tflops, "source": src}, "gpu": torch.cuda.get_device_name(0)}

    if "gemm" in only:
        rows = []
        for (M, Nn, K) in [(8192, 8192, 8192), (4096, 2304, 768), (4096, 3072, 768), (4096, 768, 3072), (16384, 4096, 4096),
                           (2048, 14336, 4096), (2048, 4096, 14336)]:
            a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            w = torch.randn(Nn, K, device=dev, dtype=torch.bfloat16)
            c = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
            ms = timeit(lambda: G.gemm_bf16(a, w, out=c, variant="1cta"), flush=False)
            ms_ref = timeit(lambda: torch.matmul(a, w.t(), out=c), flush=False)
            fl = 2.0 * M * Nn * K
            rows.append({"M": M, "N": Nn, "K": K, "ms": ms, "tflops": fl / ms / 1e9, "frac_of_peak": fl / ms / 1e9 / tflops,
                         "cublas_ms": ms_ref, "cublas_tflops": fl / ms_ref / 1e9, "vs_cublas": ms_ref / ms})
            if os.environ.get("V6B200_BENCH_2CTA", "1") == "1":
                try:
                    ms2 = timeit(lambda: G.gemm_bf16(a, w, out=c, variant="2cta"), flush=False)
                    rows[-1].update({"ms_2cta": ms2, "tflops_2cta": fl / ms2 / 1e9, "frac_of_peak_2cta": fl / ms2 / 1e9 / tflops})
                except Exception as e:  # noqa: BLE001
                    rows[-1]["error_2cta"] = repr(e)
            print(rows[-1], flush=True)
        out["gemm_tcgen05"] = rows

    if "norm" in only:
        rows = []
        for (r, c, rms) in [(32768, 768, False), (16384, 4096, True), (65536, 1024, False)]:
            x = torch.randn(r, c, device=dev, dtype=torch.bfloat16)
            res = torch.randn_like(x)
            g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            fn = (lambda: N.rms_norm(x, g, 1e-5, res)) if rms else (lambda: N.layer_norm(x, g, b, 1e-5, res))
            with torch.no_grad():
                ms = timeit(fn)
            byts = r * c * 2 * 4          # read x, res; write y, res_out
            tfn = (lambda: torch.nn.functional.rms_norm((x + res), (c,), g.to(x.dtype), 1e-5)) if rms else \
                (lambda: torch.nn.functional.layer_norm(x + res, (c,), g.to(x.dtype), b.to(x.dtype), 1e-5))
            with torch.no_grad():
                ms_t = timeit(tfn)
            rows.append({"rows": r, "cols": c, "rms": rms, "fwd_ms": ms, "GBps": byts / ms / 1e6, "frac_of_hbm": byts / ms / 1e6 / hbm,
                         "torch_ms": ms_t, "vs_torch": ms_t / ms})
            xg = x.clone().requires_grad_()
            gg = g.clone().requires_grad_()
            y, h = (N.rms_norm(xg, gg, 1e-5, None) if rms else N.layer_norm(xg, gg, b.clone().requires_grad_(), 1e-5, None))
            dy = torch.randn_like(y)
            msb = timeit(lambda: torch.autograd.grad(y, xg, dy, retain_graph=True))
            rows[-1].update({"bwd_ms": msb, "bwd_GBps": r * c * 2 * 3 / msb / 1e6})
            print(rows[-1], flush=True)
        out["norm"] = rows

    if "optim" in only:
        n = 110_000_000 // 8 * 8
        w, gr = torch.randn(n, device=dev), torch.randn(n, device=dev)
        sgd = O.FlatSGD(w, lr=0.1, momentum=0.9, weight_decay=1e-4)
        ms = timeit(lambda: sgd.step(gr))
        p = torch.nn.Parameter(w.clone())
        p.grad = gr
        tsgd = torch.optim.SGD([p], lr=0.1, momentum=0.9, weight_decay=1e-4)
        ms_t = timeit(lambda: tsgd.step())
        adam = O.FlatAdamW(w, lr=1e-3)
        sh = torch.empty(n, device=dev, dtype=torch.bfloat16)
        up = torch.empty(n, device=dev, dtype=torch.bfloat16)
        ref = torch.zeros(n, device=dev)
        msa = timeit(lambda: adam.step(gr))
        msaf = timeit(lambda: adam.step(gr, shadow=sh, upload=up, w_ref=ref, publish=O.PUBLISH_DELTA_BF16))
        tadam = torch.optim.AdamW([p], lr=1e-3, fused=True)
        ms_ta = timeit(lambda: tadam.step())
        out["optim"] = {"n": n, "sgd_ms": ms, "sgd_GBps": n * 4 * 5 / ms / 1e6, "sgd_frac_of_hbm": n * 20 / ms / 1e6 / hbm,
                        "torch_sgd_ms": ms_t, "adamw_ms": msa, "adamw_GBps": n * 4 * 7 / msa / 1e6,
                        "adamw_frac_of_hbm": n * 28 / msa / 1e6 / hbm, "adamw_fused_publish_shadow_ms": msaf,
                        "torch_fused_adamw_ms": ms_ta}
        print(out["optim"], flush=True)
        del w, gr, p, sh, up, ref

    if "glm" in only:
        rows, F = 1_000_000, 256
        X = torch.randn(rows, F, device=dev, dtype=torch.bfloat16)
        y = (torch.rand(rows, device=dev) < 0.5).float()
        wv = torch.randn(F + 1, device=dev) * 0.1
        o = torch.zeros(K8.payload_len(F), device=dev)
        sc = torch.empty(148 * 4 * (F + 2), device=dev)
        ms = timeit(lambda: K8.logistic_grad(X, y, wv, o, sc))
        Xf = X.float()

        def ref():
            z = Xf @ wv[:F] + wv[F]
            r = torch.sigmoid(z) - y
            return Xf.t() @ r
        ms_t = timeit(ref)
        out["glm"] = {"rows": rows, "F": F, "ms": ms, "GBps": rows * F * 2 / ms / 1e6, "frac_of_hbm": rows * F * 2 / ms / 1e6 / hbm,
                      "torch_fp32_two_gemv_ms": ms_t}
        print(out["glm"], flush=True)

    if "rope" in only:
        B, S, Hq, Hkv, D = 4, 4096, 32, 8, 128
        q = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
        k = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
        cos, sin = R.rope_tables(S, D, device=dev)
        C = native()
        ms = timeit(lambda: C.rope(q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(), 0, B, S, Hq, Hkv, D, False, stream_ptr()))
        byts = (q.numel() + k.numel()) * 2 * 2
        ms_t = timeit(lambda: R.reference_rope(q, k, cos, sin))
        out["rope"] = {"ms": ms, "GBps": byts / ms / 1e6, "frac_of_hbm": byts / ms / 1e6 / hbm, "torch_ms": ms_t}
        print(out["rope"], flush=True)

    if "attn" in only:
        from vantage6_b200.ops import attention as A

        rows = []
        for (B, S, Hq, Hkv, D, causal) in [(32, 128, 12, 12, 64, False), (8, 512, 12, 12, 64, False), (4, 2048, 32, 8, 128, True),
                                           (2, 4096, 32, 8, 128, True), (8, 1024, 32, 8, 128, False)]:
            q = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
            k = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
            v = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
            fl = 4.0 * B * Hq * S * S * D * (0.5 if causal else 1.0)
            ms = timeit(lambda: A.flash_attn_fwd(q, k, v, causal, variant="1cta"), flush=False)
            row = {"B": B, "S": S, "Hq": Hq, "Hkv": Hkv, "D": D, "causal": causal, "fwd_ms": ms, "fwd_tflops": fl / ms / 1e9}
            try:
                ms2 = timeit(lambda: A.flash_attn_fwd(q, k, v, causal, variant="2cta"), flush=False)
                row.update(fwd2_ms=ms2, fwd2_tflops=fl / ms2 / 1e9)
            except Exception as e:  # noqa: BLE001
                row["fwd2_error"] = repr(e)[:100]
            try:
                from flash_attn import flash_attn_func

                ms_fa = timeit(lambda: flash_attn_func(q, k, v, causal=causal), flush=False)
                row.update(flash_attn2_ms=ms_fa, flash_attn2_tflops=fl / ms_fa / 1e9, vs_flash_attn2=ms_fa / ms)
            except Exception as e:  # noqa: BLE001
                row["flash_attn2_error"] = repr(e)[:100]
            qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2).repeat_interleave(Hq // Hkv, 1), v.transpose(1, 2).repeat_interleave(Hq // Hkv, 1)
            ms_sd = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=causal), flush=False)
            row.update(sdpa_ms=ms_sd, sdpa_tflops=fl / ms_sd / 1e9)
            try:
                o, lse = A.flash_attn_fwd(q, k, v, causal)
                do = torch.randn_like(o)
                msb = timeit(lambda: A.flash_attn_bwd(do, q, k, v, o, lse, causal), flush=False)
                row.update(bwd_native_ms=msb, bwd_native_tflops=2.5 * fl / msb / 1e9)
                from flash_attn.flash_attn_interface import _flash_attn_backward

                dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
                sc = 1.0 / (D ** 0.5)
                msl = timeit(lambda: _flash_attn_backward(do, q, k, v, o, lse, dq, dk, dv, 0.0, sc, causal, -1, -1, 0.0, None, False, None), flush=False)
                row.update(bwd_flash_attn2_ms=msl, bwd_flash_attn2_tflops=2.5 * fl / msl / 1e9)
            except Exception as e:  # noqa: BLE001
                row["bwd_native_error"] = repr(e)[:100]
            rows.append(row)
            print(row, flush=True)
        out["attention"] = rows

    if "k2" in only:
        from vantage6_b200.parallel.fedavg import FedAvgEngine, ServerOptConfig

        n = 25_610_152
        rows = {}
        for opt in ("fedavg", "fedadam"):
            eng = FedAvgEngine(n, 0, 1, dev, data_plane="native", server_opt=ServerOptConfig(opt, 1.0))
            eng.w.normal_()
            eng.initialize_global()
            ms = timeit(lambda: eng.aggregate(1.0))
            traffic = eng.n * 4 * (4 if opt == "fedavg" else 8)        # read w, w_global; write w_global, w (+ m, v r/w)
            rows[opt] = {"ms": ms, "GBps": traffic / ms / 1e6, "frac_of_hbm": traffic / ms / 1e6 / hbm}
            eng.close()
        out["k2_world1"] = rows
        print(rows, flush=True)

    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kernel_bench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
