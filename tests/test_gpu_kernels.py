"""Single-GPU numerics tests: every hand-written sm_100a kernel against a plain PyTorch fp32
reference of the same op (test tier (ii) of SURVEY.md section 4)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from vantage6_b200.ops import native

    native()            # loud failure if the extension is missing on a GPU box
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


# ------------------------------------------------------------------ K7 optimizers
@pytest.mark.parametrize("publish", [0, 1, 2, 3])
def test_flat_sgd_matches_reference(dev, publish):
    from vantage6_b200.ops import optim as O

    torch.manual_seed(0)
    n = 1 << 20
    w = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    w_ref_buf = torch.randn(n, device=dev)
    up_dtype = torch.bfloat16 if publish == 2 else torch.float32
    w2, buf2, ref2 = w.clone(), torch.zeros(n, device=dev), w_ref_buf.clone()
    up1, up2 = torch.zeros(n, device=dev, dtype=up_dtype), torch.zeros(n, device=dev, dtype=up_dtype)
    sh1, sh2 = torch.zeros(n, device=dev, dtype=torch.bfloat16), torch.zeros(n, device=dev, dtype=torch.bfloat16)
    opt = O.FlatSGD(w, lr=0.1, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        first = step == 0
        opt.step(g, w_ref=w_ref_buf, save_ref=first, upload=up1, publish=publish, contrib_scale=3.0, shadow=sh1)
        O.reference_sgd_step(w2, g, buf2, 0.1, 0.9, 0.0, 1e-4, False, first, w_ref=ref2, save_ref=first, upload=up2,
                             publish=publish, contrib_scale=3.0, shadow=sh2)
    torch.testing.assert_close(w, w2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(opt.buf, buf2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(w_ref_buf, ref2)
    torch.testing.assert_close(up1.float(), up2.float(), rtol=2e-2 if publish == 2 else 1e-5, atol=1e-2 if publish == 2 else 1e-5)
    torch.testing.assert_close(sh1.float(), sh2.float(), rtol=1e-2, atol=1e-2)


def test_flat_adamw_matches_torch(dev):
    from vantage6_b200.ops import optim as O

    torch.manual_seed(1)
    n = 1 << 18
    p = torch.nn.Parameter(torch.randn(n, device=dev))
    w = p.detach().clone()
    topt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    opt = O.FlatAdamW(w, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1)
    for _ in range(4):
        g = torch.randn(n, device=dev)
        p.grad = g.clone()
        topt.step()
        opt.step(g)
    torch.testing.assert_close(w, p.detach(), rtol=2e-5, atol=2e-6)


def test_clip_coef(dev):
    from vantage6_b200.ops import optim as O

    g = torch.randn(1 << 20, device=dev) * 3
    c = O.clip_grad_coef(g, 1.0)
    ref = min(1.0, 1.0 / (g.norm().item() + 1e-6))
    assert abs(c.item() - ref) / ref < 1e-3


# ------------------------------------------------------------------ K5 norms
@pytest.mark.parametrize("rows,cols", [(1000, 768), (33, 4096), (7, 64), (4096, 1024)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_fwd_bwd(dev, rows, cols, dtype, with_res):
    from vantage6_b200.ops import norm as N

    torch.manual_seed(2)
    x = torch.randn(rows, cols, device=dev, dtype=dtype, requires_grad=True)
    r = torch.randn(rows, cols, device=dev, dtype=dtype, requires_grad=True) if with_res else None
    gamma = torch.randn(cols, device=dev).add_(1.0).requires_grad_()
    beta = torch.randn(cols, device=dev).requires_grad_()
    y, h = N.layer_norm(x, gamma, beta, 1e-5, r)
    dy = torch.randn_like(y)
    dh = torch.randn_like(y) if with_res else None
    (y.float() * dy.float()).sum().add((h.float() * dh.float()).sum() if with_res else 0).backward()
    gx, gg, gb = x.grad.clone(), gamma.grad.clone(), beta.grad.clone()
    gr = r.grad.clone() if with_res else None
    # fp32 reference
    xf = x.detach().float().requires_grad_()
    rf = r.detach().float().requires_grad_() if with_res else None
    g2, b2 = gamma.detach().clone().requires_grad_(), beta.detach().clone().requires_grad_()
    hf = xf + rf if with_res else xf
    if dtype == torch.bfloat16 and with_res:
        hf = hf + (hf.to(dtype).float() - hf).detach()          # kernel rounds the residual stream to bf16
    yr = torch.nn.functional.layer_norm(hf, (cols,), g2, b2, 1e-5)
    (yr * dy.float()).sum().add((hf * dh.float()).sum() if with_res else 0).backward()
    tol = dict(rtol=3e-2, atol=3e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(y.float(), yr, **tol)
    torch.testing.assert_close(gx.float(), xf.grad, **tol)
    if with_res:
        torch.testing.assert_close(gr.float(), rf.grad, **tol)
    ptol = dict(rtol=3e-2, atol=0.3) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(gg, g2.grad, **ptol)
    torch.testing.assert_close(gb, b2.grad, **ptol)


@pytest.mark.parametrize("rows,cols", [(512, 4096), (100, 1024)])
def test_rmsnorm_fwd_bwd(dev, rows, cols):
    from vantage6_b200.ops import norm as N

    torch.manual_seed(3)
    x = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    gamma = torch.randn(cols, device=dev).add_(1.0).requires_grad_()
    y, _ = N.rms_norm(x, gamma, 1e-5)
    dy = torch.randn_like(y)
    (y.float() * dy.float()).sum().backward()
    xf = x.detach().float().requires_grad_()
    g2 = gamma.detach().clone().requires_grad_()
    yr = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * g2
    (yr * dy.float()).sum().backward()
    torch.testing.assert_close(y.float(), yr, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(x.grad.float(), xf.grad, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(gamma.grad, g2.grad, rtol=3e-2, atol=0.3)


# ------------------------------------------------------------------ K6 rope
def test_rope_fwd_bwd(dev):
    from vantage6_b200.ops import rope as R

    torch.manual_seed(4)
    B, S, Hq, Hkv, D = 2, 64, 8, 2, 128
    cos, sin = R.rope_tables(S, D, device=dev)
    q = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
    qr, kr = R.reference_rope(q, k, cos, sin)
    q1, k1 = q.clone().requires_grad_(), k.clone().requires_grad_()
    qo, ko = R.apply_rope(q1.clone(), k1.clone(), cos, sin)
    torch.testing.assert_close(qo.float(), qr.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(ko.float(), kr.float(), rtol=2e-2, atol=2e-2)
    # inverse rotation restores the input (backward = inverse)
    from vantage6_b200.ops import native, stream_ptr

    q2, k2 = qo.detach().clone(), ko.detach().clone()
    native().rope(q2.data_ptr(), k2.data_ptr(), cos.data_ptr(), sin.data_ptr(), 0, B, S, Hq, Hkv, D, True, stream_ptr())
    torch.testing.assert_close(q2.float(), q.float(), rtol=3e-2, atol=3e-2)


# ------------------------------------------------------------------ K8 glm
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_glm_logistic_grad(dev, dtype):
    from vantage6_b200.ops import glm as G

    torch.manual_seed(5)
    rows, F = 20000, 256
    X = torch.randn(rows, F, device=dev).to(dtype)
    w = torch.randn(F + 1, device=dev) * 0.1
    y = (torch.rand(rows, device=dev) < 0.4).float()
    out = G.logistic_grad(X, y, w)
    ref = G.reference_logistic_grad(X, y, w)
    torch.testing.assert_close(out[: F + 3], ref, rtol=2e-3, atol=5e-2)


# ------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (4096, 2304, 768), (1000, 3072, 768),
                                   (77, 264, 136), (8192, 4096, 4096)])
def test_tcgen05_gemm(dev, M, N, K):
    from vantage6_b200.ops import gemm as G

    torch.manual_seed(6)
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev)
    c = G.gemm_bf16(a, w, bias, G.ACT_NONE)
    ref = a.float() @ w.float().t() + bias
    err = (c.float() - ref).abs().max() / ref.abs().max()
    assert err < 1e-2, f"rel err {err}"
    c2 = G.gemm_bf16(a, w, bias, G.ACT_GELU)
    ref2 = torch.nn.functional.gelu(ref)
    err2 = (c2.float() - ref2).abs().max() / ref2.abs().max()
    assert err2 < 1e-2, f"gelu rel err {err2}"


def test_linear_autograd(dev):
    from vantage6_b200.ops import gemm as G

    torch.manual_seed(7)
    x = torch.randn(4, 128, 768, device=dev, dtype=torch.bfloat16, requires_grad=True)
    W = (torch.randn(3072, 768, device=dev, dtype=torch.bfloat16) * 0.02).requires_grad_()
    b = torch.zeros(3072, device=dev, requires_grad=True)
    y = G.linear(x, W, b, G.ACT_GELU)
    y.float().pow(2).mean().backward()
    xr = x.detach().float().requires_grad_()
    Wr = W.detach().float().requires_grad_()
    br = b.detach().clone().requires_grad_()
    yr = torch.nn.functional.gelu(xr @ Wr.t() + br)
    yr.pow(2).mean().backward()
    torch.testing.assert_close(y.float(), yr, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(x.grad.float(), xr.grad, rtol=5e-2, atol=1e-4)
    torch.testing.assert_close(W.grad.float(), Wr.grad, rtol=5e-2, atol=1e-4)


# ------------------------------------------------------------------ K2 / K3 single-rank + trainer
@pytest.mark.parametrize("opt", ["fedavg", "fedavgm", "fedadam"])
@pytest.mark.parametrize("upload", ["weights_f32", "delta_f32", "delta_bf16"])
def test_fedavg_engine_world1_matches_collective(dev, opt, upload):
    from vantage6_b200.parallel.fedavg import FedAvgEngine, ServerOptConfig

    torch.manual_seed(8)
    n = 100_003
    cfg = ServerOptConfig(opt, 0.7)
    e1 = FedAvgEngine(n, 0, 1, dev, data_plane="native", server_opt=cfg, upload=upload)
    e2 = FedAvgEngine(n, 0, 1, dev, data_plane="collective", server_opt=ServerOptConfig(opt, 0.7), upload=upload)
    w0 = torch.randn(e1.n, device=dev)
    for e in (e1, e2):
        e.w.copy_(w0)
        e.initialize_global()
    for rnd in range(3):
        step = torch.randn(e1.n, device=dev) * 0.1
        for e in (e1, e2):
            if upload == "weights_f32":
                e.w.add_(step)
            else:
                e.upload.copy_(step.to(e.upload.dtype))
            e.aggregate(5.0)
        torch.cuda.synchronize()
        tol = dict(rtol=2e-2, atol=2e-3) if upload == "delta_bf16" else dict(rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(e1.w, e2.w, **tol)
    assert e1.poll_status() == 0
    e1.close()


def test_small_allreduce_world1(dev):
    from vantage6_b200.parallel.fedavg import SmallAggregator

    agg = SmallAggregator(1000, 0, 1, dev)
    v = torch.randn(1000, device=dev)
    agg.slot()[:1000].copy_(v)
    out = agg.allreduce(3.0)
    torch.cuda.synchronize()
    torch.testing.assert_close(out[:1000], v)
    agg.close()


def test_trainer_graph_matches_eager(dev):
    from vantage6_b200.models.resnet import imagenet_forward_loss, resnet_tiny
    from vantage6_b200.parallel.trainer import FederatedTrainer

    def make(graph):
        torch.manual_seed(9)
        return FederatedTrainer(resnet_tiny(10).to(memory_format=torch.channels_last), imagenet_forward_loss, rank=0,
                                world=1, device=dev, lr=0.05, use_cuda_graph=graph, amp_dtype=None)

    a, b = make(True), make(False)
    x = torch.randint(0, 256, (3, 8, 3, 64, 64), dtype=torch.uint8, device=dev)
    y = torch.randint(0, 10, (3, 8), device=dev)
    batches = [(x[i], y[i]) for i in range(3)]
    for t in (a, b):
        t.initialize_global()
    la = [a.run_round(batches, 24.0).item() for _ in range(2)]
    lb = [b.run_round(batches, 24.0).item() for _ in range(2)]
    assert abs(la[0] - lb[0]) < 1e-3 and abs(la[1] - lb[1]) < 5e-2, (la, lb)
    torch.testing.assert_close(a.engine.w, b.engine.w, rtol=1e-2, atol=1e-3)
    a.close()
    b.close()


@pytest.mark.parametrize("T,V", [(64, 1000), (33, 128256), (128, 512), (7, 30522)])
def test_fused_cross_entropy(dev, T, V):
    """csrc/ce.cu: mean cross-entropy over bf16 logits (online log-sum-exp, ignore_index) and its in-place gradient vs PyTorch fp32."""
    from vantage6_b200.ops.ce import fused_cross_entropy

    torch.manual_seed(12)
    Vp = (V + 7) // 8 * 8
    buf = (torch.randn(T, Vp, device=dev) * 3).to(torch.bfloat16)
    logits = buf[:, :V].requires_grad_() if Vp == V else buf[:, :V].detach().requires_grad_()
    labels = torch.randint(0, V, (T,), device=dev)
    labels[::5] = -100
    ref_in = logits.detach().float().requires_grad_()
    ref = torch.nn.functional.cross_entropy(ref_in, labels, ignore_index=-100)
    ref.backward()
    x = logits.detach().clone().requires_grad_() if Vp == V else logits
    loss = fused_cross_entropy(x, labels)
    (loss * 1.0).backward()
    torch.cuda.synchronize()
    torch.testing.assert_close(loss.float(), ref, rtol=2e-3, atol=2e-3)
    g = x.grad.float()
    torch.testing.assert_close(g, ref_in.grad, rtol=5e-2, atol=2e-2 * float(ref_in.grad.abs().max()))
