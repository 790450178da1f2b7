"""Host-side logic of the FedAvg engine / trainer on CPU: closed-form checks of the aggregation
math for every server optimizer / upload mode / server placement, partial participation, dead
nodes, checkpoint round-trip, multi-process gloo (world_size 2) and the reference
implementations that serve as numerics oracles for the CUDA kernels."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vantage6_b200.models import zoo
from vantage6_b200.models.flat import FlatModel, flat_size
from vantage6_b200.models.resnet import imagenet_forward_loss, resnet_tiny
from vantage6_b200.ops import glm as K8
from vantage6_b200.ops import norm as N
from vantage6_b200.ops import optim as O
from vantage6_b200.ops import rope as R
from vantage6_b200.parallel.fedavg import FedAvgEngine, ServerOptConfig, SmallAggregator
from vantage6_b200.parallel.trainer import FederatedTrainer
from vantage6_b200.utils.checkpoint import load_checkpoint, save_checkpoint


def test_flat_model_aliases_parameters_and_buffers():
    m = resnet_tiny(10)
    n = flat_size(m)
    fm = FlatModel(m)
    assert fm.n_total == n and fm.check_aliasing()
    before = m.fc.weight.clone()
    fm.flat.mul_(2.0)
    assert torch.equal(m.fc.weight, before * 2)                       # views, not copies
    assert m.bn1.running_var.data_ptr() >= fm.flat.data_ptr()         # BN stats are federated too
    x = torch.randn(2, 3, 32, 32)
    m(x).sum().backward()
    assert fm.grad.abs().sum() > 0 and m.fc.weight.grad.data_ptr() >= fm.grad.data_ptr()


@pytest.mark.parametrize("opt", ["fedavg", "fedavgm", "fedadam"])
def test_world1_server_optimizers_closed_form(opt):
    n = 64
    e = FedAvgEngine(n, data_plane="collective", server_opt=ServerOptConfig(opt, 0.5, 0.9, 0.99, 1e-3))
    w0 = torch.randn(e.n)
    e.w.copy_(w0)
    e.initialize_global()
    step = torch.randn(e.n)
    e.w.add_(step)
    e.aggregate(10.0)
    d = step
    if opt == "fedavg":
        exp = w0 + 0.5 * d
    elif opt == "fedavgm":
        exp = w0 + 0.5 * d                                            # m = d on the first step
    else:
        m, v = 0.1 * d, 0.01 * d * d
        exp = w0 + 0.5 * (m / 0.1) / ((v / 0.01).sqrt() + 1e-3)
    torch.testing.assert_close(e.w, exp, rtol=1e-5, atol=1e-6)


def test_trainer_cpu_all_upload_modes_agree():
    outs = []
    for upload in ("weights_f32", "delta_f32"):
        torch.manual_seed(0)
        tr = FederatedTrainer(resnet_tiny(10), imagenet_forward_loss, device="cpu", lr=0.01, upload=upload)
        tr.initialize_global()
        x = torch.randint(0, 256, (2, 4, 3, 32, 32), dtype=torch.uint8)
        y = torch.randint(0, 10, (2, 4))
        for _ in range(2):
            tr.run_round([(x[0], y[0]), (x[1], y[1])])
        outs.append(tr.engine.w.clone())
    torch.testing.assert_close(outs[0][: outs[1].numel()], outs[1], rtol=1e-5, atol=1e-6)


def test_checkpoint_roundtrip(tmp_path):
    torch.manual_seed(0)
    tr = FederatedTrainer(resnet_tiny(10), imagenet_forward_loss, device="cpu", lr=0.01,
                          server_opt=ServerOptConfig("fedadam", 0.1))
    tr.initialize_global()
    x = torch.randint(0, 256, (1, 4, 3, 32, 32), dtype=torch.uint8)
    y = torch.randint(0, 10, (1, 4))
    tr.run_round([(x[0], y[0])])
    path = save_checkpoint(tmp_path / "ckpt", tr, round_idx=1)
    ref_next = None
    w_after_1 = tr.engine.w.clone()
    tr.run_round([(x[0], y[0])])
    ref_next = tr.engine.w.clone()
    torch.manual_seed(1)
    tr2 = FederatedTrainer(resnet_tiny(10), imagenet_forward_loss, device="cpu", lr=0.01,
                           server_opt=ServerOptConfig("fedadam", 0.1))
    meta = load_checkpoint(path, tr2)
    assert meta["round"] == 1
    torch.testing.assert_close(tr2.engine.w, w_after_1)
    tr2.run_round([(x[0], y[0])])
    torch.testing.assert_close(tr2.engine.w, ref_next, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ multi-process (gloo)
def _worker(rank, world, port, q, server_mode, upload, opt):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1003
        e = FedAvgEngine(n, rank, world, "cpu", server_mode=server_mode, upload=upload, server_opt=ServerOptConfig(opt, 1.0))
        g = torch.Generator().manual_seed(7)
        w0 = torch.randn(e.n, generator=g)
        e.w.copy_(w0 if rank == 0 else torch.zeros(e.n))
        e.initialize_global()
        assert torch.allclose(e.w, w0)
        steps = [torch.randn(e.n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        weights = [1.0, 3.0]
        if upload == "weights_f32":
            e.w.add_(steps[rank])
        else:
            e.upload.copy_(steps[rank])
        e.aggregate(weights)
        exp = w0 + sum(w * s for w, s in zip(weights, steps)) / sum(weights)
        ok1 = torch.allclose(e.w, exp, rtol=1e-5, atol=1e-6)
        # partial participation: node 1 does not report -> renormalise over the reporters
        w1 = e.w.clone()
        if upload == "weights_f32":
            e.w.add_(steps[rank])
        else:
            e.upload.copy_(steps[rank])
        e.aggregate([2.0, 0.0])
        ok2 = torch.allclose(e.w, w1 + steps[0], rtol=1e-5, atol=1e-6)
        # K3 path
        agg = SmallAggregator(10, rank, world, "cpu")
        agg.slot()[:10] = float(rank + 1)
        out = agg.allreduce([1.0, 3.0])
        ok3 = torch.allclose(out[:10], torch.full((10,), (1 * 1 + 3 * 2) / 4.0))
        q.put((rank, bool(ok1), bool(ok2), bool(ok3)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("server_mode,upload,opt", [("sharded", "weights_f32", "fedavg"), ("central", "delta_f32", "fedavg"),
                                                   ("sharded", "delta_f32", "fedavg")])
def test_two_process_gloo_aggregation(server_mode, upload, opt):
    from vantage6_b200.dev import free_port

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, server_mode, upload, opt)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(30)
    assert res == [(0, True, True, True), (1, True, True, True)], res


def _unequal_worker(rank, world, port, q, server_mode, upload):
    """Every rank passes only ITS OWN sample count (what a real node knows): the engine must still apply the
    sample-weighted mean (a scalar used to be expanded to [n_r] * world on every rank)."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torch import nn

        from vantage6_b200.parallel.trainer import FederatedTrainer

        torch.manual_seed(0)
        model = nn.Linear(4, 1, bias=False)
        tr = FederatedTrainer(model, lambda m, x, y: ((m(x) - y) ** 2).mean(), rank=rank, world=world, device="cpu",
                              optimizer="sgd", lr=0.0, momentum=0.0, server_mode=server_mode, upload=upload, amp_dtype=None)
        tr.initialize_global()
        with torch.no_grad():                                   # local "training": node r moves every weight to a known value
            target = 1.0 if rank == 0 else 3.0
        x, y = torch.zeros(2, 4), torch.zeros(2, 1)
        n_i = 100.0 if rank == 0 else 300.0
        # lr = 0: the local step leaves the weights alone; set them by hand between the step and the aggregation
        orig = tr.engine.aggregate

        def agg(*a, **k):
            with torch.no_grad():
                if upload == "weights_f32":
                    tr.fm.params.fill_(target)
                else:
                    tr.engine.upload[: tr.fm.n_trainable].copy_(((target - tr.w_ref[: tr.fm.n_trainable]) * n_i).to(tr.engine.upload.dtype))
            return orig(*a, **k)
        tr.engine.aggregate = agg
        tr.run_round([(x, y)], n_samples=n_i)
        q.put((rank, [round(float(v), 4) for v in tr.fm.params[:4]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("server_mode,upload", [("sharded", "weights_f32"), ("central", "weights_f32"), ("sharded", "delta_f32")])
def test_unequal_sample_counts_give_the_sample_weighted_mean(server_mode, upload):
    from vantage6_b200.dev import free_port

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_unequal_worker, args=(r, 2, port, q, server_mode, upload)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(30)
    # w=1 with n=100 and w=3 with n=300 -> (100*1 + 300*3) / 400 = 2.5 everywhere (plain mean would be 2.0)
    assert res == [(0, [2.5] * 4), (1, [2.5] * 4)], res


def test_dead_node_is_excluded():
    e = FedAvgEngine(16, data_plane="collective")
    e.world = 3                                   # host-side logic only
    e.live_mask = 0b111
    e.mark_dead(1)
    assert e.live_mask == 0b101
    w = [1.0, 2.0, 3.0]
    assert [x if (e.live_mask >> r) & 1 else 0.0 for r, x in enumerate(w)] == [1.0, 0.0, 3.0]


# ------------------------------------------------------------------ reference oracles
def test_reference_sgd_matches_torch_optim():
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(100))
    w, buf = p.detach().clone(), torch.zeros(100)
    topt = torch.optim.SGD([p], lr=0.1, momentum=0.9, weight_decay=1e-3, nesterov=True)
    for i in range(3):
        g = torch.randn(100)
        p.grad = g.clone()
        topt.step()
        O.reference_sgd_step(w, g, buf, 0.1, 0.9, 0.0, 1e-3, True, i == 0)
    torch.testing.assert_close(w, p.detach(), rtol=1e-6, atol=1e-6)


def test_flat_optimizers_cpu_publish_modes():
    w = torch.randn(32)
    w0 = w.clone()
    ref, up = torch.zeros(32), torch.zeros(32, dtype=torch.bfloat16)
    sh = torch.zeros(32, dtype=torch.bfloat16)
    opt = O.FlatAdamW(w, lr=1e-2)
    opt.step(torch.randn(32), w_ref=ref, save_ref=True, upload=up, publish=O.PUBLISH_DELTA_BF16, contrib_scale=2.0, shadow=sh)
    torch.testing.assert_close(ref, w0)
    torch.testing.assert_close(up.float(), (2.0 * (w - w0)).to(torch.bfloat16).float())
    torch.testing.assert_close(sh.float(), w.to(torch.bfloat16).float())
    c = O.clip_grad_coef(torch.full((100,), 3.0), 1.0)
    assert abs(c.item() - 1.0 / 30.0) < 1e-4


def test_reference_norms_rope_glm_linear():
    x = torch.randn(5, 64)
    g, b = torch.rand(64) + 0.5, torch.randn(64)
    y, _ = N.layer_norm(x, g, b, 1e-5)
    torch.testing.assert_close(y, torch.nn.functional.layer_norm(x, (64,), g, b, 1e-5))
    r = torch.randn(5, 64)
    y2, h = N.rms_norm(x, g, 1e-5, residual=r)
    hh = x + r
    torch.testing.assert_close(h, hh)
    torch.testing.assert_close(y2, hh * torch.rsqrt(hh.pow(2).mean(-1, keepdim=True) + 1e-5) * g)
    cos, sin = R.rope_tables(8, 16)
    q, k = torch.randn(1, 8, 2, 16), torch.randn(1, 8, 1, 16)
    q2, k2 = R.apply_rope(q, k, cos, sin)
    torch.testing.assert_close(q2.norm(dim=-1), q.norm(dim=-1))        # rotations preserve norms
    torch.testing.assert_close(q2[:, 0], q[:, 0])                       # position 0: identity
    X, yv, w = torch.randn(50, 256), (torch.rand(50) < 0.5).float(), torch.randn(257) * 0.1
    out = K8.logistic_grad(X, yv, w)
    wp = w.clone().requires_grad_()
    loss = torch.nn.functional.binary_cross_entropy_with_logits(X @ wp[:256] + wp[256], yv, reduction="sum")
    loss.backward()
    torch.testing.assert_close(out[:257], wp.grad, rtol=1e-4, atol=1e-4)
    assert abs(out[257].item() - loss.item()) < 1e-3 and out[258].item() == 50


@pytest.mark.parametrize("name", ["bert_tiny", "llama_tiny_lora", "resnet_tiny"])
def test_zoo_models_train_on_cpu(name):
    torch.manual_seed(0)
    tr, spec = zoo.build_trainer(name, rank=0, world=1, device="cpu")
    tr.initialize_global()
    batches = spec.make_batches(spec.local_steps, spec.batch, seed=3)
    losses = [tr.run_round(batches).item() for _ in range(3)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    if name == "llama_tiny_lora":           # only the adapters are federated
        names = [s.name for s in tr.fm.segments]
        assert names and all("lora_" in n for n in names)


def test_flat_model_keeps_channels_last_filters_and_gradient_sink():
    """Conv filters live [O,H,W,I] inside the flat buffers (no per-step layout conversion); the gradient sink adds
    bf16 filter gradients into the flat fp32 gradient buffer in memory order."""
    import torch
    from torch import nn

    from vantage6_b200.models.flat import FlatModel
    from vantage6_b200.ops.optim import multi_accumulate

    torch.manual_seed(0)
    m = nn.Sequential(nn.Conv2d(3, 8, 3, bias=False), nn.Conv2d(8, 4, 1, bias=False)).to(memory_format=torch.channels_last)
    ref = [p.detach().clone() for p in m.parameters()]
    fm = FlatModel(m)
    seg3, seg1 = fm.segment("0.weight"), fm.segment("1.weight")
    assert seg3.channels_last and not seg1.channels_last           # 1x1 filters: both layouts coincide
    for p, r in zip(m.parameters(), ref):
        assert torch.equal(p, r)
    w = m[0].weight
    assert w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous()
    flat = fm.flat[seg3.offset: seg3.offset + seg3.numel]
    assert torch.equal(flat.view(8, 3, 3, 3), ref[0].permute(0, 2, 3, 1))     # stored O,H,W,I
    assert m[0].weight.grad.is_contiguous(memory_format=torch.channels_last)
    # gradient sink (CPU fallback of the multi-tensor kernel)
    g3 = torch.randn(8, 3, 3, 3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    g1 = torch.randn(4, 8, 1, 1).to(torch.bfloat16)
    fm.grad_sink.extend([(g3, seg3.offset), (g1, seg1.offset)])
    assert fm.flush_grad_sink() == 2 and not fm.grad_sink
    torch.testing.assert_close(m[0].weight.grad, g3.float())
    torch.testing.assert_close(m[1].weight.grad, g1.float())
    fm.grad_sink.append((g3, seg3.offset))
    fm.flush_grad_sink()
    torch.testing.assert_close(m[0].weight.grad, 2 * g3.float())
    dst = torch.zeros(64)
    multi_accumulate(dst, [(torch.ones(5, dtype=torch.bfloat16), 8)])
    assert dst[8:13].sum() == 5 and dst.sum() == 5


def test_shadow_conv_falls_back_to_plain_conv_without_shadow():
    import torch

    from vantage6_b200.models.conv import ShadowConv2d

    torch.manual_seed(1)
    c = ShadowConv2d(4, 6, 3, padding=1, bias=False)
    x = torch.randn(2, 4, 5, 5, requires_grad=True)
    y = c(x)
    ref = torch.nn.functional.conv2d(x, c.weight, None, 1, 1)
    torch.testing.assert_close(y, ref)
    y.sum().backward()
    assert c.weight.grad is not None and x.grad is not None


def test_resnet_stem_paths_agree_on_cpu():
    """uint8 input (normalised inside the model) == pre-normalised float input."""
    import torch

    from vantage6_b200.models.resnet import _MEAN, _STD, resnet_tiny

    torch.manual_seed(2)
    m = resnet_tiny(10).eval()
    img = torch.randint(0, 256, (2, 3, 32, 32), dtype=torch.uint8)
    mean = torch.tensor(_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(_STD).view(1, 3, 1, 1)
    with torch.no_grad():
        a = m(img)
        b = m((img.float() - mean) / std)
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)


def test_stock_optimizer_arm_matches_fused_arm_in_delta_mode():
    """The baseline arm (torch.optim, no fused publish) must federate the same model as the fused arm: the round
    reference is saved before the first local step and the delta published after the last one."""
    import torch

    from vantage6_b200.models import zoo

    def run(fused):
        torch.manual_seed(21)
        tr, spec = zoo.build_trainer("bert_tiny", rank=0, world=1, device="cpu", data_plane="auto" if fused else "collective",
                                     fused_local_optimizer=fused)
        batches = spec.make_batches(2, 4, 77)
        tr.initialize_global()
        losses = [float(tr.run_round(batches)) for _ in range(4)]
        tr.close()
        return losses

    a, b = run(True), run(False)
    assert a[-1] < a[0] - 0.3                     # the (fixed, synthetic) shard is being fitted
    for x, y in zip(a, b):
        assert abs(x - y) < 5e-3, (a, b)


def test_lora_linear_single_node_matches_composed_expression():
    """_LoRALinearFn (base GEMM + x A^T + fused addmm, hand-written backward) == x W^T + s (x A^T) B^T."""
    import torch

    from vantage6_b200.models.transformer import LoRALinear

    torch.manual_seed(0)
    m = LoRALinear(32, 48, r=4, alpha=8.0)
    with torch.no_grad():
        m.lora_B.normal_(0, 0.1)
    x = torch.randn(5, 7, 32, requires_grad=True)
    y = m(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    W = m.base.weight_bf16.float()
    A = m.lora_A.detach().clone().requires_grad_()
    B = m.lora_B.detach().clone().requires_grad_()
    xr = x.detach().clone().requires_grad_()
    (xr @ W.t() + m.scaling * (xr @ A.t()) @ B.t()).backward(dy)
    torch.testing.assert_close(x.grad, xr.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(m.lora_A.grad, A.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(m.lora_B.grad, B.grad, rtol=1e-5, atol=1e-5)


def test_swiglu_cpu_fallback_and_frozen_norm_weights():
    import torch

    from vantage6_b200.ops.act import swiglu
    from vantage6_b200.ops.norm import rms_norm

    g, u = torch.randn(4, 16), torch.randn(4, 16)
    torch.testing.assert_close(swiglu(g, u), torch.nn.functional.silu(g) * u)
    w = torch.nn.Parameter(torch.ones(16), requires_grad=False)
    x = torch.randn(4, 16, requires_grad=True)
    y, _ = rms_norm(x, w, 1e-5)
    y.sum().backward()
    assert x.grad is not None and w.grad is None


def test_bias_act_backward_reference_path_and_token_stride_detection():
    import torch

    from vantage6_b200.ops import attention as A
    from vantage6_b200.ops import gemm as G

    torch.manual_seed(4)
    dy, pre = torch.randn(6, 64), torch.randn(6, 64)
    bias = torch.nn.Parameter(torch.zeros(64))
    for act in (G.ACT_NONE, G.ACT_GELU, G.ACT_RELU):
        p = pre.clone().requires_grad_()
        y = {G.ACT_NONE: lambda t: t, G.ACT_GELU: torch.nn.functional.gelu, G.ACT_RELU: torch.relu}[act](p)
        y.backward(dy)
        dpre, db = G.bias_act_backward(dy, pre, act, bias)
        torch.testing.assert_close(dpre, p.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(db, p.grad.sum(0), rtol=1e-4, atol=1e-5)
    # packed QKV slices are token-strided views that the attention kernels can read in place
    qkv = torch.zeros(2, 16, 3, 4, 8, dtype=torch.bfloat16)
    q, ld = A._token_strided(qkv[:, :, 0])
    assert ld == 3 * 4 * 8 and q.data_ptr() == qkv.data_ptr()
    k, ld = A._token_strided(qkv[:, :, 1].transpose(1, 2).transpose(1, 2))
    assert ld == 3 * 4 * 8
    odd, ld = A._token_strided(torch.zeros(2, 16, 4, 9, dtype=torch.bfloat16)[..., :8])     # head dim not dense
    assert ld == 4 * 8 and odd.is_contiguous()


def test_packed_attention_falls_back_off_gpu():
    import torch

    from vantage6_b200.models.transformer import attention, packed_attention

    torch.manual_seed(5)
    qkv = torch.randn(2, 8, 3, 2, 16)
    ref = attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal=False)
    torch.testing.assert_close(packed_attention(qkv, causal=False), ref)


def test_flat_model_puts_first_consumer_weights_in_a_prefix():
    """``_v6_first`` parameters (the weights K1 delivers: trainer bcast="fused") open the flat buffers as one contiguous range."""
    from vantage6_b200.models.flat import FlatModel

    m = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 24), torch.nn.Linear(24, 8))
    m[1].weight._v6_first = True
    m[2].weight._v6_first = True
    ref = {n: p.detach().clone() for n, p in m.named_parameters()}
    fm = FlatModel(m)
    assert fm.segment("1.weight").offset == 0 and fm.segment("2.weight").offset == 16 * 24
    assert fm.n_first == 16 * 24 + 24 * 8
    assert fm.segment("0.weight").offset == fm.n_first
    for n, p in m.named_parameters():                  # re-homing keeps the values, whatever the order
        torch.testing.assert_close(p.detach(), ref[n])
    assert fm.check_aliasing()


def test_engine_shard_alignment():
    from vantage6_b200.parallel.fedavg import FedAvgEngine

    e = FedAvgEngine(10_000, 0, 1, "cpu", shard_align=256)
    e.reducers = [0, 1, 2]
    e.rank = 1
    e._reshard()
    assert e.lo % 256 == 0 and e.hi % 256 == 0 and e.hi > e.lo
    assert e.k1_layer(0, 256, 8) is None               # no NVLink data plane on the CPU: K2 keeps pushing everything


def test_trainer_bcast_fused_is_a_noop_without_the_nvlink_plane():
    """bcast="fused" (K1 on the engine path) needs the NVLink data plane + multicast: on the CPU / collective plane the trainer
    runs the plain path (no K1 layers, no skipped shadow range) and trains as usual."""
    from vantage6_b200.models import zoo

    tr, spec = zoo.build_trainer("bert_tiny", rank=0, world=1, device="cpu", bcast="fused")
    assert tr.k1_layers == 0 and tr.engine.shadow_skip == (0, 0)
    tr.initialize_global()
    batches = spec.make_batches(2, 4, 3)
    l0 = float(tr.run_round(batches).item())
    l1 = float(tr.run_round(batches).item())
    assert l1 < l0
    tr.close()


def test_slice_and_k1_block_ownership_properties():
    """Pure layout math of the engine (hypothesis): reducer slices tile the flat buffer exactly on aligned boundaries, and
    the 256-row blocks of a K1 layer are owned by exactly one reducer each -- or the layer is refused on every rank alike."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from vantage6_b200.parallel.fedavg import k1_own_blocks, slice_bounds

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 5_000_000), st.integers(1, 8), st.sampled_from([8, 64, 256, 256 * 64, 256 * 768]))
    def slices(n, nr, align):
        bounds = [slice_bounds(n, nr, pos, align) for pos in range(nr)]
        chunk = bounds[0][2]
        assert all(b[2] == chunk for b in bounds) and chunk % align == 0 and chunk * nr >= n
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        for (lo, hi, _), (lo2, _, _) in zip(bounds, bounds[1:]):
            assert lo <= hi == lo2 and (lo % align == 0 or lo == n)         # empty tail slices sit at n
        assert slice_bounds(n, nr, None, align)[:2] == (0, 0)

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 8), st.integers(1, 12), st.sampled_from([64, 768, 1024]), st.integers(0, 40), st.integers(0, 5_000_000),
           st.booleans())
    def blocks(nr, n_blocks, k_in, prefix_blocks, tail, aligned):
        n_out, blk = 256 * n_blocks, 256 * k_in
        offset = prefix_blocks * blk if aligned else prefix_blocks * blk + 8 * 3            # a layer that starts off the block grid
        n = offset + n_out * k_in + tail
        align = blk if aligned else 8
        bounds = [slice_bounds(n, nr, pos, align) for pos in range(nr)]
        owns = [k1_own_blocks(lo, hi, chunk, nr, offset, n_out, k_in) for lo, hi, chunk in bounds]
        assert all(o is None for o in owns) or all(o is not None for o in owns)            # collective decision
        if owns[0] is None:
            return
        covered = []
        for a, z in owns:
            assert 0 <= a <= z <= n_blocks
            covered += list(range(a, z))
        assert covered == list(range(n_blocks))                                            # every block exactly once, in order
        for (lo, hi, _), (a, z) in zip(bounds, owns):                                      # and an owner really holds its blocks
            assert z == a or (lo <= offset + a * blk and offset + z * blk <= hi)

    slices()
    blocks()
    # block-aligned shards (what the trainer asks for with bcast="fused") are never refused
    for nr in range(1, 9):
        bounds = [slice_bounds(10_000_000, nr, pos, 256 * 768) for pos in range(nr)]
        assert all(k1_own_blocks(lo, hi, c, nr, 3 * 256 * 768, 2304, 768) is not None for lo, hi, c in bounds)
