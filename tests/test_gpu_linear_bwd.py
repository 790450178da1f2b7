"""Fused bias / activation backward (csrc/act.cu), LayerNorm parameter-gradient fold + in-place accumulation, and the
ShadowLinear gradient-sink path against plain PyTorch references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from vantage6_b200.ops import native

    native()
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


@pytest.mark.parametrize("R,C", [(4096, 3072), (4096, 768), (1000, 2304), (37, 64), (20000, 128)])
@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("direct", [False, True])
def test_bias_act_backward(dev, R, C, act, direct):
    from vantage6_b200.ops import gemm as G

    torch.manual_seed(0)
    dy = torch.randn(R, C, device=dev).to(torch.bfloat16)
    pre = (torch.randn(R, C, device=dev) * 1.5).to(torch.bfloat16)
    bias = torch.nn.Parameter(torch.zeros(C, device=dev))
    if direct:
        bias.grad = torch.full((C,), 0.5, device=dev)
    for rep in range(2):                                   # second call: arrival counters must have reset
        dpre, db = G.bias_act_backward(dy, pre if act else None, act, bias)
    p = pre.float()
    if act == 1:
        g = dy.float() * (0.5 * (1 + torch.erf(p * 0.7071067811865476)) + p * torch.exp(-0.5 * p * p) * 0.3989422804014327)
    elif act == 2:
        g = dy.float() * (p > 0)
    else:
        g = dy.float()
    torch.testing.assert_close(dpre.float(), g.to(torch.bfloat16).float(), rtol=2e-2, atol=2e-2)
    ref_db = g.sum(0)
    tol = 2e-2 * max(1.0, ref_db.abs().max().item())
    if direct:
        assert db is None
        assert (bias.grad - (0.5 + 2 * ref_db)).abs().max().item() < 2 * tol
    else:
        assert (db - ref_db).abs().max().item() < tol


def test_layernorm_accumulates_param_grads_in_place(dev):
    from vantage6_b200.ops import norm as N

    torch.manual_seed(1)
    rows, cols = 4096, 768
    x = torch.randn(rows, cols, device=dev).to(torch.bfloat16).requires_grad_()
    g = torch.nn.Parameter(torch.rand(cols, device=dev) + 0.5)
    b = torch.nn.Parameter(torch.randn(cols, device=dev) * 0.1)
    dy = torch.randn(rows, cols, device=dev).to(torch.bfloat16)
    y, _ = N.layer_norm(x, g, b, 1e-5)
    y.backward(dy)
    ref_g, ref_b = g.grad.clone(), b.grad.clone()                 # first backward: grads returned to autograd
    xf = x.detach().float().requires_grad_()
    gr = g.detach().clone().requires_grad_()
    br = b.detach().clone().requires_grad_()
    torch.nn.functional.layer_norm(xf, (cols,), gr, br, 1e-5).backward(dy.float())
    assert (ref_g - gr.grad).abs().max().item() < 3e-2 * gr.grad.abs().max().item()
    assert (ref_b - br.grad).abs().max().item() < 3e-2 * br.grad.abs().max().item()
    y2, _ = N.layer_norm(x, g, b, 1e-5)                           # .grad exists now: accumulated in place by the kernel
    y2.backward(dy)
    torch.testing.assert_close(g.grad, 2 * ref_g, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(b.grad, 2 * ref_b, rtol=1e-4, atol=1e-3)


def test_bert_tiny_fused_arm_matches_stock_arm(dev):
    """ShadowLinear with gradient sink + fused bias/GELU backward + in-place LN grads vs the stock-optimizer arm."""
    from vantage6_b200.models import zoo

    def run(fused):
        torch.manual_seed(21)
        tr, spec = zoo.build_trainer("bert_tiny", rank=0, world=1, device=dev, data_plane="auto" if fused else "collective",
                                     fused_local_optimizer=fused, use_cuda_graph=fused)
        batches = [(x.to(dev), y.to(dev)) for x, y in spec.make_batches(2, 4, 77)]
        tr.initialize_global()
        losses = [float(tr.run_round(batches).item()) for _ in range(5)]
        tr.close()
        return losses

    a, b = run(True), run(False)
    assert abs(a[0] - b[0]) < 0.05, (a, b)
    assert a[-1] < a[0] and abs(a[-1] - b[-1]) < 0.3, (a, b)


@pytest.mark.parametrize("shape", [(1024, 14336), (37, 64), (8, 8)])
def test_swiglu_kernels(dev, shape):
    from vantage6_b200.ops.act import swiglu

    torch.manual_seed(3)
    g = (torch.randn(*shape, device=dev) * 2).to(torch.bfloat16).requires_grad_()
    u = torch.randn(*shape, device=dev).to(torch.bfloat16).requires_grad_()
    h = swiglu(g, u)
    dh = torch.randn_like(h)
    h.backward(dh)
    gf, uf = g.detach().float().requires_grad_(), u.detach().float().requires_grad_()
    hr = torch.nn.functional.silu(gf) * uf
    hr.backward(dh.float())
    torch.testing.assert_close(h.float(), hr, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(g.grad.float(), gf.grad, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(u.grad.float(), uf.grad, rtol=3e-2, atol=3e-2)


def test_llama_tiny_fused_arm_matches_stock_arm(dev):
    """LoRA adapters through _LoRALinearFn (bf16 shadows + gradient sink), fused SwiGLU, frozen RMSNorm weights."""
    from vantage6_b200.models import zoo

    def run(fused):
        torch.manual_seed(31)
        tr, spec = zoo.build_trainer("llama_tiny_lora", rank=0, world=1, device=dev,
                                     data_plane="auto" if fused else "collective", fused_local_optimizer=fused,
                                     use_cuda_graph=fused)
        batches = [(x.to(dev), y.to(dev)) for x, y in spec.make_batches(2, 2, 78)]
        tr.initialize_global()
        losses = [float(tr.run_round(batches).item()) for _ in range(5)]
        tr.close()
        return losses

    a, b = run(True), run(False)
    assert abs(a[0] - b[0]) < 0.05, (a, b)
    assert a[-1] < a[0] and abs(a[-1] - b[-1]) < 0.3, (a, b)


def test_frozen_linear_dx_through_transposed_copy(dev, monkeypatch):
    """FrozenLinear: dX = dY . W on the K-major tcgen05 GEMM with a transposed copy of the frozen weight made once
    (models/transformer.py::frozen_transposed) vs the fp32 product."""
    from vantage6_b200.models import transformer as T

    monkeypatch.setenv("V6B200_FROZEN_DX", "gemm")
    torch.manual_seed(5)
    lin = T.FrozenLinear(256, 384, device=dev, init_std=0.05)
    x = (torch.randn(3, 50, 256, device=dev) * 0.5).to(torch.bfloat16).requires_grad_()
    y = lin(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    assert T.frozen_transposed(lin) is not None and T.frozen_transposed(lin).shape == (256, 384)
    ref = dy.float().reshape(-1, 384) @ lin.weight_bf16.float()
    torch.testing.assert_close(x.grad.float().reshape(-1, 256), ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
    with torch.no_grad():
        lin.weight_bf16.mul_(2.0)                      # overwritten in place (checkpoint load): the copy follows
    torch.testing.assert_close(T.frozen_transposed(lin).float(), lin.weight_bf16.float().t())


def test_bert_head_padded_vocabulary_matches_fp32_reference(dev):
    """Tied output projection on the tcgen05 kernels with the vocabulary padded to a multiple of 64 + fused cross-entropy that
    ignores / zeroes the padding columns, against the same model evaluated in fp32 on the CPU (unpadded slice of the table)."""
    import copy

    from vantage6_b200.models.bert import BertConfig, BertForMaskedLM, synthetic_mlm_batch

    torch.manual_seed(4)
    cfg = BertConfig(vocab_size=1000, hidden=128, layers=1, heads=2, ffn=256, max_pos=64)
    m_cpu = BertForMaskedLM(cfg)
    assert m_cpu.vocab_padded == 1024
    with torch.no_grad():
        m_cpu.head_bias.normal_(0.0, 0.2)
    m = copy.deepcopy(m_cpu).to(dev)
    ids, labels = synthetic_mlm_batch(1000, 8, 32, 40)
    ref = m_cpu(ids, labels)
    ref.backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = m(ids.to(dev), labels.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(ref.detach())) < 0.1, (float(loss.detach()), float(ref.detach()))      # bf16 GPU path vs fp32 CPU
    g_word, g_bias = m.word.weight.grad.float().cpu(), m.head_bias.grad.float().cpu()
    assert float(g_word[1000:].abs().max()) == 0.0 and float(g_bias[1000:].abs().max()) == 0.0       # padding rows: no gradient
    assert torch.nn.functional.cosine_similarity(g_bias[:1000], m_cpu.head_bias.grad[:1000], dim=0) > 0.97
    assert torch.nn.functional.cosine_similarity(g_word[:1000].flatten(), m_cpu.word.weight.grad[:1000].flatten(), dim=0) > 0.95
