"""K4 (tcgen05 flash attention) numerics vs an fp32 PyTorch reference; K1 / model-level GPU tests."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from vantage6_b200.ops import native

    native()
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


@pytest.mark.parametrize("B,S,Hq,Hkv,D,causal", [
    (2, 128, 12, 12, 64, False),      # BERT-base
    (2, 512, 12, 12, 64, False),
    (1, 256, 4, 4, 64, True),
    (2, 1024, 8, 2, 128, True),       # Llama-style GQA
    (1, 384, 4, 1, 128, False),       # several KV tiles, odd tile count
    (1, 200, 2, 2, 64, True),         # S not a multiple of 128
    (1, 2048, 32, 8, 128, True),
])
@pytest.mark.parametrize("variant", ["1cta", "2cta"])
def test_flash_attention_forward(dev, B, S, Hq, Hkv, D, causal, variant):
    from vantage6_b200.ops import attention as A

    torch.manual_seed(0)
    q = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
    o, lse = A.flash_attn_fwd(q, k, v, causal, variant=variant)
    torch.cuda.synchronize()
    ro, rlse = A.reference_attention(q, k, v, causal)
    err = (o.float() - ro).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"
    torch.testing.assert_close(lse, rlse, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("variant", ["1cta", "2cta"])
def test_flash_attention_large_logits_trigger_rescale(dev, variant):
    """Scores grow along the key axis so the running max is raised repeatedly (lazy-rescale path)."""
    from vantage6_b200.ops import attention as A

    torch.manual_seed(1)
    B, S, H, D = 1, 1024, 2, 64
    q = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    ramp = torch.linspace(0.5, 6.0, S, device=dev)[None, :, None, None]
    k = (k * ramp).to(torch.bfloat16)
    v = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    o, lse = A.flash_attn_fwd(q, k, v, False, variant=variant)
    ro, rlse = A.reference_attention(q, k, v, False)
    assert (o.float() - ro).abs().max().item() < 5e-2
    torch.testing.assert_close(lse, rlse, rtol=1e-3, atol=5e-3)


def test_flash_attention_backward_matches_reference(dev):
    from vantage6_b200.ops import attention as A

    torch.manual_seed(2)
    B, S, Hq, Hkv, D = 2, 256, 8, 2, 128
    q = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    do = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
    o = A.flash_attention(q, k, v, causal=True)
    o.backward(do)
    qf, kf, vf = (t.detach().float().requires_grad_() for t in (q, k, v))
    ro, _ = A.reference_attention(qf, kf, vf, True)
    ro.backward(do.float())
    for a, b in ((q.grad, qf.grad), (k.grad, kf.grad), (v.grad, vf.grad)):
        assert (a.float() - b).abs().max().item() < 0.1 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("name", ["bert_tiny", "llama_tiny_lora"])
def test_zoo_transformers_train_on_gpu(dev, name):
    from vantage6_b200.models import zoo

    torch.manual_seed(0)
    tr, spec = zoo.build_trainer(name, rank=0, world=1, device=dev)
    tr.initialize_global()
    batches = [(x.to(dev), y.to(dev)) for x, y in spec.make_batches(spec.local_steps, spec.batch, seed=3)]
    losses = [tr.run_round(batches).item() for _ in range(3)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    assert tr.engine.poll_status() == 0
    tr.close()


def test_glm_fused_converges_on_gpu(dev):
    from vantage6_b200.models.glm import FederatedGLM, synthetic_glm_shard

    X, y, w_true = synthetic_glm_shard(50_000, 256, seed=1, device=dev)
    glm = FederatedGLM(X, y, 0, 1, lr=2.0)
    l0 = glm.step().item()
    for _ in range(60):
        glm.step()
    assert glm.last_loss.item() < 0.8 * l0
    assert (glm.w - w_true).abs().max().item() < 0.6
    glm.close()


@pytest.mark.parametrize("variant", ["1cta", "2cta"])
def test_packed_qkv_attention_forward_backward(dev, variant, monkeypatch):
    """q / k consumed in place from a packed [B,S,3,H,D] projection output; dq/dk/dv written into one dqkv buffer."""
    from vantage6_b200.ops import attention as A

    monkeypatch.setenv("V6B200_ATTN_FWD", variant)
    torch.manual_seed(5)
    B, S, H, D = 2, 256, 12, 64
    qkv = torch.randn(B, S, 3, H, D, device=dev, dtype=torch.bfloat16).requires_grad_()
    o = A.packed_qkv_attention(qkv, False)
    do = torch.randn_like(o)
    o.backward(do)
    ref = qkv.detach().float().requires_grad_()
    ro, _ = A.reference_attention(ref[:, :, 0], ref[:, :, 1], ref[:, :, 2], False)
    ro.backward(do.float())
    assert (o.float() - ro).abs().max().item() < 3e-2
    assert (qkv.grad.float() - ref.grad).abs().max().item() < 6e-2


@pytest.mark.parametrize("mode", ["mn", "t"])
@pytest.mark.parametrize("B,S,Hq,Hkv,D,causal", [(2, 256, 12, 12, 64, False), (1, 1024, 8, 2, 128, True), (1, 200, 2, 2, 64, True)])
def test_attention_v_in_place_and_transposed(dev, B, S, Hq, Hkv, D, causal, mode, monkeypatch):
    """V consumed in its natural layout as an MN-major UMMA operand (default) and through the transposed copy (V6B200_ATTN_V=t)."""
    from vantage6_b200.ops import attention as A

    monkeypatch.setenv("V6B200_ATTN_V", mode)
    torch.manual_seed(7)
    q = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
    o, lse = A.flash_attn_fwd(q, k, v, causal, variant="2cta")
    ro, rlse = A.reference_attention(q, k, v, causal)
    assert (o.float() - ro).abs().max().item() < 3e-2
    torch.testing.assert_close(lse, rlse, rtol=1e-3, atol=2e-3)
