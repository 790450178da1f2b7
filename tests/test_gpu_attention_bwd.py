"""K4 backward (tcgen05 dQ and dK/dV kernels) vs fp32 PyTorch autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,S,Hq,Hkv,D,causal", [
    (1, 128, 2, 2, 64, False),
    (2, 256, 4, 4, 64, True),
    (1, 512, 12, 12, 64, False),      # BERT-like
    (2, 384, 8, 2, 128, True),        # GQA, 3 tiles
    (1, 200, 2, 1, 128, False),       # S not a multiple of 128
    (1, 1024, 8, 2, 128, True),
])
def test_flash_attention_backward_native(B, S, Hq, Hkv, D, causal):
    from vantage6_b200.ops import attention as A
    from vantage6_b200.ops import native

    if not hasattr(native(), "flash_attn_bwd"):
        pytest.skip("extension built without the attention backward kernels")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    q = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, S, Hkv, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, Hq, D, device=dev, dtype=torch.bfloat16)
    o, lse = A.flash_attn_fwd(q, k, v, causal)
    dq, dk, dv = A.flash_attn_bwd(do, q, k, v, o, lse, causal)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().requires_grad_() for t in (q, k, v))
    ro, _ = A.reference_attention(qf, kf, vf, causal)
    ro.backward(do.float())
    for name, a, b in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        err = (a.float() - b).abs().max().item()
        ref = b.abs().max().item()
        assert err < 0.03 * max(1.0, ref), f"{name}: max abs err {err} (ref max {ref})"
