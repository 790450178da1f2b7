"""2-CTA (cta_group::2) tcgen05 GEMM numerics vs fp32 PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 256), (4096, 2304, 768), (1000, 3072, 768), (300, 264, 136),
                                   (8192, 4096, 4096)])
def test_gemm_2cta(M, N, K):
    from vantage6_b200.ops import gemm as G
    from vantage6_b200.ops import native

    if not hasattr(native(), "gemm2_bf16"):
        pytest.skip("extension built without the 2-CTA kernel")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev)
    c = G.gemm_bf16(a, w, bias, G.ACT_NONE, variant="2cta")
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    err = ((c.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, f"rel err {err}"
    c2 = G.gemm_bf16(a, w, bias, G.ACT_RELU, variant="2cta")
    err2 = ((c2.float() - torch.relu(ref)).abs().max() / ref.abs().max()).item()
    assert err2 < 1e-2
