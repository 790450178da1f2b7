"""K8 GLM: the tensor-core kernel (csrc/glm_tc.cu, MN-major UMMA operand for X^T r) and the CUDA-core kernel against
the PyTorch reference; an opt-in timing print (``V6B200_EXPERIMENTAL=1``)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
opt_in = pytest.mark.skipif(os.environ.get("V6B200_EXPERIMENTAL") != "1", reason="set V6B200_EXPERIMENTAL=1")


@pytest.mark.parametrize("rows", [128, 1000, 125_000])
@pytest.mark.parametrize("kernel", ["tc", "cuda"])
def test_glm_kernels_match_reference(rows, kernel, monkeypatch):
    from vantage6_b200.ops import glm as K8
    from vantage6_b200.ops import native

    native()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    X = torch.randn(rows, 256, device=dev).to(torch.bfloat16)
    y = (torch.rand(rows, device=dev) < 0.4).float()
    w = torch.randn(257, device=dev) * 0.1
    ref = K8.reference_logistic_grad(X, y, w)
    monkeypatch.setenv("V6B200_GLM", kernel)
    out = K8.logistic_grad(X, y, w)
    torch.cuda.synchronize()
    scale = ref[:256].abs().max().item()
    assert (out[:256] - ref[:256]).abs().max().item() < 2e-2 * scale
    torch.testing.assert_close(out[256:259], ref[256:259], rtol=2e-3, atol=5e-2)


@opt_in
def test_glm_kernel_timing(monkeypatch, capsys):
    """Not an assertion on speed -- prints both kernels' time on the benchmark shape (1M x 256 bf16)."""
    from vantage6_b200.ops import glm as K8

    dev = torch.device("cuda", 0)
    X = torch.randn(1_000_000, 256, device=dev).to(torch.bfloat16)
    y = (torch.rand(1_000_000, device=dev) < 0.5).float()
    w = torch.zeros(257, device=dev)

    def timed():
        for _ in range(3):
            K8.logistic_grad(X, y, w)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            K8.logistic_grad(X, y, w)
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / 10

    monkeypatch.setenv("V6B200_GLM", "cuda")
    base = timed()
    monkeypatch.setenv("V6B200_GLM", "tc")
    tc = timed()
    with capsys.disabled():
        print(f"\n[glm] cuda-core {base * 1e3:.1f} us, tensor-core {tc * 1e3:.1f} us "
              f"({512e6 / tc / 1e6:.0f} GB/s of X)")


def test_fused_glm_iteration_matches_composed_path(monkeypatch):
    """Gradient kernel + one fold / all-reduce / update kernel (default) == the composed path (V6B200_GLM_FUSED=0)."""
    from vantage6_b200.models.glm import FederatedGLM, synthetic_glm_shard

    dev = torch.device("cuda", 0)
    X, y, _ = synthetic_glm_shard(20_000, 256, seed=3, device=dev, dtype=torch.bfloat16)

    def run(fused):
        monkeypatch.setenv("V6B200_GLM_FUSED", "1" if fused else "0")
        glm = FederatedGLM(X, y, 0, 1, lr=1.0)
        losses = [float(glm.step().item()) for _ in range(5)]
        w = glm.w.clone()
        glm.close()
        return losses, w

    (la, wa), (lb, wb) = run(True), run(False)
    assert max(abs(a - b) for a, b in zip(la, lb)) < 1e-3
    torch.testing.assert_close(wa, wb, rtol=1e-3, atol=1e-4)
