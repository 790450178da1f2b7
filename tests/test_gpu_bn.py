"""Fused BatchNorm(+add)(+ReLU) kernels vs PyTorch fp32 references; ResNet with fused vs stock BN."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from vantage6_b200.ops import native

    native()
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


@pytest.mark.parametrize("N,C,H,W", [(8, 64, 56, 56), (4, 256, 14, 14), (2, 2048, 7, 7), (16, 8, 16, 16), (3, 128, 5, 9)])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("with_res", [False, True])
def test_bn_fwd_bwd_matches_fp32_reference(dev, N, C, H, W, relu, with_res):
    from vantage6_b200.ops.bn import FusedBatchNormAct

    torch.manual_seed(0)
    cl = torch.channels_last
    x = (torch.randn(N, C, H, W, device=dev) * 2 + 0.5).to(torch.bfloat16).contiguous(memory_format=cl).requires_grad_()
    res = torch.randn(N, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=cl).requires_grad_() if with_res else None
    bn = FusedBatchNormAct(C, relu=relu).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    ref = torch.nn.BatchNorm2d(C).to(dev)
    ref.load_state_dict(bn.state_dict())
    y = bn(x, res)
    dy = torch.randn_like(y)
    y.backward(dy)
    xf = x.detach().float().requires_grad_()
    rf = res.detach().float().requires_grad_() if with_res else None
    yr = ref(xf)
    if with_res:
        yr = yr + rf
    if relu:
        yr = torch.relu(yr)
    yr.backward(dy.float())
    assert y.is_contiguous(memory_format=cl)
    torch.testing.assert_close(y.float(), yr, rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(x.grad.float(), xf.grad, rtol=5e-2, atol=5e-2)
    if with_res:
        torch.testing.assert_close(res.grad.float(), rf.grad, rtol=2e-2, atol=2e-2)
    scale = max(1.0, ref.weight.grad.abs().max().item())
    assert (bn.weight.grad - ref.weight.grad).abs().max().item() < 3e-2 * scale
    assert (bn.bias.grad - ref.bias.grad).abs().max().item() < 3e-2 * max(1.0, ref.bias.grad.abs().max().item())
    torch.testing.assert_close(bn.running_mean, ref.running_mean, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(bn.running_var, ref.running_var, rtol=1e-2, atol=1e-2)
    assert int(bn.num_batches_tracked) == 1


def test_bn_eval_mode_uses_running_stats(dev):
    from vantage6_b200.ops.bn import FusedBatchNormAct

    torch.manual_seed(1)
    bn = FusedBatchNormAct(64).to(dev)
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2.0)
    bn.eval()
    x = torch.randn(4, 64, 8, 8, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = bn(x)
    ref = torch.relu(torch.nn.functional.batch_norm(x.float(), bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.1, bn.eps))
    torch.testing.assert_close(y.float(), ref, rtol=2e-2, atol=2e-2)


def test_resnet_fused_bn_matches_stock_bn_training(dev):
    from vantage6_b200.models.resnet import imagenet_forward_loss, resnet_tiny

    def run(fused):
        torch.manual_seed(3)
        m = resnet_tiny(10, fused_bn=fused).to(dev).to(memory_format=torch.channels_last)
        opt = torch.optim.SGD(m.parameters(), lr=0.05)
        g = torch.Generator(device="cpu").manual_seed(5)
        x = torch.randint(0, 256, (16, 3, 64, 64), dtype=torch.uint8, generator=g).to(dev)
        y = torch.randint(0, 10, (16,), generator=g).to(dev)
        losses = []
        for _ in range(4):
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = imagenet_forward_loss(m, x, y)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        return losses

    a, b = run(True), run(False)
    assert abs(a[0] - b[0]) < 0.05 and abs(a[-1] - b[-1]) < 0.25, (a, b)
    assert a[-1] < a[0]
