"""EfficientNet-B0 on the B200 inference kernels against torchvision's own eager module -- the module
the reference star-imports (`nbdt/models/__init__.py:3-4`) -- and the MBConv kernels against plain PyTorch fp32
references of the same ops.  bf16 activations: tolerances are relative to the tensor's scale."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(got, want, tol, what=""):
    scale = max(want.abs().max().item(), 1e-6)
    err = (got.float() - want.float()).abs().max().item()
    assert err <= tol * scale, f"{what} max err {err:.4g} vs scale {scale:.4g} (tol {tol})"


def _bn(c, g):
    bn = nn.BatchNorm2d(c).to(DEV).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
        bn.running_mean.uniform_(-0.3, 0.3, generator=g)
        bn.running_var.uniform_(0.5, 2.0, generator=g)
    return bn


@pytest.mark.parametrize("k,s", [(3, 1), (3, 2), (5, 1), (5, 2)])
@pytest.mark.parametrize("shape", [(2, 14, 14, 64), (3, 7, 9, 128), (1, 33, 18, 192), (2, 4, 4, 704)])
def test_depthwise_conv_bn_silu(k, s, shape):
    from nbdt._b200 import mbconv_ops as M

    N, H, W, C = shape
    c_real = C - 24  # padded physical channels stay zero
    g = torch.Generator(device=DEV).manual_seed(k * 10 + s + C)
    x = torch.randn(shape, device=DEV, generator=g).bfloat16()
    x[..., c_real:] = 0
    conv = nn.Conv2d(c_real, c_real, k, s, padding=k // 2, groups=c_real, bias=False).to(DEV)
    with torch.no_grad():
        conv.weight.normal_(0, 0.4, generator=g)
    bn = _bn(c_real, g)
    with torch.no_grad():
        y = M.dwconv_bn_act(x, conv, bn, M.ACT_SILU)
        want = F.silu(bn(conv(x[..., :c_real].float().permute(0, 3, 1, 2)))).permute(0, 2, 3, 1)
    assert y.shape == (N, (H + s - 1) // s, (W + s - 1) // s, C)
    close(y[..., :c_real], want, 1e-2, f"dwconv k{k} s{s}")
    assert (y[..., c_real:] == 0).all()


@pytest.mark.parametrize("shape", [(3, 7, 7, 64), (2, 16, 16, 192), (5, 1, 1, 1280), (2, 3, 5, 2112)])
def test_se_scale_and_bn_silu(shape):
    from nbdt._b200 import mbconv_ops as M

    N, H, W, C = shape
    g = torch.Generator(device=DEV).manual_seed(C)
    x = torch.randn(shape, device=DEV, generator=g).bfloat16()
    s = torch.rand((N, C), device=DEV, generator=g)
    close(M.se_scale(x, s), x.float() * s[:, None, None, :], 1e-2, "se scale")
    bn = _bn(C, g)
    res = torch.randn(shape, device=DEV, generator=g).bfloat16()
    with torch.no_grad():
        z = bn(x.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        close(M.bn_act_eval(x, bn, M.ACT_SILU), F.silu(z), 1e-2, "bn+silu")
        close(M.bn_act_eval(x, bn, M.ACT_NONE, residual=res), z + res.float(), 1e-2, "bn+residual")


@pytest.mark.parametrize("stride,hw", [(2, (32, 32)), (2, (15, 21)), (1, (9, 12))])
def test_strided_stem(stride, hw):
    from nbdt._b200 import mbconv_ops as M

    g = torch.Generator(device=DEV).manual_seed(stride)
    x = torch.randn((3, 3) + hw, device=DEV, generator=g)
    conv = nn.Conv2d(3, 32, 3, stride, padding=1, bias=False).to(DEV)
    with torch.no_grad():
        y = M.stem(x, conv, 64)
        want = conv(x).permute(0, 2, 3, 1)
    close(y[..., :32], want, 1e-2, "stem")
    assert (y[..., 32:] == 0).all()


@pytest.mark.parametrize("cin,cout", [(16, 96), (96, 24), (40, 240), (320, 1280)])
def test_pointwise_conv_padded_channels(cin, cout):
    from nbdt._b200 import conv_ops as C
    from nbdt._b200 import mbconv_ops as M

    g = torch.Generator(device=DEV).manual_seed(cin + cout)
    cp = C.pad64(cin)
    x = torch.zeros((3, 6, 10, cp), device=DEV, dtype=torch.bfloat16)
    x[..., :cin] = torch.randn((3, 6, 10, cin), device=DEV, generator=g).bfloat16()
    conv = nn.Conv2d(cin, cout, 1, bias=False).to(DEV)
    with torch.no_grad():
        conv.weight.normal_(0, 0.2, generator=g)
        y = M.conv1x1(x, conv)
        want = F.conv2d(x[..., :cin].float().permute(0, 3, 1, 2), conv.weight.bfloat16().float()).permute(0, 2, 3, 1)
    assert y.shape == (3, 6, 10, C.pad64(cout))
    close(y[..., :cout], want, 1e-2, "1x1 conv")
    assert (y[..., cout:] == 0).all()
    # cache: second call reuses the packed weight; an in-place update invalidates it
    with torch.no_grad():
        store = conv.__dict__["_nbdt_b200_cache"]
        first = store[f"wp{cp}"][1]
        M.conv1x1(x, conv)
        assert store[f"wp{cp}"][1] is first
        conv.weight.mul_(2.0)
        close(M.conv1x1(x, conv)[..., :cout], 2 * want, 1e-2, "1x1 conv after update")


@pytest.mark.parametrize("size,batch", [(64, 3), (224, 2)])
def test_efficientnet_b0_matches_torchvision(size, batch):
    import torchvision
    from nbdt.models import efficientnet_b0

    torch.manual_seed(0)
    ref = torchvision.models.efficientnet_b0(weights=None).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for m in ref.modules():  # non-trivial BN statistics / affine parameters
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.uniform_(-0.2, 0.2, generator=g)
                m.running_var.uniform_(0.6, 1.6, generator=g)
                m.weight.uniform_(0.7, 1.3, generator=g)
                m.bias.uniform_(-0.2, 0.2, generator=g)
    net = efficientnet_b0().to(DEV).eval()
    net.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(batch, 3, size, size, generator=g)
    with torch.no_grad():
        want = ref(x)
        got = net(x.to(DEV))
    assert got.shape == (batch, 1000) and got.dtype == torch.float32
    close(got.cpu(), want, 4e-2, f"efficientnet_b0 logits @{size}")
    assert (got.argmax(1).cpu() == want.argmax(1)).float().mean() >= 0.5 or size == 64


def test_soft_nbdt_on_efficientnet():
    """BASELINE config 5 shape: EfficientNet-B0 + SoftNBDT on the ImageNet-1k induced hierarchy."""
    from nbdt.model import HardNBDT, SoftNBDT
    from nbdt.models import efficientnet_b0

    torch.manual_seed(0)
    net = efficientnet_b0().to(DEV).eval()
    soft = SoftNBDT(dataset="Imagenet1000", model=net, hierarchy="induced-efficientnet_b7b")
    hard = HardNBDT(dataset="Imagenet1000", model=net, hierarchy="induced-efficientnet_b7b")
    x = torch.randn(4, 3, 64, 64, device=DEV)
    with torch.no_grad():
        p = soft(x)
        h = hard(x)
    assert p.shape == (4, 1000) and torch.allclose(p.sum(1), torch.ones(4, device=DEV), atol=1e-4)
    assert h.shape == (4, 1000) and (h.sum(1) == 1).all()


def test_inference_ops_refuse_autograd():
    """The fused inference ops (folded BN, cached packed weights) must not silently detach trainable parameters."""
    from nbdt._b200 import mbconv_ops as M

    conv = nn.Conv2d(64, 64, 1, bias=False).to(DEV)
    x = torch.zeros((1, 4, 4, 64), device=DEV, dtype=torch.bfloat16)
    with pytest.raises(NotImplementedError):
        M.conv1x1(x, conv)
    with torch.no_grad():
        M.conv1x1(x, conv)


@pytest.mark.parametrize("k,s", [(3, 1), (3, 2), (5, 1), (5, 2)])
@pytest.mark.parametrize("shape", [(2, 14, 14, 64), (3, 7, 9, 128), (2, 16, 12, 704)])
def test_depthwise_conv_training(k, s, shape):
    from nbdt._b200 import mbconv_ops as M

    N, H, W, C = shape
    c_real = C - 16
    g = torch.Generator(device=DEV).manual_seed(k * 7 + s + C)
    x = torch.randn(shape, device=DEV, generator=g).bfloat16()
    x[..., c_real:] = 0
    w = (torch.randn((c_real, 1, k, k), device=DEV, generator=g) * 0.3).requires_grad_(True)
    xs = x.clone().requires_grad_(True)
    y = M.DepthwiseConvFunction.apply(xs, w, s)
    dy = torch.randn(y.shape, device=DEV, generator=g).bfloat16()
    dy[..., c_real:] = 0
    y.backward(dy)
    xr = x[..., :c_real].float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, s, k // 2, groups=c_real)
    yr.backward(dy[..., :c_real].float().permute(0, 3, 1, 2))
    close(y[..., :c_real], yr.permute(0, 2, 3, 1), 1e-2, "dw fwd")
    close(xs.grad[..., :c_real], xr.grad.permute(0, 2, 3, 1), 1e-2, "dw dgrad")
    close(w.grad, wr.grad, 2e-3, "dw wgrad")
    assert (xs.grad[..., c_real:] == 0).all()


@pytest.mark.parametrize("shape", [(3, 7, 7, 64), (2, 28, 28, 192), (4, 1, 1, 1280)])
def test_se_scale_training(shape):
    from nbdt._b200 import mbconv_ops as M

    N, H, W, C = shape
    g = torch.Generator(device=DEV).manual_seed(C + H)
    x = torch.randn(shape, device=DEV, generator=g).bfloat16().requires_grad_(True)
    sc = torch.rand((N, C), device=DEV, generator=g).requires_grad_(True)
    dy = torch.randn(shape, device=DEV, generator=g).bfloat16()
    M.SEScaleFunction.apply(x, sc).backward(dy)
    close(x.grad, dy.float() * sc.detach()[:, None, None, :], 1e-2, "se dx")
    close(sc.grad, (dy.float() * x.detach().float()).sum((1, 2)), 2e-3, "se ds")


@pytest.mark.parametrize("shape", [(4, 8, 8, 64), (2, 14, 14, 128), (3, 5, 7, 704)])
def test_batchnorm_silu_training(shape):
    from nbdt._b200.nn_ops import BatchNormActFunction

    N, H, W, C = shape
    c_real = C - 8
    g = torch.Generator(device=DEV).manual_seed(C)
    x = (torch.randn(shape, device=DEV, generator=g) * 1.5 + 0.2).bfloat16()
    x[..., c_real:] = 0
    dy = torch.randn(shape, device=DEV, generator=g).bfloat16()
    bn = nn.BatchNorm2d(c_real).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
    ref = nn.BatchNorm2d(c_real).to(DEV)
    ref.load_state_dict(bn.state_dict())
    xs = x.clone().requires_grad_(True)
    y = BatchNormActFunction.apply(xs, bn.weight, bn.bias, bn.running_mean, bn.running_var, True, 0.1, bn.eps, 2, None)
    y.backward(dy)
    xr = x[..., :c_real].float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.silu(ref(xr))
    yr.backward(dy[..., :c_real].float().permute(0, 3, 1, 2))
    close(y[..., :c_real], yr.permute(0, 2, 3, 1), 1e-2, "bn+silu fwd")
    close(xs.grad[..., :c_real], xr.grad.permute(0, 2, 3, 1), 2e-2, "bn+silu dx")
    close(bn.weight.grad, ref.weight.grad, 1e-2, "dgamma")
    close(bn.bias.grad, ref.bias.grad, 1e-2, "dbeta")
    close(bn.running_var, ref.running_var, 1e-3, "running var")


def test_efficientnet_b0_train_step_matches_torchvision():
    """One training forward/backward (batch statistics, no stochastic depth / dropout so that it is deterministic) against
    torchvision's eager module in fp32: logits, and the gradients of parameters along the whole depth."""
    import torchvision
    from nbdt.models import efficientnet_b0

    torch.manual_seed(0)
    ref = torchvision.models.efficientnet_b0(weights=None, dropout=0.0, stochastic_depth_prob=0.0).train()
    net = efficientnet_b0(dropout=0.0, stochastic_depth_prob=0.0).to(DEV).train()
    net.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 3, 128, 128, generator=g)  # 4x4 final maps: 96 samples per channel in the last BatchNorms
    t = torch.randint(1000, (6,), generator=g)
    F.cross_entropy(ref(x), t).backward()
    out = net(x.to(DEV))
    F.cross_entropy(out, t.to(DEV)).backward()
    want = ref(x).detach()  # second forward only moves the running statistics of the reference
    # batch-statistics BatchNorm re-normalises the bf16 rounding noise of 81 layers: looser than the eval-mode 4e-2
    close(out.detach().cpu(), want, 1e-1, "train-mode logits")
    refp = dict(ref.named_parameters())
    checked = 0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        if name.endswith("bias") and "fc" not in name and "classifier" not in name:
            continue
        gr = refp[name].grad.flatten()
        cos = F.cosine_similarity(p.grad.detach().cpu().flatten().double(), gr.double(), dim=0).item()
        lim = 0.8 if p.grad.numel() < 512 else 0.9
        assert cos > lim, f"{name}: gradient cosine {cos:.3f}"
        checked += 1
    assert checked > 100


def test_efficientnet_stochastic_depth_and_dropout_run():
    from nbdt.models import efficientnet_b0

    torch.manual_seed(0)
    net = efficientnet_b0(num_classes=10).to(DEV).train()
    out = net(torch.randn(4, 3, 32, 32, device=DEV))
    out.sum().backward()
    assert torch.isfinite(out).all() and all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
