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


def test_training_mode_raises():
    from nbdt.models import efficientnet_b0

    net = efficientnet_b0().to(DEV)
    with pytest.raises(NotImplementedError):
        net.train()(torch.randn(2, 3, 32, 32, device=DEV))
    with pytest.raises(NotImplementedError):  # eval mode but autograd on trainable parameters
        net.eval()(torch.randn(2, 3, 32, 32, device=DEV))
