"""tcgen05 implicit-GEMM convolution kernels (fwd / dgrad / wgrad, through the C ABI) against a plain
PyTorch fp32 reference of the same op evaluated on the bf16-rounded operands.

Tolerance: outputs are rounded to bf16 (rel 2^-8); accumulation is fp32 in TMEM.  We assert
|err| <= 1e-2 * max|ref| for bf16 outputs and 2e-3 * max|ref| for the fp32 weight gradient."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (N, H, W, Cin, Cout, ksize, stride)
SHAPES = [
    (4, 32, 32, 64, 160, 3, 1),     # WRN stage-1 first unit (16 channels padded to 64)
    (2, 32, 32, 160, 160, 3, 1),    # WRN stage 1
    (2, 32, 32, 160, 320, 3, 2),    # WRN stage-2 entry (stride 2)
    (2, 32, 32, 160, 320, 1, 2),    # WRN identity conv (1x1 stride 2)
    (3, 16, 16, 320, 320, 3, 1),    # WRN stage 2, odd batch
    (4, 8, 8, 640, 640, 3, 1),      # WRN stage 3 (two images per 128-pixel tile)
    (3, 8, 8, 320, 640, 3, 2),      # -> 4x4 outputs, partial tiles
    (2, 32, 32, 64, 64, 3, 1),      # ResNet18 layer1 (BLOCK_N = 64)
    (2, 16, 16, 128, 128, 3, 1),    # ResNet18 layer2 (BLOCK_N = 128)
    (2, 16, 16, 128, 256, 1, 2),    # ResNet18 shortcut
    (1, 64, 64, 64, 128, 3, 2),     # TinyImagenet-sized input
    (2, 8, 8, 256, 512, 3, 1),
    (2, 32, 32, 64, 160, 1, 1),
    (2, 16, 16, 160, 64, 3, 1),     # partial last k-block (160 = 2.5 x 64) with a 64-wide output
]


def make(N, H, W, cin, cout, k, stride, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((N, H, W, cin), device=DEV, generator=g).bfloat16()
    w = (torch.randn((cout, cin, k, k), device=DEV, generator=g) / (cin * k * k) ** 0.5).float()
    dy = torch.randn((N, H // stride, W // stride, cout), device=DEV, generator=g).bfloat16()
    return x, w, dy


def ref_conv(x, w, k, stride):
    xr = x.float().permute(0, 3, 1, 2)
    return F.conv2d(xr, w.bfloat16().float(), None, stride, k // 2)


def close(got, want, tol):
    scale = want.abs().max().item()
    err = (got.float() - want).abs().max().item()
    assert err <= tol * scale, f"max err {err:.4g} vs scale {scale:.4g} (tol {tol})"


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_fwd(shape):
    from nbdt._b200 import conv_ops as C

    N, H, W, cin, cout, k, stride = shape
    x, w, _ = make(*shape, seed=1)
    wp, wpt = C.pack_weights(w, cin_pad=cin, cout_pad=cout)
    y = C.conv_fwd(x, wp, k, stride)
    want = ref_conv(x, w, k, stride).permute(0, 2, 3, 1)
    close(y, want, 1e-2)
    # fused residual add
    res = torch.randn_like(y)
    y2 = C.conv_fwd(x, wp, k, stride, residual=res)
    close(y2, want + res.float(), 1e-2)


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_dgrad(shape):
    from nbdt._b200 import conv_ops as C

    N, H, W, cin, cout, k, stride = shape
    x, w, dy = make(*shape, seed=2)
    _, wpt = C.pack_weights(w, cin_pad=cin, cout_pad=cout)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv2d(xr, w.bfloat16().float(), None, stride, k // 2).backward(dy.float().permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1)
    dx = C.conv_dgrad(dy, wpt, (H, W), k, stride)
    close(dx, want, 1e-2)
    if not (k == 1 and stride == 2):
        base = torch.randn((N, H, W, cin), device=DEV).bfloat16()
        dx2 = C.conv_dgrad(dy, wpt, (H, W), k, stride, accumulate=base.clone())
        close(dx2, want + base.float(), 1e-2)


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_wgrad(shape):
    from nbdt._b200 import conv_ops as C

    N, H, W, cin, cout, k, stride = shape
    x, w, dy = make(*shape, seed=3)
    wr = w.bfloat16().float().requires_grad_(True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wr, None, stride, k // 2).backward(dy.float().permute(0, 3, 1, 2))
    dwp = C.conv_wgrad(x, dy, k, stride)
    assert dwp.shape == (k * k, cout, cin)
    got = C.unpack_wgrad(dwp, w.shape)
    close(got, wr.grad, 2e-3)
    acc = torch.ones_like(got)
    C.unpack_wgrad(dwp, w.shape, out=acc, accumulate=True)
    close(acc, wr.grad + 1, 2e-3)


def test_pack_weights_padding():
    from nbdt._b200 import conv_ops as C

    w = torch.randn((160, 16, 3, 3), device=DEV)
    wp, wpt = C.pack_weights(w, cin_pad=64, cout_pad=192)
    assert wp.shape == (9, 192, 64) and wpt.shape == (9, 64, 192)
    assert torch.equal(wp[:, :160, :16], w.permute(2, 3, 0, 1).reshape(9, 160, 16).bfloat16())
    assert (wp[:, 160:] == 0).all() and (wp[:, :, 16:] == 0).all()
    assert torch.equal(wpt, wp.transpose(1, 2))


# The CTA-pair kernels (tcgen05.mma.cta_group::2; H-halo variant for 16/32-wide stride-1 3x3) only engage from one pair
# tile per SM pair on; NBDT_B200_CONV_PAIR_MIN=1 forces them for the small parity shapes (incl. odd tile counts), and two
# shapes are large enough to take them by default.
@pytest.fixture
def force_pair_kernels(monkeypatch):
    monkeypatch.setenv("NBDT_B200_CONV_PAIR_MIN", "1")


@pytest.mark.parametrize("shape", SHAPES + [(3, 32, 32, 160, 160, 3, 1), (5, 16, 16, 320, 640, 3, 1), (1, 32, 32, 64, 64, 3, 1)])
def test_conv_fwd_dgrad_pair_kernels(shape, force_pair_kernels):
    test_conv_fwd(shape)
    test_conv_dgrad(shape)


@pytest.mark.parametrize("shape", [(160, 32, 32, 160, 160, 3, 1), (301, 8, 8, 128, 160, 3, 1), (80, 32, 32, 64, 128, 3, 2)])
def test_conv_fwd_dgrad_large(shape):
    """Default dispatch at sizes where the pair kernels are selected (>= 74 pair tiles), odd tile count included."""
    test_conv_fwd(shape)
    test_conv_dgrad(shape)
    test_conv_wgrad(shape)


def _check_fused_stats(shape, with_residual):
    """y of the statistics variant is bit-identical to the plain forward; the per-CTA partial rows add up to the per-channel sum
    and sum of squares of the stored bf16 values (what the BatchNorm statistics pass would have read)."""
    from nbdt._b200 import conv_ops as C

    N, H, W, cin, cout, k, stride = shape
    x, w, dy = make(*shape, seed=5)
    wp, _ = C.pack_weights(w, cin_pad=cin, cout_pad=cout)
    res = dy if with_residual else None
    y0 = C.conv_fwd(x, wp, k, stride, res)
    y, partials = C.conv_fwd(x, wp, k, stride, res, bn_stats=True)
    assert torch.equal(y, y0)
    assert partials is not None and partials.dim() == 3 and partials.shape[1:] == (2, cout)
    yd = y.double().reshape(-1, cout)
    got = partials.double().sum(0)
    want_s, want_q = yd.sum(0), (yd * yd).sum(0)
    tol_s = 2e-5 * yd.abs().sum(0).max().item() + 1e-6
    assert (got[0] - want_s).abs().max().item() <= tol_s, ((got[0] - want_s).abs().max().item(), tol_s)
    assert (got[1] - want_q).abs().max().item() <= 2e-5 * want_q.max().item() + 1e-6


@pytest.mark.parametrize("with_residual", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_fwd_fused_bn_statistics(shape, with_residual):
    _check_fused_stats(shape, with_residual)


@pytest.mark.parametrize("shape", [(3, 32, 32, 160, 160, 3, 1), (5, 16, 16, 320, 640, 3, 1), (1, 32, 32, 64, 64, 3, 1), (3, 8, 8, 320, 640, 3, 2),
                                   (160, 32, 32, 160, 160, 3, 1), (301, 8, 8, 128, 160, 3, 1)])
def test_conv_fwd_fused_bn_statistics_pair_kernels(shape, force_pair_kernels):
    """cta_group::2 kernels (with and without the H-halo tiles), odd tile counts, several tiles per CTA."""
    _check_fused_stats(shape, True)


def test_batchnorm_from_fused_statistics_equals_the_separate_pass(monkeypatch):
    """bn_act() on a convolution output that carries its partial sums gives the same result, coefficients and running
    statistics as the statistics kernel reading the tensor (the sums are the same numbers added in a different order)."""
    import torch.nn as nn
    from nbdt._b200 import backbone as B

    monkeypatch.setattr(B, "_FUSED_BN_STATS", True)  # opt-in (NBDT_B200_BN_FUSED_STATS=1)
    torch.manual_seed(0)
    conv = nn.Conv2d(160, 320, 3, stride=2, padding=1, bias=False).to(DEV).train()
    bns = [nn.BatchNorm2d(320).to(DEV).train() for _ in range(2)]
    x = torch.randn((6, 16, 16, 160), device=DEV).bfloat16()
    y_plain = B.conv(x, conv)
    y_stats = B.conv(x, conv, bn_stats=True)
    assert getattr(y_stats, "_nbdt_bn_partials", None) is not None and not hasattr(y_plain, "_nbdt_bn_partials")
    a = B.bn_act(y_plain, bns[0], relu=True)
    b = B.bn_act(y_stats, bns[1], relu=True)
    close(b, a.float(), 1e-2)
    assert (b.float() - a.float()).abs().max().item() <= 2 ** -6 * a.float().abs().max().item()
    torch.testing.assert_close(bns[1].running_mean, bns[0].running_mean, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(bns[1].running_var, bns[0].running_var, rtol=1e-4, atol=1e-6)
    # the gradient flows through both outputs of the convolution function
    w0 = conv.weight.detach().clone().requires_grad_(True)
    conv.weight = nn.Parameter(w0.detach().clone())
    B.bn_act(B.conv(x, conv, bn_stats=True), bns[1], relu=True).float().square().sum().backward()
    assert conv.weight.grad is not None and torch.isfinite(conv.weight.grad).all() and bns[1].weight.grad is not None
