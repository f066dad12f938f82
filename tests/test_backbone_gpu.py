"""Backbone kernels (BN / stem / pool / SGD) and whole networks on the B200 path against plain PyTorch
fp32 references of the same ops (oracle/torch_port.py eager modules).

Activations are bf16 between layers, so whole-network comparisons use the north-star's bf16 tolerance
(1e-2 relative to the tensor's scale) per layer and a cosine-similarity bound for end-to-end gradients."""
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


@pytest.mark.parametrize("shape", [(4, 8, 8, 64), (2, 32, 32, 160), (3, 5, 7, 640), (2, 16, 16, 16)])
@pytest.mark.parametrize("relu,use_res", [(True, False), (False, False), (True, True)])
def test_batchnorm_forward_backward(shape, relu, use_res):
    from nbdt._b200.nn_ops import BatchNormActFunction

    N, H, W, C = shape
    g = torch.Generator(device=DEV).manual_seed(C + relu + 2 * use_res)
    x = (torch.randn(shape, device=DEV, generator=g) * 1.7 + 0.3).bfloat16()
    res = torch.randn(shape, device=DEV, generator=g).bfloat16() if use_res else None
    dy = torch.randn(shape, device=DEV, generator=g).bfloat16()
    bn = nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()

    # reference: fp32 on the bf16-rounded operands, NCHW
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    rr = res.float().permute(0, 3, 1, 2).requires_grad_(True) if use_res else None
    ref_bn = nn.BatchNorm2d(C).to(DEV)
    ref_bn.load_state_dict(bn.state_dict())
    yr = ref_bn(xr)
    if use_res:
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    # gradient masks must come from OUR bf16 output (a reference value within rounding of 0 may flip)
    xs = x.clone().requires_grad_(True)
    rs = res.clone().requires_grad_(True) if use_res else None
    y = BatchNormActFunction.apply(xs, bn.weight, bn.bias, rm, rv, True, 0.1, bn.eps, relu, rs)
    close(y, yr.permute(0, 2, 3, 1), 1e-2, "bn fwd")
    close(rm, ref_bn.running_mean, 1e-3, "running mean")
    close(rv, ref_bn.running_var, 1e-3, "running var")

    mask = (y.float().permute(0, 3, 1, 2) > 0).float() if relu else 1.0
    # reference backward with our mask: d/dz of the pre-activation z
    z = ref_bn(xr) if not use_res else ref_bn(xr) + rr  # noqa: F841 (second forward only moves running stats)
    gz = dy.float().permute(0, 3, 1, 2) * mask
    ref_bn2 = nn.BatchNorm2d(C).to(DEV)
    ref_bn2.load_state_dict(bn.state_dict())
    xr2 = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref_bn2(xr2).backward(gz)
    y.backward(dy)
    close(xs.grad, xr2.grad.permute(0, 2, 3, 1), 1.5e-2, "bn dx")
    close(bn.weight.grad, ref_bn2.weight.grad, 5e-3, "dgamma")
    close(bn.bias.grad, ref_bn2.bias.grad, 5e-3, "dbeta")
    if use_res:
        close(rs.grad, gz.permute(0, 2, 3, 1), 1e-2, "dres")


def test_batchnorm_padded_channels_and_eval():
    from nbdt._b200.nn_ops import BatchNormActFunction

    x = torch.zeros((2, 4, 4, 64), device=DEV, dtype=torch.bfloat16)
    x[..., :16] = torch.randn((2, 4, 4, 16), device=DEV).bfloat16()
    bn = nn.BatchNorm2d(16).to(DEV)
    y = BatchNormActFunction.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, True, 0.1, bn.eps, True, None)
    assert (y[..., 16:] == 0).all()
    ref = F.relu(F.batch_norm(x[..., :16].float().permute(0, 3, 1, 2), None, None, bn.weight, bn.bias, True, 0.1, bn.eps))
    close(y[..., :16], ref.permute(0, 2, 3, 1), 1e-2)
    bn.eval()
    ye = BatchNormActFunction.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, False, 0.1, bn.eps, False, None)
    close(ye[..., :16], bn(x[..., :16].float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1), 1e-2)


@pytest.mark.parametrize("cout,c_phys,hw", [(16, 64, 32), (64, 64, 32), (64, 64, 64)])
def test_stem_and_pool(cout, c_phys, hw):
    from nbdt._b200.nn_ops import AvgPoolFunction, StemFunction

    g = torch.Generator(device=DEV).manual_seed(cout)
    x = torch.randn((3, 3, hw, hw), device=DEV, generator=g)
    w = (torch.randn((cout, 3, 3, 3), device=DEV, generator=g) * 0.2).requires_grad_(True)
    y = StemFunction.apply(x, w, c_phys)
    wr = w.detach().clone().requires_grad_(True)
    yr = F.conv2d(x, wr, None, 1, 1)
    close(y[..., :cout], yr.permute(0, 2, 3, 1), 1e-2, "stem fwd")
    assert (y[..., cout:] == 0).all()
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy[..., :cout].float().permute(0, 3, 1, 2))
    close(w.grad, wr.grad, 2e-3, "stem wgrad")

    a = torch.randn((5, 8, 8, 640), device=DEV).bfloat16().requires_grad_(True)
    feat = AvgPoolFunction.apply(a)
    close(feat, a.float().mean(dim=(1, 2)), 1e-5, "pool")
    df = torch.randn_like(feat)
    feat.backward(df)
    close(a.grad, (df / 64)[:, None, None, :].expand(5, 8, 8, 640), 1e-2, "pool bwd")


def _copy_state(src, dst):
    missing = dst.load_state_dict(src.state_dict())
    assert not missing.missing_keys and not missing.unexpected_keys


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("fused_bn_stats", [False, True])
@pytest.mark.parametrize("arch", ["ResNet18", "wrn"])
def test_network_forward_backward_matches_eager(arch, fused_bn_stats, monkeypatch):
    """Same parameters, same batch: B200 kernels (bf16 activations, fp32 accumulation) vs the fp32 eager oracle
    module.  Random-init 18/28-layer nets amplify bf16 rounding in the gradients, so the yardstick for the
    backward is the SAME eager module run under bf16 autocast (cuDNN): our per-parameter gradient agreement
    with fp32 must be as good as autocast's (scripts/diag_network.py prints both: medians ~0.95).
    fused_bn_stats: the same with the BatchNorm statistics accumulated in the convolution epilogues (NBDT_B200_BN_FUSED_STATS=1)."""
    from nbdt.loss import SoftTreeSupLoss
    from nbdt import models
    from nbdt._b200 import backbone as B
    from oracle import torch_port as P

    monkeypatch.setattr(B, "_FUSED_BN_STATS", fused_bn_stats)
    torch.manual_seed(0)
    if arch == "ResNet18":
        net, mk, hierarchy = models.ResNet18(num_classes=10).to(DEV), lambda: P.EagerResNet(num_classes=10), "induced-ResNet18"
    else:
        net, mk, hierarchy = models.wrn28_10_cifar10().to(DEV), lambda: P.EagerWRN(num_classes=10), "induced-wrn28_10_cifar10"
    ref, ref16 = mk().to(DEV), mk().to(DEV)
    _copy_state(net, ref)
    _copy_state(net, ref16)
    net.train(); ref.train(); ref16.train()
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn((32, 3, 32, 32), device=DEV, generator=g)
    y = torch.randint(10, (32,), device=DEV, generator=g)
    crit = SoftTreeSupLoss(dataset="CIFAR10", criterion=nn.CrossEntropyLoss(), hierarchy=hierarchy)
    out, out_ref = net(x), ref(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out16 = ref16(x).float()
    assert out.shape == (32, 10) and out.dtype == torch.float32
    close(out, out_ref, 4e-2, "logits")
    loss, loss_ref = crit(out, y), crit(out_ref, y)
    assert abs(loss.item() - loss_ref.item()) <= 3e-2 * abs(loss_ref.item())
    loss.backward(); loss_ref.backward(); crit(out16, y).backward()
    ours, floor = [], []
    for (name, p), (_, q), (_, r) in zip(net.named_parameters(), ref.named_parameters(), ref16.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        if q.grad.norm() > 1e-6:
            c_ours, c_floor = _cos(p.grad, q.grad), _cos(r.grad, q.grad)
            ours.append(c_ours); floor.append(c_floor)
            assert c_ours > c_floor - 0.05, f"{name}: cosine {c_ours:.4f} vs bf16-autocast {c_floor:.4f}"
            assert 0.8 < float(p.grad.norm() / q.grad.norm()) < 1.25, name
    ours.sort(); floor.sort()
    assert ours[len(ours) // 2] > floor[len(floor) // 2] - 0.01 and ours[len(ours) // 2] > 0.9
    # running statistics moved identically
    for (name, b), (_, c) in zip(net.named_buffers(), ref.named_buffers()):
        if b.dtype.is_floating_point:
            close(b, c, 2e-2, name)
        else:
            assert torch.equal(b, c), name
    # eval mode
    net.eval(); ref.eval()
    with torch.no_grad():
        close(net(x), ref(x), 4e-2, "eval logits")


def test_fused_sgd_matches_torch():
    from nbdt._b200.optim import FusedSGD

    torch.manual_seed(0)
    a = nn.Sequential(nn.Linear(7, 5), nn.Linear(5, 3)).to(DEV)
    b = nn.Sequential(nn.Linear(7, 5), nn.Linear(5, 3)).to(DEV)
    b.load_state_dict(a.state_dict())
    oa = torch.optim.SGD(a.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    ob = FusedSGD(b.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    for step in range(4):
        x = torch.randn(6, 7, device=DEV)
        for net, opt in ((a, oa), (b, ob)):
            opt.zero_grad()
            net(x).square().mean().backward()
            opt.step()
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), step
    ob.param_groups[0]["lr"] = 0.01  # schedulers mutate param_groups
    x = torch.randn(6, 7, device=DEV)
    for g in oa.param_groups:
        g["lr"] = 0.01
    for net, opt in ((a, oa), (b, ob)):
        opt.zero_grad(); net(x).square().mean().backward(); opt.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)
