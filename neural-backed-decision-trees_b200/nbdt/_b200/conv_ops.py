"""Tensor-level wrappers over the tcgen05 implicit-GEMM convolution entry points (csrc/conv_gemm.cu).

Activations are NHWC bf16 tensors of shape (N, H, W, C) with C a multiple of 64; weights stay fp32 OIHW
(`nn.Conv2d.weight`) and are repacked to bf16 tap-major matrices per call."""
import torch

from . import lib as L
from .ops import _count, _p


PROFILE = None  # bench.py sets this to a list to collect (kind, flops, start_event, end_event) per conv launch


class _timed:
    def __init__(self, kind, flops):
        self.kind, self.flops = kind, flops

    def __enter__(self):
        if PROFILE is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            PROFILE.append((self.kind, self.flops, self.start, end))


def pad64(c):
    return (c + 63) // 64 * 64


def _check_act(t, name):
    L.require_cuda(t, name)
    if t.dtype != torch.bfloat16 or t.dim() != 4 or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous NHWC bf16 tensor, got {t.dtype} {tuple(t.shape)}")


def pack_weights(weight, cin_pad=None, cout_pad=None, want=(True, True)):
    """fp32 OIHW -> (wpack [taps][cout_pad][cin_pad], wpack_t [taps][cin_pad][cout_pad]) bf16."""
    L.require_cuda(weight, "pack_weights")
    cout, cin, k, _ = weight.shape
    cin_pad, cout_pad = cin_pad or cin, cout_pad or cout
    w = weight.detach().float().contiguous()
    wp = torch.empty((k * k, cout_pad, cin_pad), dtype=torch.bfloat16, device=w.device) if want[0] else None
    wpt = torch.empty((k * k, cin_pad, cout_pad), dtype=torch.bfloat16, device=w.device) if want[1] else None
    with torch.cuda.device(w.device):
        L.check(L.load().nbdt_conv_pack_weights(_p(w), _p(wp), _p(wpt), cout, cin, cout_pad, cin_pad, k,
                                                L.stream_ptr(w.device)))
    _count()
    return wp, wpt


def conv_fwd(x, wp, ksize, stride, residual=None, real=None):
    """real = (Cout, Cin) of the layer: the FLOP count bench.py reports is the algorithmic one, not that of the padded channels."""
    _check_act(x, "conv_fwd input")
    N, H, W, cin = x.shape
    taps, cout, cin_w = wp.shape
    assert taps == ksize * ksize and cin_w == cin, (wp.shape, x.shape, ksize)
    y = torch.empty((N, H // stride, W // stride, cout), dtype=torch.bfloat16, device=x.device)
    if residual is not None:
        _check_act(residual, "conv_fwd residual")
        assert residual.shape == y.shape
    flops = 2.0 * y.numel() * cin * ksize * ksize if real is None else 2.0 * N * (H // stride) * (W // stride) * real[0] * real[1] * ksize * ksize
    with torch.cuda.device(x.device), _timed("fwd", flops):
        L.check(L.load().nbdt_conv2d_fwd(_p(x), _p(wp), _p(y), _p(residual), N, H, W, cin, cout, ksize, stride,
                                         L.stream_ptr(x.device)))
    _count()
    return y


def conv_dgrad(dy, wpt, in_hw, ksize, stride, accumulate=None, real=None):
    """dx (N, H, W, Cin) for dy (N, H/s, W/s, Cout); `accumulate` (same shape as dx) is added and
    overwritten in place."""
    _check_act(dy, "conv_dgrad dy")
    N, Ho, Wo, cout = dy.shape
    H, W = in_hw
    taps, cin, cout_w = wpt.shape
    assert taps == ksize * ksize and cout_w == cout and Ho == H // stride and Wo == W // stride
    dx = accumulate if accumulate is not None else torch.empty((N, H, W, cin), dtype=torch.bfloat16, device=dy.device)
    if accumulate is not None:
        _check_act(accumulate, "conv_dgrad accumulate")
        assert accumulate.shape == (N, H, W, cin)
    flops = 2.0 * dy.numel() * cin * ksize * ksize if real is None else 2.0 * N * Ho * Wo * real[0] * real[1] * ksize * ksize
    with torch.cuda.device(dy.device), _timed("dgrad", flops):
        L.check(L.load().nbdt_conv2d_dgrad(_p(dy), _p(wpt), _p(dx), _p(accumulate), N, H, W, cin, cout, ksize, stride,
                                           L.stream_ptr(dy.device)))
    _count(4 if stride == 2 else 1)
    return dx


def conv_wgrad(x, dy, ksize, stride, real=None):
    """fp32 [taps][Cout][Cin] weight gradient (padded channel counts, as the activations carry them)."""
    _check_act(x, "conv_wgrad x")
    _check_act(dy, "conv_wgrad dy")
    N, H, W, cin = x.shape
    cout = dy.shape[3]
    assert dy.shape == (N, H // stride, W // stride, cout)
    dwp = torch.empty((ksize * ksize, cout, cin), dtype=torch.float32, device=x.device)
    flops = 2.0 * dy.numel() * cin * ksize * ksize if real is None else 2.0 * N * (H // stride) * (W // stride) * real[0] * real[1] * ksize * ksize
    with torch.cuda.device(x.device), _timed("wgrad", flops):
        L.check(L.load().nbdt_conv2d_wgrad(_p(x), _p(dy), _p(dwp), N, H, W, cin, cout, ksize, stride,
                                           L.stream_ptr(x.device)))
    _count()
    return dwp


def unpack_wgrad(dwp, weight_shape, out=None, accumulate=False):
    cout, cin, k, _ = weight_shape
    taps, cout_pad, cin_pad = dwp.shape
    if out is None:
        out = torch.empty(weight_shape, dtype=torch.float32, device=dwp.device)
        accumulate = False
    with torch.cuda.device(dwp.device):
        L.check(L.load().nbdt_conv_unpack_wgrad(_p(dwp), _p(out), cout, cin, cout_pad, cin_pad, k, int(accumulate),
                                                L.stream_ptr(dwp.device)))
    _count()
    return out
