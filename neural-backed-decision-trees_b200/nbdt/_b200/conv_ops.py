"""Tensor-level wrappers over the tcgen05 implicit-GEMM convolution entry points (csrc/conv_gemm.cu).

Activations are NHWC bf16 tensors of shape (N, H, W, C) with C a multiple of 64; weights stay fp32 OIHW
(`nn.Conv2d.weight`) and are repacked to bf16 tap-major matrices per call."""
import ctypes

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


def conv_fwd(x, wp, ksize, stride, residual=None, real=None, bn_stats=False):
    """real = (Cout, Cin) of the layer: the FLOP count bench.py reports is the algorithmic one, not that of the padded channels.
    bn_stats: also accumulate the BatchNorm batch statistics of y in the epilogue; returns (y, partials) where partials is None
    (too many channels for the fused path) or a fp32 tensor [parts][2][Cout] of per-CTA partial sums."""
    _check_act(x, "conv_fwd input")
    N, H, W, cin = x.shape
    taps, cout, cin_w = wp.shape
    assert taps == ksize * ksize and cin_w == cin, (wp.shape, x.shape, ksize)
    y = torch.empty((N, H // stride, W // stride, cout), dtype=torch.bfloat16, device=x.device)
    if residual is not None:
        _check_act(residual, "conv_fwd residual")
        assert residual.shape == y.shape
    flops = 2.0 * y.numel() * cin * ksize * ksize if real is None else 2.0 * N * (H // stride) * (W // stride) * real[0] * real[1] * ksize * ksize
    lib = L.load()
    partials = None
    with torch.cuda.device(x.device):
        nfloats = lib.nbdt_conv2d_stats_floats(cout) if bn_stats else 0
        buf = torch.empty((nfloats,), dtype=torch.float32, device=x.device) if nfloats > 0 else None
        parts = ctypes.c_int(0)
        with _timed("fwd", flops):
            if nfloats > 0:
                L.check(lib.nbdt_conv2d_fwd_stats(_p(x), _p(wp), _p(y), _p(residual), N, H, W, cin, cout, ksize, stride, _p(buf),
                                                  ctypes.byref(parts), L.stream_ptr(x.device)))
            else:
                L.check(lib.nbdt_conv2d_fwd(_p(x), _p(wp), _p(y), _p(residual), N, H, W, cin, cout, ksize, stride,
                                            L.stream_ptr(x.device)))
        if nfloats > 0:
            partials = buf[: parts.value * 2 * cout].view(parts.value, 2, cout)
    _count()
    return (y, partials) if bn_stats else y


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


def conv_wgrad(x, dy, ksize, stride, real=None, out=None):
    """fp32 [taps][Cout][Cin] weight gradient (padded channel counts, as the activations carry them); `out`: buffer to overwrite."""
    _check_act(x, "conv_wgrad x")
    _check_act(dy, "conv_wgrad dy")
    N, H, W, cin = x.shape
    cout = dy.shape[3]
    assert dy.shape == (N, H // stride, W // stride, cout)
    dwp = out if out is not None else torch.empty((ksize * ksize, cout, cin), dtype=torch.float32, device=x.device)
    assert dwp.shape == (ksize * ksize, cout, cin) and dwp.dtype == torch.float32 and dwp.is_contiguous()
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


class ConvWeightBank:
    """Persistent packed forms of every convolution weight a FusedSGD owns, maintained with ONE launch per step each way:

      * bf16 tap-major matrices wp / wpt (what the tcgen05 kernels read), re-packed from the fp32 master weights by
        nbdt_conv_pack_weights_batch at the first convolution of the next forward after an optimizer step;
      * fp32 tap-major gradient buffers the weight-gradient kernels overwrite, added into the optimizer's flat (OIHW) gradient
        buffer by nbdt_conv_unpack_wgrad_batch when the gradients are needed (all-reduce / optimizer step / flush()).

    Before: one pack and one unpack launch per layer per step -- 56 launch-bound kernels, 1.3 ms of a WRN-28-10 step
    (profiles/r01_bench_launches_final.md)."""

    def __init__(self):
        self.entries = {}      # (id(param), cin_pad, cout_pad) -> dict
        self.valid = False     # packed copies match the master weights
        self.pending = []      # entries whose gradient buffer holds a not yet unpacked gradient
        self.defer_unpack = True
        self._pack_table = None
        self._unpack_tables = {}

    @staticmethod
    def _table(rows, device):
        return torch.tensor(rows, dtype=torch.int64, device=device)

    def _entry(self, weight, cin_pad, cout_pad):
        key = (id(weight), cin_pad, cout_pad)
        e = self.entries.get(key)
        if e is None:
            cout, cin, k, _ = weight.shape
            dev = weight.device
            e = self.entries[key] = dict(
                key=key, weight=weight, shape=(cout, cin, k), pads=(cout_pad, cin_pad),
                wp=torch.zeros((k * k, cout_pad, cin_pad), dtype=torch.bfloat16, device=dev),
                wpt=torch.zeros((k * k, cin_pad, cout_pad), dtype=torch.bfloat16, device=dev),
                dwp=torch.zeros((k * k, cout_pad, cin_pad), dtype=torch.float32, device=dev))
            self._pack_table = None
            self.valid = False
        return e

    def packed(self, weight, cin_pad, cout_pad):
        e = self._entry(weight, cin_pad, cout_pad)
        if not self.valid:
            self._pack_all(weight.device)
        return e["wp"], e["wpt"]

    def _pack_all(self, device):
        if self._pack_table is None:
            rows, first = [], 0
            for e in self.entries.values():
                cout, cin, k = e["shape"]
                cout_pad, cin_pad = e["pads"]
                rows.append([e["weight"].data_ptr(), e["wp"].data_ptr(), e["wpt"].data_ptr(), cout, cin, cout_pad, cin_pad, (k * k) | (first << 32)])
                first += ((cin_pad + 31) // 32) * ((cout_pad + 31) // 32)
            self._pack_table = (self._table(rows, device), len(rows), first)
        tab, n, tiles = self._pack_table
        with torch.cuda.device(device):
            L.check(L.load().nbdt_conv_pack_weights_batch(_p(tab), n, tiles, L.stream_ptr(device)))
        _count()
        self.valid = True

    def invalidate(self):
        """The master weights changed (optimizer step, load_state_dict): re-pack at the next forward."""
        self.valid = False

    def grad_buffer(self, weight, cin_pad, cout_pad):
        e = self._entry(weight, cin_pad, cout_pad)
        if any(p is e for p in self.pending):
            self.flush()  # a second backward before the first gradient was consumed (gradient accumulation)
        self.pending.append(e)
        return e["dwp"]

    def discard(self):
        self.pending = []

    def flush(self):
        """Add every pending packed gradient into its parameter's .grad (the optimizer's flat buffer)."""
        if not self.pending:
            return
        from .nn_ops import join_wgrad_streams

        dev = self.pending[0]["weight"].device
        join_wgrad_streams(dev)
        key = tuple(e["key"] for e in self.pending)
        cached = self._unpack_tables.get(key)
        if cached is None or any(e["weight"].grad is None or e["weight"].grad.data_ptr() != g for e, g in zip(self.pending, cached[3])):
            rows, first, gptrs = [], 0, []
            for e in self.pending:
                cout, cin, k = e["shape"]
                cout_pad, cin_pad = e["pads"]
                if e["weight"].grad is None:
                    e["weight"].grad = torch.zeros_like(e["weight"])
                gptrs.append(e["weight"].grad.data_ptr())
                rows.append([gptrs[-1], e["dwp"].data_ptr(), 0, cout, cin, cout_pad, cin_pad, (k * k) | (first << 32)])
                first += ((cin + 31) // 32) * ((cout + 31) // 32)
            cached = self._unpack_tables[key] = (self._table(rows, dev), len(rows), first, gptrs)
        tab, n, tiles, _ = cached
        with torch.cuda.device(dev):
            L.check(L.load().nbdt_conv_unpack_wgrad_batch(_p(tab), n, tiles, L.stream_ptr(dev)))
        _count()
        self.pending = []
