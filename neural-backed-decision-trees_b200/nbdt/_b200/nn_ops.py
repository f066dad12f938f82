"""Autograd functions of the B200 backbone: every forward/backward is a hand-written kernel behind the
C ABI (csrc/conv_gemm.cu, csrc/bn_ops.cu).  Activations are NHWC bf16 tensors (N, H, W, C); parameters stay
fp32 `nn.Parameter`s in their PyTorch layouts, so state_dicts, optimizers and checkpoints are unchanged."""
import os

import torch

from . import conv_ops as C
from . import lib as L
from .ops import _count, _p


def _stream(t):
    return L.stream_ptr(t.device)


# Weight-gradient kernels are off the critical path of backward (nothing downstream reads dW before the optimizer
# step), and they are tensor-bound while the BatchNorm backward kernels that follow the data-gradient are HBM-bound:
# they run on a side stream and overlap (the wgrad CTAs leave enough shared memory for the BN CTAs to co-reside).
# Only when the gradient goes straight into a FusedSGD flat buffer -- the optimizer step joins the side stream.
# OPT-IN (enable_wgrad_overlap(True) or NBDT_B200_WGRAD_STREAM=1): with it on, parameter gradients are complete only after
# join_wgrad_streams() -- FusedSGD.step() and FlatGradientAllReduce.reduce() call it, code that reads .grad right after
# backward() (gradient clipping, logging) must call it itself.  bench.py turns it on.
_SIDE_STREAMS = {}
# callables told about every convolution weight whose gradient kernel has just been enqueued into the optimizer's flat buffer
# (autograd sees no gradient for those parameters, so no accumulate hook fires): the overlapped all-reduce listens here
GRAD_WRITTEN_CALLBACKS = []
_OVERLAP_WGRAD = os.environ.get("NBDT_B200_WGRAD_STREAM", "0") == "1"


def enable_wgrad_overlap(on=True):
    global _OVERLAP_WGRAD
    _OVERLAP_WGRAD = bool(on)


def wgrad_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return st


def join_wgrad_streams(device=None):
    """Make the current stream wait for every weight-gradient kernel launched so far (optimizer step, all-reduce,
    end of a CUDA-graph capture)."""
    for key, st in _SIDE_STREAMS.items():
        if device is None or torch.device(device).index in (None, key):
            torch.cuda.current_stream(key).wait_stream(st)


def _pad_vec(v, n, fill=0.0):
    v = v.detach().float()
    if v.numel() == n:
        return v.contiguous()
    out = torch.full((n,), fill, dtype=torch.float32, device=v.device)
    out[: v.numel()] = v
    return out


class StemFunction(torch.autograd.Function):
    """3x3 conv (stride 1 or 2, padding 1) on the 3-channel NCHW fp32 input -> NHWC bf16 with `c_phys` physical channels."""

    @staticmethod
    def forward(ctx, x, weight, c_phys, stride=1):
        L.require_cuda(x, "stem conv")
        x = x.float().contiguous()
        w = weight.detach().float().contiguous()
        N, cin, H, W = x.shape
        assert cin == 3 and tuple(w.shape[1:]) == (3, 3, 3), "the stem kernel handles 3x3 convs on 3 input channels"
        cout = w.shape[0]
        y = torch.empty((N, (H + stride - 1) // stride, (W + stride - 1) // stride, c_phys), dtype=torch.bfloat16,
                        device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.load().nbdt_stem_fwd_strided(_p(x), _p(w), _p(y), N, H, W, cout, c_phys, int(stride), _stream(x)))
        _count()
        ctx.save_for_backward(x)
        ctx.cout, ctx.c_phys, ctx.stride = cout, c_phys, int(stride)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        N, _, H, W = x.shape
        dw = torch.empty((ctx.cout, 3, 3, 3), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.load().nbdt_stem_wgrad_strided(_p(x), _p(dy), _p(dw), N, H, W, ctx.cout, ctx.c_phys, ctx.stride,
                                                     _stream(x)))
        _count()
        return None, dw, None, None


class ConvFunction(torch.autograd.Function):
    """y = conv(x, weight) (+ residual) on tcgen05 tensor cores; backward = dgrad + wgrad kernels.

    When the weight's gradient lives in a FusedSGD flat buffer (`weight._nbdt_direct_grad`), the weight-gradient
    unpack kernel accumulates straight into it and autograd gets no weight gradient (one pass over the gradient
    instead of unpack + autograd's accumulate)."""

    @staticmethod
    def forward(ctx, x, weight, stride, residual, bn_stats=False):
        cout, cin, k, _ = weight.shape
        cin_phys = x.shape[3]
        assert cin_phys >= cin
        # output channels the GEMM tiles can cover (multiples of 64 / 128 / 160); others are zero-padded to 64s
        cout_phys = cout if (cout % 64 == 0 or cout % 160 == 0) else C.pad64(cout)
        bank = getattr(weight, "_nbdt_bank", None)
        if bank is not None:  # persistent packed copies, refreshed by one launch per optimizer step
            wp, wpt = bank.packed(weight, cin_phys, cout_phys)
        else:
            wp, wpt = C.pack_weights(weight, cin_pad=cin_phys, cout_pad=cout_phys)
        partials = None
        if bn_stats:  # second output: per-CTA partial sums of y for the BatchNorm that consumes it (not differentiable)
            y, partials = C.conv_fwd(x, wp, k, stride, residual, real=(cout, cin), bn_stats=True)
        else:
            y = C.conv_fwd(x, wp, k, stride, residual, real=(cout, cin))
        ctx.bn_stats = bn_stats
        ctx.save_for_backward(x, wpt)
        ctx.meta = (tuple(weight.shape), k, stride, residual is not None)
        ctx.direct_grad = weight.grad if getattr(weight, "_nbdt_direct_grad", False) else None
        ctx.param = weight if ctx.direct_grad is not None else None
        ctx.bank = bank if (bank is not None and ctx.direct_grad is not None and bank.defer_unpack) else None
        ctx.pads = (cin_phys, cout_phys)
        if bn_stats:
            if partials is None:
                partials = torch.empty((0,), dtype=torch.float32, device=x.device)
            ctx.mark_non_differentiable(partials)
            return y, partials
        return y

    @staticmethod
    def backward(ctx, dy, _dpartials=None):
        x, wpt = ctx.saved_tensors
        wshape, k, stride, has_res = ctx.meta
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = C.conv_dgrad(dy, wpt, (x.shape[1], x.shape[2]), k, stride, real=wshape[:2])
        if ctx.needs_input_grad[1]:
            if ctx.bank is not None:
                # the gradient kernel overwrites the layer's persistent tap-major buffer; ConvWeightBank.flush() adds all of them
                # into the flat gradient buffer with one launch when they are needed (all-reduce / optimizer step)
                dwp = ctx.bank.grad_buffer(ctx.param, *ctx.pads)
                if _OVERLAP_WGRAD:
                    main, side = torch.cuda.current_stream(x.device), wgrad_stream(x.device)
                    side.wait_stream(main)  # x and dy are complete on the main stream at this point
                    with torch.cuda.stream(side):
                        C.conv_wgrad(x, dy, k, stride, real=wshape[:2], out=dwp)
                    x.record_stream(side)
                    dy.record_stream(side)
                else:
                    C.conv_wgrad(x, dy, k, stride, real=wshape[:2], out=dwp)
            elif ctx.direct_grad is not None and _OVERLAP_WGRAD:
                main, side = torch.cuda.current_stream(x.device), wgrad_stream(x.device)
                side.wait_stream(main)  # x and dy are complete on the main stream at this point
                with torch.cuda.stream(side):
                    dwp = C.conv_wgrad(x, dy, k, stride, real=wshape[:2])
                    C.unpack_wgrad(dwp, wshape, out=ctx.direct_grad, accumulate=True)
                x.record_stream(side)
                dy.record_stream(side)
            elif ctx.direct_grad is not None:
                C.unpack_wgrad(C.conv_wgrad(x, dy, k, stride, real=wshape[:2]), wshape, out=ctx.direct_grad, accumulate=True)
            else:
                dw = C.unpack_wgrad(C.conv_wgrad(x, dy, k, stride, real=wshape[:2]), wshape)
            if ctx.param is not None:
                for cb in GRAD_WRITTEN_CALLBACKS:
                    cb(ctx.param)
        return dx, dw, None, (dy if has_res else None), None


# NBDT_B200_BN_FOLD_FINALIZE=1: BatchNorm finalize folded into the consumer kernels (csrc/bn_ops.cu); measured slower, off by default
_FOLD_BN_FINALIZE = os.environ.get("NBDT_B200_BN_FOLD_FINALIZE") is not None


class BatchNormActFunction(torch.autograd.Function):
    """y = act(batch_norm(x) (+ residual)); training mode uses batch statistics and updates the running ones.

    `want_skip`: also return x itself as a second output (the identity branch of a pre-activation residual unit):
    its gradient then arrives in backward() and is added inside the BN-backward kernel instead of by a separate
    autograd accumulation pass over the tensor."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, residual, want_skip=False,
                partials=None):
        C._check_act(x, "batch norm input")
        N, H, W, Cp = x.shape
        M = N * H * W
        c_real = gamma.numel()
        dev = x.device
        g, b = _pad_vec(gamma, Cp, 1.0), _pad_vec(beta, Cp)
        scale_shift = torch.empty((2, Cp), dtype=torch.float32, device=dev)
        mean_invstd = torch.empty((2, Cp), dtype=torch.float32, device=dev)
        lib = L.load()
        act = int(relu)  # activation code: 0 none, 1 ReLU, 2 SiLU
        if act == 2 and residual is not None:
            raise NotImplementedError("SiLU after a residual add is not implemented")
        y = torch.empty_like(x)
        with torch.cuda.device(dev):
            rm = running_mean if running_mean is not None else None
            rv = running_var if running_var is not None else None
            if training and partials is not None and partials.numel() > 0:
                # the producing convolution already summed x and x^2 per channel in its epilogue (csrc/conv_gemm.cu)
                assert partials.shape[1:] == (2, Cp), (partials.shape, Cp)
                L.check(lib.nbdt_bn_finalize_partials(_p(partials), partials.shape[0], M, Cp, c_real, _p(g), _p(b), _p(rm), _p(rv),
                                                      float(momentum), float(eps), _p(mean_invstd), _p(scale_shift), _stream(x)))
                L.check(lib.nbdt_bn_apply(_p(x), _p(residual), _p(y), M, Cp, _p(scale_shift), act, _stream(x)))
                _count(2)
            elif training and not _FOLD_BN_FINALIZE:
                sums = torch.empty((lib.nbdt_bn_scratch_floats(Cp),), dtype=torch.float32, device=dev)
                L.check(lib.nbdt_bn_stats(_p(x), M, Cp, c_real, _p(g), _p(b), _p(rm), _p(rv), float(momentum), float(eps),
                                          _p(sums), _p(mean_invstd), _p(scale_shift), _stream(x)))
                L.check(lib.nbdt_bn_apply(_p(x), _p(residual), _p(y), M, Cp, _p(scale_shift), act, _stream(x)))
                _count(3)
            elif training:
                # experiment (measured slower): statistics kernel -> fp64 accumulator, the apply kernel finalizes in its prologue
                sums = torch.empty((lib.nbdt_bn_scratch_floats(Cp),), dtype=torch.float32, device=dev)
                L.check(lib.nbdt_bn_stats_apply(_p(x), _p(residual), _p(y), M, Cp, c_real, _p(g), _p(b), _p(rm), _p(rv),
                                                float(momentum), float(eps), _p(sums), _p(mean_invstd), _p(scale_shift), act,
                                                _stream(x)))
                _count(2)
            else:
                L.check(lib.nbdt_bn_eval_scale(Cp, c_real, _p(g), _p(b), _p(running_mean), _p(running_var), float(eps),
                                               _p(scale_shift), _stream(x)))
                L.check(lib.nbdt_bn_apply(_p(x), _p(residual), _p(y), M, Cp, _p(scale_shift), act, _stream(x)))
                _count(2)
        ctx.training = training
        ctx.meta = (M, Cp, c_real, act, residual is not None)
        # ReLU mask in backward: recomputed from x and scale/shift (bit-identical fmaf) unless a residual was added
        # before the activation, in which case the saved output is needed; SiLU's gradient is always recomputed
        mask_from_act = act == 1 and residual is not None
        ctx.save_for_backward(x, y if mask_from_act else None, scale_shift if (act and not mask_from_act) else None,
                              g, mean_invstd)
        ctx.set_materialize_grads(False)
        if want_skip:
            return y, x
        return y

    @staticmethod
    def backward(ctx, dy, dskip=None):
        if not ctx.training:
            raise NotImplementedError("backward through eval-mode batch norm is not implemented on the B200 path")
        x, act, scale_shift, g, mean_invstd = ctx.saved_tensors
        M, Cp, c_real, relu, has_res = ctx.meta
        if dy is None:
            return (dskip,) + (None,) * 11
        dy = dy.contiguous()
        extra = dskip.contiguous() if dskip is not None else None
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        dsums = torch.empty((L.load().nbdt_bn_scratch_floats(Cp),), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            if relu == 2:
                L.check(L.load().nbdt_bn_backward_silu(_p(x), _p(scale_shift), _p(dy), _p(extra), _p(dx), M, Cp, c_real,
                                                       _p(g), _p(mean_invstd), _p(dsums), _stream(x)))
            else:
                L.check(L.load().nbdt_bn_backward(_p(x), _p(act), _p(scale_shift), _p(dy), _p(extra), _p(dx), _p(dres), M,
                                                  Cp, c_real, _p(g), _p(mean_invstd), _p(dsums), _stream(x)))
        _count(2 if _FOLD_BN_FINALIZE else 3)
        return dx, dsums[Cp:Cp + c_real], dsums[:c_real], None, None, None, None, None, None, dres, None, None


class AvgPoolFunction(torch.autograd.Function):
    """(N, H, W, C) bf16 -> (N, C) fp32 mean over the spatial positions."""

    @staticmethod
    def forward(ctx, a):
        C._check_act(a, "average pool input")
        N, H, W, Cp = a.shape
        feat = torch.empty((N, Cp), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            L.check(L.load().nbdt_avgpool_fwd(_p(a), _p(feat), N, H * W, Cp, _stream(a)))
        _count()
        ctx.shape = (N, H, W, Cp)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        N, H, W, Cp = ctx.shape
        dfeat = dfeat.float().contiguous()
        da = torch.empty((N, H, W, Cp), dtype=torch.bfloat16, device=dfeat.device)
        with torch.cuda.device(dfeat.device):
            L.check(L.load().nbdt_avgpool_bwd(_p(dfeat), _p(da), N, H * W, Cp, _stream(dfeat)))
        _count()
        return da


class LinearFunction(torch.autograd.Function):
    """y = x W^T (+ b) in fp32 on this repository's SGEMM kernel (csrc/linear.cu): the classifier head and the squeeze-excitation
    layers.  x: (N, D_phys >= D) -- only the first D columns are read (padding channels of the pooled features); W: (C, D)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        L.require_cuda(x, "linear input")
        x = x.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        w = weight.float().contiguous() if (weight.dtype != torch.float32 or not weight.is_contiguous()) else weight
        N, (C, D) = x.shape[0], w.shape
        y = torch.empty((N, C), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.load().nbdt_sgemm(N, C, D, _p(x), x.stride(0), 1, _p(w), 1, D, _p(y), C, _p(bias), 0, _stream(x)))
        _count()
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.float().contiguous()
        N, (C, D) = x.shape[0], w.shape
        dx = dw = db = None
        lib = L.load()
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                dx = torch.zeros_like(x) if x.shape[1] != D else torch.empty_like(x)
                L.check(lib.nbdt_sgemm(N, D, C, _p(dy), C, 1, _p(w), D, 1, _p(dx), dx.stride(0), None, 0, _stream(x)))
                _count()
            if ctx.needs_input_grad[1]:
                dw = torch.empty_like(w)
                L.check(lib.nbdt_sgemm(C, D, N, _p(dy), 1, C, _p(x), x.stride(0), 1, _p(dw), D, None, 0, _stream(x)))
                _count()
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = torch.empty((C,), dtype=torch.float32, device=x.device)
                L.check(lib.nbdt_colsum(_p(dy), N, C, C, _p(db), _stream(x)))
                _count()
        return dx, dw, db
