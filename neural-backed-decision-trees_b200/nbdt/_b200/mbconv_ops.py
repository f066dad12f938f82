"""Inference-path ops of the MBConv (EfficientNet) backbones: strided stem, depthwise conv with the folded
BatchNorm + SiLU fused, squeeze-excitation scaling, eval-mode BatchNorm + activation, 1x1 convolutions on the tcgen05
implicit-GEMM kernel.  NHWC bf16 activations with the channel count padded to a multiple of 64 (padding channels
are exactly zero everywhere: zero weights, zero BN scale/shift, SiLU(0) = 0).

The functions in the first half are the inference path (folded BatchNorm, cached packed weights): calling them with
autograd enabled on parameters that require grad raises instead of silently detaching.  `DepthwiseConvFunction` and
`SEScaleFunction` are the autograd functions of the training path (depthwise dgrad / wgrad, squeeze-excitation gradient
kernels); the 1x1 convolutions and BatchNorm + SiLU of the training path are nn_ops.ConvFunction / BatchNormActFunction."""
import torch

from . import conv_ops as C
from . import lib as L
from .ops import _count, _p

ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2


def _stream(t):
    return L.stream_ptr(t.device)


def require_inference(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "the MBConv (EfficientNet) kernels of nbdt-b200 are inference-only so far: wrap the call in "
            "torch.no_grad() (training kernels: depthwise dgrad/wgrad, SiLU/SE backward are not built)")


def _versions(tensors):
    return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors if t is not None)


def cached(module, slot, tensors, build):
    """Derived inference constants (packed weights, folded BN) cached on the module, rebuilt when a source tensor
    is replaced or modified in place."""
    store = module.__dict__.setdefault("_nbdt_b200_cache", {})
    key = _versions(tensors)
    hit = store.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    value = build()
    store[slot] = (key, value)
    return value


def _pad_vec(v, n, fill=0.0):
    v = v.detach().float()
    if v.numel() == n:
        return v.contiguous()
    out = torch.full((n,), fill, dtype=torch.float32, device=v.device)
    out[: v.numel()] = v
    return out


def folded_scale_shift(bn, c_phys):
    """[2][c_phys] fp32: y = x*scale + shift of an eval-mode BatchNorm2d (padding channels: 0, 0)."""
    def build():
        dev = bn.weight.device
        ss = torch.empty((2, c_phys), dtype=torch.float32, device=dev)
        # keep the padded copies alive across the launch (a temporary's storage would be recycled for the next one)
        g, b = _pad_vec(bn.weight, c_phys, 1.0), _pad_vec(bn.bias, c_phys)
        rm, rv = _pad_vec(bn.running_mean, c_phys), _pad_vec(bn.running_var, c_phys, 1.0)
        with torch.cuda.device(dev):
            L.check(L.load().nbdt_bn_eval_scale(c_phys, bn.num_features, _p(g), _p(b), _p(rm), _p(rv), float(bn.eps),
                                                _p(ss), L.stream_ptr(dev)))
        _count()
        return ss
    return cached(bn, f"ss{c_phys}", (bn.weight, bn.bias, bn.running_mean, bn.running_var), build)


def bn_act_eval(x, bn, act=ACT_NONE, residual=None):
    """act(bn_eval(x) (+ residual)) in one pass over x."""
    C._check_act(x, "batch norm input")
    require_inference(bn.weight, bn.bias, x)
    N, H, W, Cp = x.shape
    ss = folded_scale_shift(bn, Cp)
    y = torch.empty_like(x)
    if residual is not None:
        C._check_act(residual, "batch norm residual")
        assert residual.shape == x.shape
    with torch.cuda.device(x.device):
        L.check(L.load().nbdt_bn_apply(_p(x), _p(residual), _p(y), N * H * W, Cp, _p(ss), int(act), _stream(x)))
    _count()
    return y


def stem(x, conv_mod, c_phys):
    """3x3 conv (stride 1 or 2, padding 1) on NCHW fp32 images -> NHWC bf16 with c_phys physical channels."""
    L.require_cuda(x, "stem conv")
    require_inference(conv_mod.weight, x)
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected (N, 3, H, W) images, got {tuple(x.shape)}")
    if conv_mod.kernel_size != (3, 3) or conv_mod.padding != (1, 1) or conv_mod.stride not in ((1, 1), (2, 2)) \
            or conv_mod.bias is not None or conv_mod.groups != 1:
        raise NotImplementedError(f"no B200 stem kernel for {conv_mod}")
    s = conv_mod.stride[0]
    x = x.float().contiguous()
    w = conv_mod.weight.detach().float().contiguous()
    N, _, H, W = x.shape
    y = torch.empty((N, (H + s - 1) // s, (W + s - 1) // s, c_phys), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().nbdt_stem_fwd_strided(_p(x), _p(w), _p(y), N, H, W, w.shape[0], c_phys, s, _stream(x)))
    _count()
    return y


def conv1x1(x, conv_mod):
    """Pointwise conv on the tcgen05 implicit-GEMM kernel; output channels padded to a multiple of 64."""
    C._check_act(x, "1x1 conv input")
    require_inference(conv_mod.weight, x)
    if conv_mod.kernel_size != (1, 1) or conv_mod.stride != (1, 1) or conv_mod.bias is not None or conv_mod.groups != 1:
        raise NotImplementedError(f"no B200 pointwise kernel for {conv_mod}")
    cin_phys = x.shape[3]
    cout_phys = C.pad64(conv_mod.out_channels)
    wp = cached(conv_mod, f"wp{cin_phys}", (conv_mod.weight,),
                lambda: C.pack_weights(conv_mod.weight, cin_pad=cin_phys, cout_pad=cout_phys, want=(True, False))[0])
    return C.conv_fwd(x, wp, 1, 1)


def dwconv_bn_act(x, conv_mod, bn, act=ACT_SILU):
    """act(bn_eval(depthwise_conv(x))) in ONE kernel (the BatchNorm is a per-channel affine of the conv output)."""
    C._check_act(x, "depthwise conv input")
    require_inference(conv_mod.weight, bn.weight, x)
    k, s = conv_mod.kernel_size[0], conv_mod.stride[0]
    ch = conv_mod.in_channels
    if (conv_mod.kernel_size != (k, k) or k not in (3, 5) or conv_mod.stride != (s, s) or s not in (1, 2)
            or conv_mod.padding != (k // 2, k // 2) or conv_mod.groups != ch or conv_mod.out_channels != ch
            or conv_mod.bias is not None):
        raise NotImplementedError(f"no B200 depthwise kernel for {conv_mod}")
    N, H, W, Cp = x.shape

    def pack():  # [C,1,k,k] -> tap-major [k*k][C_phys] fp32, zero for the padding channels
        w = torch.zeros((k * k, Cp), dtype=torch.float32, device=x.device)
        w[:, :ch] = conv_mod.weight.detach().float().reshape(ch, k * k).t()
        return w

    w = cached(conv_mod, f"dw{Cp}", (conv_mod.weight,), pack)
    ss = folded_scale_shift(bn, Cp)
    y = torch.empty((N, (H + s - 1) // s, (W + s - 1) // s, Cp), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().nbdt_dwconv_fwd(_p(x), _p(w), _p(ss), _p(y), N, H, W, Cp, k, s, int(act), _stream(x)))
    _count()
    return y


def se_scale(x, s):
    """x[n,h,w,c] * s[n,c]  (s fp32 (N, C_phys))."""
    C._check_act(x, "squeeze-excitation input")
    N, H, W, Cp = x.shape
    s = s.float().contiguous()
    assert s.shape == (N, Cp)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.load().nbdt_se_scale(_p(x), _p(s), _p(y), N, H * W, Cp, _stream(x)))
    _count()
    return y


def _pack_dw(weight, c_phys):
    """[C,1,k,k] -> tap-major [k*k][c_phys] fp32, zero for the padding channels."""
    ch, _, k, _ = weight.shape
    w = torch.zeros((k * k, c_phys), dtype=torch.float32, device=weight.device)
    w[:, :ch] = weight.detach().float().reshape(ch, k * k).t()
    return w


class DepthwiseConvFunction(torch.autograd.Function):
    """Depthwise k x k convolution (k in {3, 5}, stride 1 or 2, padding k//2) on NHWC bf16 activations."""

    @staticmethod
    def forward(ctx, x, weight, stride):
        C._check_act(x, "depthwise conv input")
        N, H, W, Cp = x.shape
        ch, _, k, _ = weight.shape
        assert Cp >= ch
        w = _pack_dw(weight, Cp)
        y = torch.empty((N, (H + stride - 1) // stride, (W + stride - 1) // stride, Cp), dtype=torch.bfloat16, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.load().nbdt_dwconv_fwd(_p(x), _p(w), None, _p(y), N, H, W, Cp, k, int(stride), ACT_NONE, _stream(x)))
        _count()
        ctx.save_for_backward(x, w)
        ctx.meta = (tuple(weight.shape), int(stride))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        wshape, stride = ctx.meta
        ch, _, k, _ = wshape
        N, H, W, Cp = x.shape
        dy = dy.contiguous()
        dx = dw = None
        lib = L.load()
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                L.check(lib.nbdt_dwconv_dgrad(_p(dy), _p(w), _p(dx), N, H, W, Cp, k, stride, _stream(x)))
                _count()
            if ctx.needs_input_grad[1]:
                dwp = torch.empty((k * k, Cp), dtype=torch.float32, device=x.device)
                L.check(lib.nbdt_dwconv_wgrad(_p(x), _p(dy), _p(dwp), N, H, W, Cp, k, stride, _stream(x)))
                _count(2)
                dw = dwp[:, :ch].t().reshape(wshape)
        return dx, dw, None


class SEScaleFunction(torch.autograd.Function):
    """y[n,h,w,c] = x[n,h,w,c] * s[n,c]  (s fp32 (N, C_phys)); also used for stochastic depth (s = mask / keep_prob)."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.save_for_backward(x, s)
        return se_scale(x, s)

    @staticmethod
    def backward(ctx, dy):
        x, s = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ds = None
        if ctx.needs_input_grad[0]:
            dx = se_scale(dy, s)
        if ctx.needs_input_grad[1]:
            N, H, W, Cp = x.shape
            ds = torch.empty((N, Cp), dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                L.check(L.load().nbdt_se_dscale(_p(x), _p(dy), _p(ds), N, H * W, Cp, _stream(x)))
            _count(2)
        return dx, ds
