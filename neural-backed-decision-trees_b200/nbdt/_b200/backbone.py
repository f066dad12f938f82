"""Functional backbone ops used by nbdt.models: every conv / batch-norm / activation / pool runs as a
hand-written sm_100a kernel (tcgen05 implicit-GEMM convolutions, fused BN(+ReLU,+residual) passes) on
NHWC bf16 activations; the classifier (features x FC weight, fp32) on the small SGEMM of csrc/linear.cu.  CUDA tensors only;
there is no CPU or ATen/cuDNN/cuBLAS path behind these.
"""
import torch

from . import lib as L
from .conv_ops import pad64
from .nn_ops import AvgPoolFunction, BatchNormActFunction, ConvFunction, LinearFunction, StemFunction

# BatchNorm statistics from the convolution epilogue (csrc/conv_gemm.cu: nbdt_conv2d_fwd_stats + nbdt_bn_finalize_partials) instead
# of a separate pass over the convolution output: NBDT_B200_BN_FUSED_STATS=1.  Off by default -- measured on one B200 in the same
# run (profiles/r02_conv_bn_stats_ab.md): the butterfly reduction in the epilogue costs the forward convolutions what the
# statistics pass cost (WRN-28-10 step 19.05 vs 19.03 ms), it only moves the time into the tensor-core kernels.
import os as _os

_FUSED_BN_STATS = _os.environ.get("NBDT_B200_BN_FUSED_STATS", "0") == "1"

BACKBONE_STATUS = "own kernels: tcgen05/TMA implicit-GEMM conv fwd/dgrad/wgrad, fused BN+ReLU(+residual), pool, stem"


def _bn_args(bn):
    training = bn.training or not bn.track_running_stats
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    if bn.momentum is None:  # cumulative moving average (nn.BatchNorm2d semantics): 1 / number of batches seen
        momentum = 1.0 / float(bn.num_batches_tracked) if (training and bn.track_running_stats and bn.num_batches_tracked is not None) else 0.0
    else:
        momentum = bn.momentum
    return bn.weight, bn.bias, bn.running_mean, bn.running_var, training, momentum, bn.eps


def _check_conv(conv_mod):
    k = conv_mod.kernel_size[0]
    s = conv_mod.stride[0]
    if (conv_mod.kernel_size != (k, k) or k not in (1, 3) or conv_mod.stride != (s, s) or s not in (1, 2)
            or conv_mod.padding != (k // 2, k // 2) or conv_mod.bias is not None or conv_mod.groups != 1
            or conv_mod.dilation != (1, 1)):
        raise NotImplementedError(f"no B200 kernel for {conv_mod}; supported: 1x1/3x3, stride 1/2, pad k//2, no bias")
    return s


def conv(x, conv_mod, residual=None, bn_stats=False):
    """NHWC bf16 (N,H,W,Cin) -> (N,H/s,W/s,Cout), optionally + residual (fused in the epilogue).

    bn_stats=True (the output goes into a BatchNorm, possibly in the next module): in training mode the epilogue also
    accumulates the per-channel sum and sum of squares of the output; they travel with the returned tensor and bn_act() uses
    them instead of reading the tensor once more for its batch statistics."""
    stride = _check_conv(conv_mod)
    if bn_stats and conv_mod.training and _FUSED_BN_STATS:
        y, partials = ConvFunction.apply(x, conv_mod.weight, stride, residual, True)
        y._nbdt_bn_partials = partials if partials.numel() else None
        return y
    return ConvFunction.apply(x, conv_mod.weight, stride, residual)


def bn_act(x, bn, relu=True, residual=None, want_skip=False):
    """act(bn(x) (+ residual)).  want_skip=True returns (y, x_skip): use x_skip as the identity branch so that its
    gradient is added inside the BN-backward kernel."""
    return BatchNormActFunction.apply(x, *_bn_args(bn), relu, residual, want_skip, getattr(x, "_nbdt_bn_partials", None))


def conv_bn_act(x, conv_mod, bn, relu=True, residual=None):
    return bn_act(conv(x, conv_mod, bn_stats=True), bn, relu=relu, residual=residual)


def stem(x, conv_mod, bn):
    """NCHW fp32 images -> NHWC bf16 feature map (channels padded to a multiple of 64)."""
    L.require_cuda(x, "backbone stem")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected (N, 3, H, W) images, got {tuple(x.shape)}")
    stride = _check_conv(conv_mod)
    y = StemFunction.apply(x, conv_mod.weight, pad64(conv_mod.out_channels), stride)
    return bn_act(y, bn, relu=True) if bn is not None else y


def global_avg_pool(x):
    return AvgPoolFunction.apply(x)


def avg_pool(x, pool):
    k = getattr(pool, "kernel_size", None)
    if isinstance(pool, torch.nn.AdaptiveAvgPool2d) or k in (x.shape[1], (x.shape[1], x.shape[2])):
        return AvgPoolFunction.apply(x)
    raise NotImplementedError(f"final pool {pool} on a {x.shape[1]}x{x.shape[2]} map is not a global pool")


def linear(feat, lin):
    """Classifier: fp32 features (N, C_phys) x FC weight (reads the first in_features columns; no copy)."""
    return LinearFunction.apply(feat, lin.weight, lin.bias)
