"""Functional backbone ops used by nbdt.models (conv / batch-norm / activation / pool / linear).

INTERIM (round 1, first slice): these dispatch to ATen/cuDNN on the CUDA device while the
hand-written tcgen05 implicit-GEMM conv + BN kernels (csrc/conv_*.cu) are brought up; each op is
replaced behind this interface with parity tests against the ATen result.  CUDA tensors only.
"""
import torch
import torch.nn.functional as F

from . import lib as L


def _bn(x, bn):
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                        bn.training or not bn.track_running_stats, bn.momentum or 0.0, bn.eps)


def conv(x, conv_mod, residual=None):
    L.require_cuda(x, "backbone conv")
    y = F.conv2d(x, conv_mod.weight.to(x.dtype), None, conv_mod.stride, conv_mod.padding)
    return y if residual is None else y + residual


def bn_act(x, bn, relu=True):
    L.require_cuda(x, "backbone batch-norm")
    y = _bn(x, bn)
    return F.relu(y) if relu else y


def conv_bn_act(x, conv_mod, bn, relu=True, residual=None):
    y = _bn(conv(x, conv_mod), bn)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def stem(x, conv_mod, bn):
    L.require_cuda(x, "backbone stem")
    return conv_bn_act(x, conv_mod, bn, relu=True) if bn is not None else conv(x, conv_mod)


def global_avg_pool(x):
    return x.mean(dim=(2, 3))


def avg_pool(x, pool):
    return pool(x).flatten(1)


def linear(x, lin):
    return F.linear(x, lin.weight.to(x.dtype), lin.bias.to(x.dtype))
