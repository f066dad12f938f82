"""Tensor-level wrappers over the tree-head / tree-loss entry points of libnbdt_b200.so.

PyTorch is plumbing here: it owns the device buffers and the stream; the arithmetic happens
in csrc/tree.cu.  Every function requires CUDA tensors and raises otherwise.
"""
import ctypes

import torch

from . import lib as L

_launches = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def launch_count():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _rows(logits, what):
    L.require_cuda(logits, what)
    if logits.dim() != 2:
        raise ValueError(f"{what}: expected (rows, classes) logits, got shape {tuple(logits.shape)}")
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    if logits.size(0) > 1 and logits.stride(0) < logits.size(1):
        logits = logits.contiguous()
    ld = logits.stride(0) if logits.size(0) > 1 else logits.size(1)
    return logits, ld


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


def _check_classes(dt, logits):
    if logits.size(1) != dt.num_classes:
        raise ValueError(f"logits have {logits.size(1)} classes but the hierarchy has {dt.num_classes} leaves")


def soft_forward(tree, logits):
    """class_probs (fp32, B x C) of the soft decision rules."""
    logits, ld = _rows(logits, "soft rules")
    dt = tree.device_tree(logits.device)
    _check_classes(dt, logits)
    out = torch.empty((logits.size(0), dt.num_classes), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        L.check(L.load().nbdt_soft_forward(dt.handle, _p(logits), L.dtype_code(logits.dtype), logits.size(0), ld,
                                           _p(out), L.stream_ptr(logits.device)))
    _count()
    return out


def soft_backward(tree, logits, grad_probs):
    logits, ld = _rows(logits, "soft rules backward")
    dt = tree.device_tree(logits.device)
    grad_probs = grad_probs.to(torch.float32).contiguous()
    grad = torch.empty_strided(logits.shape, (ld, 1), dtype=logits.dtype, device=logits.device)
    with torch.cuda.device(logits.device):
        L.check(L.load().nbdt_soft_backward(dt.handle, _p(logits), L.dtype_code(logits.dtype), logits.size(0), ld,
                                            _p(grad_probs), 1.0, _p(grad), L.stream_ptr(logits.device)))
    _count()
    return grad


def hard_forward(tree, logits, with_paths=False):
    """pred (int64, B) and, if requested, (path_child, path_inode, path_prob, path_entropy) B x Dmax."""
    logits, ld = _rows(logits, "hard rules")
    dt = tree.device_tree(logits.device)
    _check_classes(dt, logits)
    B, dev = logits.size(0), logits.device
    pred = torch.empty((B,), dtype=torch.int64, device=dev)
    paths = None
    if with_paths:
        D = dt.max_depth
        paths = (
            torch.empty((B, D), dtype=torch.int32, device=dev),
            torch.empty((B, D), dtype=torch.int32, device=dev),
            torch.empty((B, D), dtype=torch.float32, device=dev),
            torch.empty((B, D), dtype=torch.float32, device=dev),
        )
    a = paths or (None, None, None, None)
    with torch.cuda.device(dev):
        L.check(L.load().nbdt_hard_forward(dt.handle, _p(logits), L.dtype_code(logits.dtype), B, ld, _p(pred),
                                           _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), L.stream_ptr(dev)))
    _count()
    return pred, paths


def node_outputs(tree, logits):
    """Packed per-node outputs: logits, probs (B x S fp32), preds (B x N int64), entropy (B x N fp32)."""
    logits, ld = _rows(logits, "node outputs")
    dt = tree.device_tree(logits.device)
    _check_classes(dt, logits)
    B, dev = logits.size(0), logits.device
    nl = torch.empty((B, dt.num_slots), dtype=torch.float32, device=dev)
    pr = torch.empty((B, dt.num_slots), dtype=torch.float32, device=dev)
    preds = torch.empty((B, dt.num_inodes), dtype=torch.int64, device=dev)
    ent = torch.empty((B, dt.num_inodes), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().nbdt_node_outputs(dt.handle, _p(logits), L.dtype_code(logits.dtype), B, ld, _p(nl), _p(pr),
                                           _p(preds), _p(ent), L.stream_ptr(dev)))
    _count()
    return nl, pr, preds, ent, dt


def tree_loss(tree, logits, targets, kind, w_xent, w_tree, tsw=1.0, batch_rows=0, want_grad=True, out=None):
    """Fused loss (+ gradient w.r.t. logits).  Returns (loss fp32 scalar tensor, grad or None).
    `out` = (loss, grad) preallocated outputs (benchmarks: no allocation inside the timed region)."""
    logits, ld = _rows(logits, "tree-supervision loss")
    L.require_cuda(targets, "tree-supervision loss targets")
    dt = tree.device_tree(logits.device)
    _check_classes(dt, logits)
    B, dev = logits.size(0), logits.device
    if targets.shape != (B,):
        raise ValueError(f"targets must have shape ({B},), got {tuple(targets.shape)}")
    targets = targets.to(device=dev, dtype=torch.int64).contiguous()
    if out is not None:
        loss, grad = out
    else:
        loss = torch.empty((), dtype=torch.float32, device=dev)
        grad = torch.empty_strided(logits.shape, (ld, 1), dtype=logits.dtype, device=dev) if want_grad else None
    lib = L.load()
    with torch.cuda.device(dev):
        if kind == "soft":
            rc = lib.nbdt_soft_loss(dt.handle, _p(logits), L.dtype_code(logits.dtype), _p(targets), B, ld,
                                    int(batch_rows), float(w_xent), float(w_tree), 1.0, _p(loss), _p(grad),
                                    L.stream_ptr(dev))
        elif kind == "hard":
            rc = lib.nbdt_hard_loss(dt.handle, _p(logits), L.dtype_code(logits.dtype), _p(targets), B, ld,
                                    int(batch_rows), float(w_xent), float(w_tree), float(tsw), 1.0, _p(loss),
                                    _p(grad), L.stream_ptr(dev))
        else:
            raise ValueError(kind)
    L.check(rc)
    _count(1)
    return loss, grad


class SoftRulesFunction(torch.autograd.Function):
    """Differentiable soft decision rules (for criteria other than the fused cross-entropy)."""

    @staticmethod
    def forward(ctx, logits, tree):
        ctx.tree = tree
        ctx.save_for_backward(logits)
        return soft_forward(tree, logits)

    @staticmethod
    def backward(ctx, grad_probs):
        (logits,) = ctx.saved_tensors
        return soft_backward(ctx.tree, logits, grad_probs).reshape(logits.shape), None


class FusedTreeLossFunction(torch.autograd.Function):
    """loss and dloss/dlogits from ONE kernel launch; backward only rescales the saved gradient."""

    @staticmethod
    def forward(ctx, logits, targets, tree, kind, w_xent, w_tree, tsw, batch_rows):
        want_grad = logits.requires_grad
        loss, grad = tree_loss(tree, logits, targets, kind, w_xent, w_tree, tsw, batch_rows, want_grad)
        ctx.save_for_backward(grad)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return (None,) * 8
        g = grad.reshape(ctx.shape)
        return g * grad_out.to(g.dtype), None, None, None, None, None, None, None
