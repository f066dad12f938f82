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


def _node_tables(dt, device):
    """(M, seg) for the backward of node_outputs, cached on the device tree: M (S x C fp32) maps a gradient on the packed node
    logits to the class logits -- M[e, c] = 1/#leaves(e) for the classes c under child edge e (a node logit is the MEAN of the
    class logits under the child, reference nbdt/model.py:83-97) -- and seg (S int64) is the inner node every edge belongs to."""
    cache = dt.__dict__.setdefault("_node_tables", {})
    key = str(device)
    if key not in cache:
        import numpy as np

        C, N = dt.num_classes, dt.num_inodes
        ptr, node = dt.child_ptr, dt.child_node
        under = {}
        stack = [int(dt.root) + C]
        while stack:  # iterative post-order: hierarchies may be deep chains
            n = stack[-1]
            kids = [int(k) for k in node[ptr[n - C]:ptr[n - C + 1]]]
            todo = [k for k in kids if k >= C and k not in under]
            if todo:
                stack.extend(todo)
                continue
            under[n] = [c for k in kids for c in ([k] if k < C else under[k])]
            stack.pop()
        S = int(ptr[N])
        M = np.zeros((S, C), dtype=np.float32)
        seg = np.zeros((S,), dtype=np.int64)
        for i in range(N):
            for e in range(int(ptr[i]), int(ptr[i + 1])):
                k = int(node[e])
                leaves = [k] if k < C else under[k]
                M[e, leaves] = 1.0 / len(leaves)
                seg[e] = i
        cache[key] = (torch.from_numpy(M).to(device), torch.from_numpy(seg).to(device))
    return cache[key]


class NodeOutputsFunction(torch.autograd.Function):
    """Differentiable per-node outputs (node logits, child probabilities, entropy; the predictions are integers).  The reference
    builds them from autograd ops (nbdt/model.py:83-123), so criteria or analyses written on top of `get_node_logits`,
    `get_all_node_outputs` / `forward_nodes` get gradients there; here the forward is the fused kernel and the backward is the
    softmax / entropy chain rule on the saved probabilities followed by ONE product with the edge-to-class mean matrix
    (csrc/linear.cu).  Not a hot path: the training losses use the fused loss kernels."""

    @staticmethod
    def forward(ctx, logits, tree):
        nl, pr, preds, ent, _ = node_outputs(tree, logits)
        ctx.tree, ctx.shape, ctx.dtype = tree, logits.shape, logits.dtype
        ctx.save_for_backward(pr)
        ctx.mark_non_differentiable(preds)
        return nl, pr, preds, ent

    @staticmethod
    def backward(ctx, g_nl, g_pr, _g_preds, g_ent):
        (pr,) = ctx.saved_tensors
        dev = pr.device
        dt = ctx.tree.device_tree(dev)
        M, seg = _node_tables(dt, dev)
        B, S = pr.shape
        g = torch.zeros_like(pr) if g_nl is None else g_nl.float().clone()
        t = None if g_pr is None else g_pr.float()
        if g_ent is not None:  # entropy = -sum p ln p  ->  d/dp = -(ln p + 1)
            te = g_ent.float()[:, seg] * (-(torch.log(pr.clamp_min(1e-30)) + 1.0))
            t = te if t is None else t + te
        if t is not None:      # softmax over the children of every node
            dot = torch.zeros((B, dt.num_inodes), dtype=torch.float32, device=dev).index_add_(1, seg, t * pr)
            g += pr * (t - dot[:, seg])
        g = g.contiguous()
        C = dt.num_classes
        grad = torch.empty((B, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().nbdt_sgemm(B, C, S, _p(g), S, 1, _p(M), C, 1, _p(grad), C, None, 0, L.stream_ptr(dev)))
        _count()
        return grad.to(ctx.dtype).reshape(ctx.shape), None


def node_outputs_autograd(tree, logits):
    """node_outputs() that records an autograd graph when the logits ask for one: (logits, probs, preds, entropy, device tree)."""
    if logits.requires_grad and torch.is_grad_enabled():
        nl, pr, preds, ent = NodeOutputsFunction.apply(logits, tree)
        return nl, pr, preds, ent, tree.device_tree(logits.device)
    return node_outputs(tree, logits)


# ---- class-major (N, C, H, W) tensors: the segmentation wrappers without coerce_tensor / uncoerce_tensor copies ----
def _nchw(logits, what):
    L.require_cuda(logits, what)
    if logits.dim() != 4:
        raise ValueError(f"{what}: expected (N, C, H, W) logits, got shape {tuple(logits.shape)}")
    return logits if logits.is_contiguous() else logits.contiguous()


def soft_forward_nchw(tree, logits):
    """Class probabilities (N, C, H, W) fp32 of the soft rules, one row of the hierarchy per pixel (reference model.py:376-399)."""
    logits = _nchw(logits, "soft rules (segmentation)")
    dt = tree.device_tree(logits.device)
    N, C, H, W = logits.shape
    if C != dt.num_classes:
        raise ValueError(f"logits have {C} classes, the hierarchy has {dt.num_classes}")
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        L.check(L.load().nbdt_soft_forward_nchw(dt.handle, _p(logits), L.dtype_code(logits.dtype), N, H * W, _p(out),
                                                L.stream_ptr(logits.device)))
    _count()
    return out


def hard_forward_nchw(tree, logits):
    """Predicted class per pixel (N, H, W) int64 of the hard rules."""
    logits = _nchw(logits, "hard rules (segmentation)")
    dt = tree.device_tree(logits.device)
    N, C, H, W = logits.shape
    if C != dt.num_classes:
        raise ValueError(f"logits have {C} classes, the hierarchy has {dt.num_classes}")
    pred = torch.empty((N, H, W), dtype=torch.int64, device=logits.device)
    with torch.cuda.device(logits.device):
        L.check(L.load().nbdt_hard_forward_nchw(dt.handle, _p(logits), L.dtype_code(logits.dtype), N, H * W, _p(pred),
                                                L.stream_ptr(logits.device)))
    _count()
    return pred


def tree_loss_nchw(tree, logits, targets, w_xent, w_tree, batch_rows=0, want_grad=True):
    """Fused soft tree-supervision loss over the pixels of (N, C, H, W) logits with (N, H, W) targets: (loss, gradient in the
    logits' layout and dtype or None).  Reference: SoftSegTreeSupLoss.forward, nbdt/loss.py:318-327."""
    logits = _nchw(logits, "tree-supervision loss (segmentation)")
    L.require_cuda(targets, "tree-supervision loss targets")
    dt = tree.device_tree(logits.device)
    N, C, H, W = logits.shape
    if C != dt.num_classes:
        raise ValueError(f"logits have {C} classes, the hierarchy has {dt.num_classes}")
    if tuple(targets.shape) != (N, H, W):
        raise ValueError(f"targets must have shape {(N, H, W)}, got {tuple(targets.shape)}")
    dev = logits.device
    targets = targets.to(device=dev, dtype=torch.int64).contiguous()
    loss = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty_like(logits) if want_grad else None
    with torch.cuda.device(dev):
        L.check(L.load().nbdt_soft_loss_nchw(dt.handle, _p(logits), L.dtype_code(logits.dtype), _p(targets), N, H * W,
                                             int(batch_rows), float(w_xent), float(w_tree), 1.0, _p(loss), _p(grad),
                                             L.stream_ptr(dev)))
    _count(2)
    return loss, grad


class FusedSegTreeLossFunction(torch.autograd.Function):
    """w_xent * CE(logits) + w_tree * CE(class probabilities of the soft rules) per pixel, loss and gradient from one launch."""

    @staticmethod
    def forward(ctx, logits, targets, tree, w_xent, w_tree, batch_rows):
        loss, grad = tree_loss_nchw(tree, logits, targets, w_xent, w_tree, batch_rows, logits.requires_grad)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return (None,) * 6
        return grad * grad_out.to(grad.dtype), None, None, None, None, None


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
