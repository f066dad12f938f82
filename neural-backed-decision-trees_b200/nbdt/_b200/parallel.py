"""Data parallelism for the train step: one process per GPU, parameters replicated, the batch split
by rows, ONE collective per step -- an all-reduce(sum) of the backbone + FC gradients in flat
buckets (NCCL over NVLink; gloo on CPU for the host-logic tests).  Replaces the reference's
single-process `torch.nn.DataParallel` (main.py:160-162), which re-broadcasts every weight and
gathers the logits to GPU 0 each step.  The tree loss needs no collective: each rank evaluates its
rows with `batch_rows` = global batch, so per-rank gradients are already scaled for a plain sum.
BatchNorm statistics stay per replica, as in the reference.
"""
import torch
import torch.distributed as dist

from .backbone import BACKBONE_STATUS  # noqa: E402,F401


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s parameters and buffers."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


class GradientAllReduce:
    """Flat-bucket gradient all-reduce.  `reduce()` after backward sums `.grad` over ranks.

    Buckets are filled in reverse parameter order (the order backward produces gradients) and capped
    at `bucket_bytes`, so a bucket's all-reduce can be issued as soon as its last gradient exists."""

    def __init__(self, params, world_size, bucket_bytes=64 << 20):
        self.world_size = world_size
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)

    def reduce(self):
        if self.world_size == 1:
            return
        handles = []
        for i, bucket in enumerate(self.buckets):
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket))
        for work, flat, bucket in handles:
            work.wait()
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = flat[off:off + n].view_as(p).clone()
                else:
                    p.grad.copy_(flat[off:off + n].view_as(p))
                off += n


class FlatGradientAllReduce:
    """All-reduce(sum) of FusedSGD's flat gradient buffer: one NCCL call per step, no copies."""

    def __init__(self, optimizer, world_size):
        self.optimizer = optimizer
        self.world_size = world_size

    def reduce(self):
        if self.world_size > 1:
            from .nn_ops import join_wgrad_streams

            join_wgrad_streams(self.optimizer.flat_grad.device)  # weight gradients are written on a side stream
            dist.all_reduce(self.optimizer.flat_grad, op=dist.ReduceOp.SUM)


def shard_rows(total_rows, rank, world_size):
    """[begin, end) of this rank's rows of a global batch (remainder rows go to the low ranks)."""
    base, rem = divmod(total_rows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
