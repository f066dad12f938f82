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

            self.optimizer.reseat_grads()  # a .grad that autograd replaced (zero_grad(set_to_none=True)) goes back into the flat buffer first
            join_wgrad_streams(self.optimizer.flat_grad.device)  # weight gradients are written on a side stream
            dist.all_reduce(self.optimizer.flat_grad, op=dist.ReduceOp.SUM)


class OverlappedGradientAllReduce:
    """The same sum, issued in a few contiguous pieces of the flat gradient buffer WHILE backward is still running.

    Parameters sit in the flat buffer in definition order and backward finishes them back to front, so at any moment a tail of the
    buffer is final.  A post-accumulate hook on a BatchNorm bias marks the moment everything behind it is complete: the piece
    [offset after that bias, previous boundary) goes to NCCL on a communication stream that waits for the main stream (and the
    weight-gradient side stream) at that point; `finish()` after backward sends the head piece and joins.  For WRN-28-10 the last
    stage holds 71 % of the gradient bytes and is in flight before the first two stages have started their backward.
    Everything is stream-ordered (events, no host sync), so it can be captured in a CUDA graph together with forward + backward.
    Needs FusedSGD (flat buffer) and its zero_grad (which keeps every .grad a view of that buffer)."""

    def __init__(self, optimizer, world_size, pieces=4, bf16=False):
        self.optimizer = optimizer
        self.world_size = world_size
        self.bf16 = bf16
        self.stream = None
        self.pending = []
        self.bounds = []       # element offsets of the piece boundaries, descending; the head piece [0, bounds[-1]) goes last
        self._hooks = []
        self._upper = None
        params = optimizer.param_groups[0]["params"]
        total = optimizer.flat_grad.numel()
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += p.numel()
        if world_size > 1 and pieces > 1:
            # candidate triggers: 1-D parameters directly followed by another parameter (BatchNorm bias: its hook fires after the layers
            # behind it have finished); one per `total / pieces` elements, scanning from the back
            want = [total * k // pieces for k in range(pieces - 1, 0, -1)]
            cands = [i for i, p in enumerate(params) if p.dim() == 1 and i + 1 < len(params) and i > 0 and params[i - 1].dim() == 1]
            for w in want:
                best = min(cands, key=lambda i: abs(offs[i + 1] - w), default=None)
                if best is not None and offs[best + 1] not in self.bounds and 0 < offs[best + 1] < total:
                    self.bounds.append(offs[best + 1])
                    self._hooks.append(params[best].register_post_accumulate_grad_hook(self._make_hook(offs[best + 1])))
            self.bounds.sort(reverse=True)

    def _make_hook(self, bound):
        def hook(param):
            self._issue(bound)

        return hook

    def _issue(self, lower):
        """all-reduce flat_grad[lower, upper) where upper is the previous boundary; ordered after everything enqueued so far."""
        if self.world_size == 1:
            return
        from .nn_ops import _SIDE_STREAMS

        flat = self.optimizer.flat_grad
        upper = flat.numel() if self._upper is None else self._upper
        if lower >= upper:
            return
        self._upper = lower
        dev = flat.device
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        self.stream.wait_stream(main)
        for key, st in _SIDE_STREAMS.items():
            if dev.index in (None, key):
                self.stream.wait_stream(st)   # weight-gradient kernels of the finished layers
        piece = flat[lower:upper]
        with torch.cuda.stream(self.stream):
            if self.bf16:
                low = piece.to(torch.bfloat16)
                work = dist.all_reduce(low, op=dist.ReduceOp.SUM, async_op=True)
                self.pending.append((work, piece, low))
            else:
                work = dist.all_reduce(piece, op=dist.ReduceOp.SUM, async_op=True)
                self.pending.append((work, piece, None))

    def finish(self):
        """After backward: send what is left (the head of the buffer), then make the current stream wait for every piece."""
        if self.world_size == 1:
            return
        self.optimizer.assert_grads_are_views()
        self._issue(0)
        dev = self.optimizer.flat_grad.device
        for work, piece, low in self.pending:
            work.wait()
            if low is not None:
                with torch.cuda.stream(self.stream):
                    piece.copy_(low)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        self.pending = []
        self._upper = None

    reduce = finish


def shard_rows(total_rows, rank, world_size):
    """[begin, end) of this rank's rows of a global batch (remainder rows go to the low ranks)."""
    base, rem = divmod(total_rows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
