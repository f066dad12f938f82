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
            if hasattr(self.optimizer, "flush_grads"):
                self.optimizer.flush_grads()  # packed convolution gradients -> flat buffer (one launch)
            dist.all_reduce(self.optimizer.flat_grad, op=dist.ReduceOp.SUM)


class OverlappedGradientAllReduce:
    """The same sum, issued in a few contiguous pieces of the flat gradient buffer WHILE backward is still running.

    Parameters sit in the flat buffer in definition order and backward finishes them roughly back to front.  Every parameter
    reports once per backward that its gradient has been written -- a post-accumulate hook for the gradients autograd
    accumulates (BatchNorm, classifier), a callback from ConvFunction.backward for the convolution weights, whose gradient kernel
    adds straight into the flat buffer -- and a piece goes to NCCL as soon as all of its parameters (and all pieces behind it:
    the collectives must be issued in the same order on every rank) have reported.  The communication stream waits for the main
    stream and the weight-gradient side stream at that moment; `finish()` after backward sends whatever is left and joins.  For
    WRN-28-10 the last stage holds 71 % of the gradient bytes and is in flight before the first two stages start their backward.
    Everything is stream-ordered (no host sync), so it is captured in a CUDA graph together with forward + backward.
    Needs FusedSGD (flat buffer) and its zero_grad (which keeps every .grad a view of that buffer)."""

    def __init__(self, optimizer, world_size, pieces=4, bf16=False):
        self.optimizer = optimizer
        self.world_size = world_size
        self.bf16 = bf16
        self.stream = None
        self.pending = []
        self.issued = []       # (lower, upper) element ranges sent in the current step, in order
        self.last_issued = []
        params = optimizer.param_groups[0]["params"]
        total = optimizer.flat_grad.numel()
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += p.numel()
        # piece k = parameters [starts[k], starts[k+1]); cut where the running size passes k/pieces of the buffer
        self.starts = [0]
        if world_size > 1:
            for k in range(1, max(1, pieces)):
                want = total * k // pieces
                i = min(range(len(params)), key=lambda j: abs(offs[j] - want))
                if i > self.starts[-1]:
                    self.starts.append(i)
        self.piece_of = [0] * len(params)
        for k, st in enumerate(self.starts):
            for i in range(st, len(params)):
                self.piece_of[i] = k
        self.bounds = [offs[i] for i in self.starts[1:]][::-1]   # element offsets of the cuts, descending (for logs / tests)
        self._elem = [offs[i] for i in self.starts] + [total]
        self._index = {id(p): i for i, p in enumerate(params)}
        self._reset()
        self._hooks = []
        if world_size > 1:
            from . import nn_ops

            for i, p in enumerate(params):
                self._hooks.append(p.register_post_accumulate_grad_hook(lambda param, i=i: self._fired(i)))
            nn_ops.GRAD_WRITTEN_CALLBACKS.append(self._param_written)
            if hasattr(optimizer, "bank"):
                optimizer.bank.defer_unpack = False  # pieces go out during backward: every layer adds its gradient into the flat buffer at once

    def _reset(self):
        self._remaining = [0] * len(self.starts)
        for k in self.piece_of:
            self._remaining[k] += 1
        self._seen = set()
        self._next = len(self.starts) - 1   # pieces are issued back to front

    def _param_written(self, param):
        i = self._index.get(id(param))
        if i is not None:
            self._fired(i)

    def _fired(self, i):
        if self.world_size == 1 or i in self._seen:
            return
        self._seen.add(i)
        self._remaining[self.piece_of[i]] -= 1
        while self._next > 0 and self._remaining[self._next] == 0:   # piece 0 (the head) always goes in finish()
            self._issue(self._next)
            self._next -= 1

    def _issue(self, k):
        """all-reduce piece k of the flat buffer, ordered after everything enqueued so far on the main and weight-gradient streams."""
        from .nn_ops import _SIDE_STREAMS

        flat = self.optimizer.flat_grad
        lower, upper = self._elem[k], self._elem[k + 1]
        self.issued.append((lower, upper))
        dev = flat.device
        if dev.type != "cuda":  # host tensors (gloo, the CPU test of the piece logic): nothing to overlap with
            dist.all_reduce(flat[lower:upper], op=dist.ReduceOp.SUM)
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        for key, st in _SIDE_STREAMS.items():
            if dev.index in (None, key):
                self.stream.wait_stream(st)   # weight-gradient kernels of the finished layers
        piece = flat[lower:upper]
        with torch.cuda.stream(self.stream):
            if self.bf16:
                low = piece.to(torch.bfloat16)
                self.pending.append((dist.all_reduce(low, op=dist.ReduceOp.SUM, async_op=True), piece, low))
            else:
                self.pending.append((dist.all_reduce(piece, op=dist.ReduceOp.SUM, async_op=True), piece, None))

    def finish(self):
        """After backward: send what is left (at least the head of the buffer), then make the current stream wait for every piece."""
        if self.world_size == 1:
            return
        self.optimizer.assert_grads_are_views()
        while self._next >= 0:   # parameters that got no gradient this step never report: their pieces go now
            self._issue(self._next)
            self._next -= 1
        dev = self.optimizer.flat_grad.device
        self.last_issued, self.issued = self.issued, []
        self._reset()
        if dev.type != "cuda":
            return
        for work, piece, low in self.pending:
            work.wait()
            if low is not None:
                with torch.cuda.stream(self.stream):
                    piece.copy_(low)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        self.pending = []

    reduce = finish


def shard_rows(total_rows, rank, world_size):
    """[begin, end) of this rank's rows of a global batch (remainder rows go to the low ranks)."""
    base, rem = divmod(total_rows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
