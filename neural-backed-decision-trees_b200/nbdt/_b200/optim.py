"""Fused SGD (momentum + weight decay) over ONE flat fp32 buffer.

All parameters (and their gradients) are re-seated as views of two flat buffers, so the optimizer step
is a single kernel launch (csrc/bn_ops.cu: sgd_kernel) and the data-parallel gradient all-reduce is a
single NCCL call on the flat gradient buffer -- no per-tensor launches, no bucket copies.  Semantics
are torch.optim.SGD's (dampening 0, no nesterov), the optimizer of the reference train step (main.py:207).
"""
import torch

from . import lib as L
from .ops import _count, _p


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.1, momentum=0.9, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        assert len(self.param_groups) == 1, "FusedSGD takes a single parameter group"
        ps = self.param_groups[0]["params"]
        for p in ps:
            L.require_cuda(p, "FusedSGD parameter")
            if p.dtype != torch.float32:
                raise TypeError("FusedSGD keeps fp32 master parameters")
        dev = ps[0].device
        total = sum(p.numel() for p in ps)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_momentum = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            p._nbdt_direct_grad = True  # conv weight gradients are accumulated into the flat buffer by the kernel
            off += n
        # packed bf16 copies / packed gradient buffers of the convolution weights: one pack and one unpack launch per step
        from .conv_ops import ConvWeightBank

        self.bank = ConvWeightBank()
        for p in ps:
            if p.dim() == 4:
                p._nbdt_bank = self.bank
        self._steps = 0
        self.grad_scale = 1.0

    def flush_grads(self):
        """Make every .grad complete: the convolution weight gradients sit in packed per-layer buffers until they are needed.
        step() and the gradient all-reduce call this; code that reads .grad right after backward() must call it itself."""
        self.bank.flush()

    def zero_grad(self, set_to_none=False):
        # gradients live in the flat buffer: keep the views, clear the storage with one memset
        self.bank.discard()
        self.flat_grad.zero_()
        off = 0
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            want = self.flat_grad[off:off + n].view_as(p)
            if p.grad is None or p.grad.data_ptr() != want.data_ptr():
                p.grad = want
            off += n

    @torch.no_grad()
    def reseat_grads(self):
        """autograd REPLACES p.grad when it was None (model.zero_grad() with set_to_none=True) instead of accumulating into the flat
        buffer's view: copy such gradients back and restore the views.  Called by step() and by the gradient all-reduce, which both
        work on the flat buffer."""
        off = 0
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad[off:off + n].data_ptr():
                self.flat_grad[off:off + n].copy_(p.grad.reshape(-1))
                p.grad = self.flat_grad[off:off + n].view_as(p)
            off += n

    def assert_grads_are_views(self):
        off = 0
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad[off:off + n].data_ptr():
                raise RuntimeError("a parameter's .grad is not a view of FusedSGD's flat gradient buffer any more: clear gradients with "
                                   "optimizer.zero_grad() (not model.zero_grad(set_to_none=True)) when gradients are all-reduced during backward")
            off += n

    def state_dict(self):
        """torch.optim's state_dict plus the flat momentum buffer and the step count (resume keeps the momentum)."""
        sd = super().state_dict()
        sd["nbdt_flat_momentum"] = self.flat_momentum.detach().clone()
        sd["nbdt_steps"] = self._steps
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        mom = state_dict.pop("nbdt_flat_momentum", None)
        steps = state_dict.pop("nbdt_steps", None)
        super().load_state_dict(state_dict)
        self.bank.invalidate()
        if mom is not None:
            self.flat_momentum.copy_(mom.to(self.flat_momentum.device))
        if steps is not None:
            self._steps = int(steps)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self.reseat_grads()
        self.bank.flush()
        from .nn_ops import join_wgrad_streams

        join_wgrad_streams(self.flat_param.device)  # weight-gradient kernels run on a side stream
        with torch.cuda.device(self.flat_param.device):
            L.check(L.load().nbdt_sgd_step(_p(self.flat_param), _p(self.flat_grad), _p(self.flat_momentum),
                                           self.flat_param.numel(), float(g["lr"]), float(g["momentum"]),
                                           float(g["weight_decay"]), float(self.grad_scale), int(self._steps == 0),
                                           L.stream_ptr(self.flat_param.device)))
        _count()
        self.bank.invalidate()  # the master weights changed: the packed bf16 copies are re-made at the next forward
        self._steps += 1
        return loss
