"""Top-k accuracy counters (reference: nbdt/metrics.py:7-35).

The reference runs torch.topk and a per-sample Python loop with a host sync on every batch.  Here CUDA batches are ranked by one
kernel (nbdt_topk_count, csrc/aux_ops.cu: one warp per row counts the scores that torch.topk would list before the target) and
the two counters stay on the device; the host reads them when `correct` / `total` / `report()` are asked for.
"""
import ctypes

import torch

__all__ = names = ("top1", "top2", "top5", "top10")


class TopK:
    def __init__(self, k=1):
        self.k = k
        self._dev = {}  # device -> uint64-as-int64 tensor [correct, total]
        self.clear()

    def clear(self):
        self._correct = 0
        self._total = 0
        for t in self._dev.values():
            t.zero_()

    def _drain(self):
        for t in self._dev.values():
            c, n = t.tolist()  # the one host sync
            self._correct += int(c)
            self._total += int(n)
            t.zero_()

    @property
    def correct(self):
        self._drain()
        return self._correct

    @correct.setter
    def correct(self, v):
        self._drain()
        self._correct = int(v)

    @property
    def total(self):
        self._drain()
        return self._total

    @total.setter
    def total(self, v):
        self._drain()
        self._total = int(v)

    def forward(self, outputs, targets):
        if outputs.dim() > 2:  # segmentation: (N, C, H, W) scores, (N, H, W) labels
            outputs = outputs.permute(0, *range(2, outputs.dim()), 1).reshape(-1, outputs.size(1))
            targets = targets.reshape(-1)
        if outputs.is_cuda and outputs.dtype in (torch.float32, torch.bfloat16, torch.float16):
            from nbdt._b200 import lib as L

            if outputs.stride(1) != 1 or (outputs.size(0) > 1 and outputs.stride(0) < outputs.size(1)):
                outputs = outputs.contiguous()
            targets = targets.to(device=outputs.device, dtype=torch.int64).contiguous()
            cnt = self._dev.get(outputs.device)
            if cnt is None:
                cnt = self._dev[outputs.device] = torch.zeros(2, dtype=torch.int64, device=outputs.device)
            ld = outputs.stride(0) if outputs.size(0) > 1 else outputs.size(1)
            with torch.cuda.device(outputs.device):
                L.check(L.load().nbdt_topk_count(ctypes.c_void_p(outputs.data_ptr()), L.dtype_code(outputs.dtype),
                                                 ctypes.c_void_p(targets.data_ptr()), outputs.size(0), outputs.size(1), ld, self.k,
                                                 ctypes.c_void_p(cnt.data_ptr()), L.stream_ptr(outputs.device)))
            return
        # host tensors (the reference's own CPU tests): same ranking rule, vectorised
        xt = outputs.gather(1, targets.reshape(-1, 1))
        idx = torch.arange(outputs.size(1), device=outputs.device).reshape(1, -1)
        rank = ((outputs > xt) | ((outputs == xt) & (idx < targets.reshape(-1, 1)))).sum(1)
        self._correct += int((rank < self.k).sum())
        self._total += targets.size(0)

    def report(self):
        return self.correct / (self.total or 1)

    def __repr__(self):
        return f"Top{self.k}: {self.report()}"

    __str__ = __repr__


def top1():
    return TopK(1)


def top2():
    return TopK(2)


def top5():
    return TopK(5)


def top10():
    return TopK(10)
