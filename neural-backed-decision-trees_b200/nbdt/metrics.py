"""Top-k accuracy counters (reference: nbdt/metrics.py:7-35), without the per-sample Python loop."""
import torch

__all__ = names = ("top1", "top2", "top5", "top10")


class TopK:
    def __init__(self, k=1):
        self.k = k
        self.clear()

    def clear(self):
        self.correct = 0
        self.total = 0

    def forward(self, outputs, targets):
        _, preds = torch.topk(outputs, self.k)
        hits = (preds == targets.reshape(-1, 1)).any(dim=1)
        self.correct += int(hits.sum().item())
        self.total += targets.size(0)

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
