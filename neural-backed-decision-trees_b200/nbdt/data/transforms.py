"""Inverse of torchvision's Normalize, for analyzers that save images (reference: nbdt/data/transforms.py)."""
import torch


def _per_channel(values):
    """(C,) statistics -> a (1, C, 1, 1) fp32 tensor that broadcasts over NCHW batches."""
    return torch.as_tensor(values, dtype=torch.float32).view(1, -1, 1, 1)


class InverseNormalize:
    """x_normalised -> x: multiplies by the standard deviation and adds the mean, channel-wise."""

    def __init__(self, mean, std):
        self.mean, self.std = _per_channel(mean), _per_channel(std)

    def to(self, device):
        self.mean, self.std = (t.to(device) for t in (self.mean, self.std))
        return self

    def __call__(self, sample):
        return torch.addcmul(self.mean, sample, self.std)
