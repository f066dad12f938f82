"""Inverse of torchvision's Normalize, for analyzers that save images (reference: nbdt/data/transforms.py)."""
import torch


class InverseNormalize:
    def __init__(self, mean, std):
        self.mean = torch.tensor(mean, dtype=torch.float32).reshape(1, -1, 1, 1)
        self.std = torch.tensor(std, dtype=torch.float32).reshape(1, -1, 1, 1)

    def __call__(self, sample):
        return sample * self.std + self.mean

    def to(self, device):
        self.mean, self.std = self.mean.to(device), self.std.to(device)
        return self
