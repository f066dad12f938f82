"""Seeded random stand-ins for the image datasets (no network, no archives): real shapes, real class lists."""
import os

import numpy as np
import torch
from torch.utils.data import Dataset


def forced():
    return os.environ.get("NBDT_SYNTHETIC_DATA", "") not in ("", "0")


class SyntheticImages(Dataset):
    """uint8 HWC images + labels drawn once from a seeded generator; `transform` is applied like torchvision datasets do."""

    def __init__(self, classes, size, n, transform=None, seed=0):
        self.classes = list(classes)
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}
        rng = np.random.default_rng(seed)
        self.data = rng.integers(0, 256, size=(n, size, size, 3), dtype=np.uint8)
        self.targets = rng.integers(0, len(self.classes), size=n).tolist()
        self.transform = transform

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        from PIL import Image

        img = Image.fromarray(self.data[i])
        if self.transform is not None:
            img = self.transform(img)
        else:
            img = torch.from_numpy(self.data[i]).permute(2, 0, 1).float() / 255
        return img, self.targets[i]


def sample_count(train):
    return int(os.environ.get("NBDT_SYNTHETIC_SAMPLES", "2048" if train else "512"))
