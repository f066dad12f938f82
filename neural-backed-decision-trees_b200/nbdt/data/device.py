"""Input pipeline on the GPU: the uint8 training set lives in HBM (CIFAR: 150 MB, TinyImagenet: 1.2 GB) and every batch is
cropped, flipped and normalised by ONE kernel (nbdt_augment_crop_flip, csrc/aux_ops.cu).  Replaces the reference's two
DataLoader worker processes running PIL transforms (main.py:136-141, nbdt/data/cifar.py:10-21), which feed ~5 k img/s --
a sixth of what one B200 trains on WRN-28-10.  Same distribution of crops / flips as torchvision's RandomCrop(padding) +
RandomHorizontalFlip; random numbers come from a torch.Generator on the device.
"""
import ctypes

import torch


class DeviceLoader:
    """Iterable of (images fp32 NCHW normalised, labels int64) batches, all on `device`.

    images_u8: uint8 array / tensor (N, H, W, 3); labels: N integers.  train=True: shuffle + RandomCrop(pad) + flip; else in order,
    no augmentation (pad = 0, no flip: ToTensor + Normalize only)."""

    def __init__(self, images_u8, labels, batch_size, mean, std, pad=4, train=True, device="cuda", seed=0, drop_last=False):
        self.device = torch.device(device)
        self.images = torch.as_tensor(images_u8).to(self.device, torch.uint8).contiguous()
        assert self.images.dim() == 4 and self.images.size(3) == 3, "expected (N, H, W, 3) uint8 images"
        self.labels = torch.as_tensor(labels).to(self.device, torch.int64)
        self.batch_size, self.pad, self.train, self.drop_last = batch_size, pad if train else 0, train, drop_last
        self.mean = (ctypes.c_float * 3)(*mean)
        self.std = (ctypes.c_float * 3)(*std)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)

    @classmethod
    def from_dataset(cls, dataset, batch_size, train=True, device="cuda", **kw):
        """From a dataset of this package that keeps its images as one uint8 array (`data`: CIFAR, synthetic)."""
        owner = dataset if hasattr(dataset, "data") else dataset.dataset
        mean = getattr(dataset, "MEAN", None) or __import__("nbdt.data.cifar", fromlist=["MEAN"]).MEAN
        std = getattr(dataset, "STD", None) or __import__("nbdt.data.cifar", fromlist=["STD"]).STD
        pad = getattr(dataset, "PAD", 4)
        return cls(owner.data, owner.targets, batch_size, mean, std, pad=pad, train=train, device=device, **kw)

    def __len__(self):
        n = self.images.size(0)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def augment(self, idx, params=None):
        """The batch of rows `idx` through the kernel; params (B, 3) int32 = (dx, dy, flip) or drawn here."""
        from nbdt._b200 import lib as L

        B, (H, W) = idx.numel(), self.images.shape[1:3]
        src = self.images.index_select(0, idx)
        if params is None:
            if self.train:
                params = torch.stack([torch.randint(0, 2 * self.pad + 1, (B,), device=self.device, generator=self.gen),
                                      torch.randint(0, 2 * self.pad + 1, (B,), device=self.device, generator=self.gen),
                                      torch.randint(0, 2, (B,), device=self.device, generator=self.gen)], dim=1).to(torch.int32)
            else:
                params = torch.zeros((B, 3), dtype=torch.int32, device=self.device)
        params = params.to(self.device, torch.int32).contiguous()
        out = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().nbdt_augment_crop_flip(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(params.data_ptr()),
                                                    ctypes.c_void_p(out.data_ptr()), B, H, W, self.pad, self.mean, self.std,
                                                    L.stream_ptr(self.device)))
        return out

    def __iter__(self):
        n = self.images.size(0)
        order = torch.randperm(n, device=self.device, generator=self.gen) if self.train else torch.arange(n, device=self.device)
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            yield self.augment(idx), self.labels.index_select(0, idx)
