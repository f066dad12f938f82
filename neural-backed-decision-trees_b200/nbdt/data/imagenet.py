"""TinyImagenet200 / Imagenet1000 image-folder datasets with the reference's transforms (reference: nbdt/data/imagenet.py).

Expected layout under `root`: tiny-imagenet-200/{train/<wnid>/images/*.JPEG, val/images + val_annotations.txt} and
imagenet-1000/{train,val}/<wnid>/*.JPEG.  Nothing is downloaded here (the reference fetches a 237 MB zip / asks for a manual
ImageNet copy); without the files, or with NBDT_SYNTHETIC_DATA=1, synthetic images with the real wnid class lists are served.
"""
import os

from torch.utils.data import Dataset
from torchvision import datasets, transforms

from . import synthetic
from . import transforms as transforms_custom

__all__ = names = ("TinyImagenet200", "Imagenet1000")


def _wnids(dataset):
    from nbdt.thirdparty.wn import get_wnids
    from nbdt.utils import dataset_to_default_path_wnids

    return get_wnids(dataset_to_default_path_wnids(dataset))


class _TinyVal(Dataset):
    """val/images/*.JPEG labelled by val_annotations.txt (file, wnid, box...)."""

    def __init__(self, root, classes, transform=None):
        self.root = os.path.join(root, "tiny-imagenet-200", "val")
        self.transform = transform
        cls_to_idx = {c: i for i, c in enumerate(classes)}
        self.samples = []
        with open(os.path.join(self.root, "val_annotations.txt")) as f:
            for line in f:
                parts = line.split()
                if len(parts) >= 2 and parts[1] in cls_to_idx:
                    self.samples.append((os.path.join(self.root, "images", parts[0]), cls_to_idx[parts[1]]))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        path, label = self.samples[i]
        img = datasets.folder.default_loader(path)
        return (self.transform(img) if self.transform else img), label


class _FolderDataset(Dataset):
    MEAN, STD = None, None
    SIZE, DIRNAME, NAME = None, None, None

    def __init__(self, root="./data", *args, train=True, download=False, transform=None, **kwargs):
        super().__init__()
        self.root = root
        base = os.path.join(root, self.DIRNAME)
        self.dataset = None
        if not synthetic.forced() and os.path.isdir(os.path.join(base, "train")):
            if train:
                self.dataset = datasets.ImageFolder(os.path.join(base, "train"), *args, transform=transform, **kwargs)
            elif self.NAME == "TinyImagenet200":
                self.dataset = _TinyVal(root, sorted(os.listdir(os.path.join(base, "train"))), transform)
            else:
                self.dataset = datasets.ImageFolder(os.path.join(base, "val"), *args, transform=transform, **kwargs)
            self.classes = self.dataset.classes if hasattr(self.dataset, "classes") else sorted(os.listdir(os.path.join(base, "train")))
        else:
            if not synthetic.forced():
                print(f"[nbdt.data] {self.NAME}: no files under {base}; serving synthetic images instead")
            self.dataset = synthetic.SyntheticImages(_wnids(self.NAME), self.SIZE, synthetic.sample_count(train), transform, seed=int(train))
            self.classes = self.dataset.classes
        self.class_to_idx = {cls: i for i, cls in enumerate(self.classes)}

    @classmethod
    def transform_val_inverse(cls):
        return transforms_custom.InverseNormalize(cls.MEAN, cls.STD)

    def __getitem__(self, i):
        return self.dataset[i]

    def __len__(self):
        return len(self.dataset)


class TinyImagenet200(_FolderDataset):
    MEAN, STD = [0.4802, 0.4481, 0.3975], [0.2302, 0.2265, 0.2262]
    SIZE, DIRNAME, NAME = 64, "tiny-imagenet-200", "TinyImagenet200"
    PAD = 8

    @staticmethod
    def transform_train(input_size=64):
        return transforms.Compose([transforms.RandomCrop(input_size, padding=TinyImagenet200.PAD), transforms.RandomHorizontalFlip(),
                                   transforms.ToTensor(), transforms.Normalize(TinyImagenet200.MEAN, TinyImagenet200.STD)])

    @staticmethod
    def transform_val(input_size=-1):
        return transforms.Compose([transforms.ToTensor(), transforms.Normalize(TinyImagenet200.MEAN, TinyImagenet200.STD)])


class Imagenet1000(_FolderDataset):
    MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    SIZE, DIRNAME, NAME = 224, "imagenet-1000", "Imagenet1000"

    @staticmethod
    def transform_train(input_size=224):
        return transforms.Compose([transforms.RandomResizedCrop(input_size), transforms.RandomHorizontalFlip(), transforms.ToTensor(),
                                   transforms.Normalize(Imagenet1000.MEAN, Imagenet1000.STD)])

    @staticmethod
    def transform_val(input_size=224):
        return transforms.Compose([transforms.Resize(input_size + 32), transforms.CenterCrop(input_size), transforms.ToTensor(),
                                   transforms.Normalize(Imagenet1000.MEAN, Imagenet1000.STD)])
