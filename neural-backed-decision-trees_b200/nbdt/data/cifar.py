"""CIFAR10 / CIFAR100 with the reference's transforms (reference: nbdt/data/cifar.py)."""
from torchvision import datasets, transforms

from . import synthetic
from . import transforms as transforms_custom

__all__ = names = ("CIFAR10", "CIFAR100")

MEAN, STD = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)
PAD = 4

CIFAR10_CLASSES = ("airplane", "automobile", "bird", "cat", "deer", "dog", "frog", "horse", "ship", "truck")


class CIFAR:
    @staticmethod
    def transform_train():
        return transforms.Compose([transforms.RandomCrop(32, padding=PAD), transforms.RandomHorizontalFlip(), transforms.ToTensor(),
                                   transforms.Normalize(MEAN, STD)])

    @staticmethod
    def transform_val():
        return transforms.Compose([transforms.ToTensor(), transforms.Normalize(MEAN, STD)])

    @staticmethod
    def transform_val_inverse():
        return transforms_custom.InverseNormalize(MEAN, STD)


def _make(base, classes_fallback):
    class Cls(base, CIFAR):
        def __init__(self, root="./data", train=True, transform=None, target_transform=None, download=False, **kwargs):
            self._synthetic = None
            if not synthetic.forced():
                try:
                    super().__init__(root=root, train=train, transform=transform, target_transform=target_transform, download=download, **kwargs)
                    return
                except Exception as e:  # no archive and no network
                    print(f"[nbdt.data] {base.__name__}: {type(e).__name__}: {e}; serving synthetic images instead")
            self._synthetic = synthetic.SyntheticImages(classes_fallback(), 32, synthetic.sample_count(train), transform, seed=int(train))
            self.classes = self._synthetic.classes
            self.class_to_idx = self._synthetic.class_to_idx
            self.data, self.targets = self._synthetic.data, self._synthetic.targets
            self.transform, self.target_transform, self.train, self.root = transform, target_transform, train, root

        def __len__(self):
            return len(self._synthetic) if self._synthetic is not None else super().__len__()

        def __getitem__(self, i):
            return self._synthetic[i] if self._synthetic is not None else super().__getitem__(i)

    Cls.__name__ = Cls.__qualname__ = base.__name__
    return Cls


def _cifar100_classes():
    from nbdt.utils import DATASET_TO_CLASSES

    classes = DATASET_TO_CLASSES.get("CIFAR100") if isinstance(DATASET_TO_CLASSES, dict) else None
    return list(classes) if classes else [f"class{i:03d}" for i in range(100)]


CIFAR10 = _make(datasets.CIFAR10, lambda: list(CIFAR10_CLASSES))
CIFAR100 = _make(datasets.CIFAR100, _cifar100_classes)
