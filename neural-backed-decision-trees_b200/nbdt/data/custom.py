"""Label-filtered views of the image datasets (reference: nbdt/data/custom.py): keep each sample of label l with probability
p[l] (Resample), keep only / all but a set of labels (Include / Exclude).  The index is built from the dataset's label list when
it has one (`targets`), not by decoding every image as the reference does."""
import random

from torch.utils.data import Dataset

from . import cifar, imagenet

__all__ = names = tuple(f"{d}{kind}Labels" for kind in ("Include", "Exclude", "Resample")
                        for d in ("CIFAR10", "CIFAR100", "TinyImagenet200", "Imagenet1000"))


def add_arguments(parser):
    parser.add_argument("--probability-labels", nargs="*", type=float)
    parser.add_argument("--include-labels", nargs="*", type=int)
    parser.add_argument("--exclude-labels", nargs="*", type=int)
    parser.add_argument("--include-classes", nargs="*", type=int)


def _labels(dataset):
    for owner in (dataset, getattr(dataset, "dataset", None)):
        t = getattr(owner, "targets", None)
        if t is not None and len(t) == len(dataset):
            return [int(v) for v in t]
    return [int(dataset[i][1]) for i in range(len(dataset))]


class ResampleLabelsDataset(Dataset):
    """Every sample with label l is kept with probability probability_labels[l] (seeded).  drop_classes renumbers the surviving
    labels to 0..m-1 and shrinks `classes`."""

    accepts_probability_labels = True

    def __init__(self, dataset, probability_labels=1, drop_classes=False, seed=0):
        self.dataset = dataset
        k = len(dataset.classes)
        ps = probability_labels
        if not isinstance(ps, (tuple, list)):
            ps = [ps] * k
        elif len(ps) == 1:
            ps = list(ps) * k
        assert len(ps) == k, f"Length of probabilities vector {len(ps)} must equal that of the dataset classes {k}."
        self.probability_labels = list(ps)
        self.drop_classes = drop_classes
        self.labels = [i for i in range(k) if not drop_classes or ps[i] > 0]
        self.classes = [dataset.classes[i] for i in self.labels]
        assert self.labels, "No labels are included in `include_labels`"
        rng = random.Random(seed)
        self.new_to_old = [old for old, label in enumerate(_labels(dataset)) if rng.random() < self.probability_labels[label]]

    def __getitem__(self, index_new):
        sample, label = self.dataset[self.new_to_old[index_new]]
        return sample, (self.labels.index(label) if self.drop_classes else label)

    def __len__(self):
        return len(self.new_to_old)


class IncludeLabelsDataset(ResampleLabelsDataset):
    accepts_include_labels = True
    accepts_probability_labels = False

    def __init__(self, dataset, include_labels=(0,)):
        include = set(include_labels)
        super().__init__(dataset, probability_labels=[int(label in include) for label in range(len(dataset.classes))])


class IncludeClassesDataset(IncludeLabelsDataset):
    accepts_include_labels = False
    accepts_include_classes = True

    def __init__(self, dataset, include_classes=()):
        super().__init__(dataset, include_labels=[dataset.classes.index(cls) for cls in include_classes])


class ExcludeLabelsDataset(IncludeLabelsDataset):
    accepts_include_labels = False
    accepts_exclude_labels = True

    def __init__(self, dataset, exclude_labels=(0,)):
        super().__init__(dataset, include_labels=set(range(len(dataset.classes))) - set(exclude_labels))


def _bind(view, base, kwarg, default):
    class Cls(view):
        def __init__(self, *args, root="./data", **kwargs):
            value = kwargs.pop(kwarg, default)
            super().__init__(base(*args, root=root, **kwargs), **{kwarg: value})

    Cls.__name__ = Cls.__qualname__ = f"{base.__name__}{view.__name__.replace('Dataset', '')}"
    for f in ("transform_train", "transform_val", "transform_val_inverse"):
        setattr(Cls, f, staticmethod(getattr(base, f)))
    return Cls


for _base in (cifar.CIFAR10, cifar.CIFAR100, imagenet.TinyImagenet200, imagenet.Imagenet1000):
    globals()[f"{_base.__name__}IncludeLabels"] = _bind(IncludeLabelsDataset, _base, "include_labels", (0,))
    globals()[f"{_base.__name__}ExcludeLabels"] = _bind(ExcludeLabelsDataset, _base, "exclude_labels", (0,))
    globals()[f"{_base.__name__}ResampleLabels"] = _bind(ResampleLabelsDataset, _base, "probability_labels", 1)
