"""Datasets and the input pipeline (reference: nbdt/data/).  `from nbdt import data` gives main.py the dataset classes
(`data.CIFAR10`, ...), the name tuples (`data.cifar.names`, `data.imagenet.names`, `data.custom.names`) and
`data.custom.add_arguments`.  Segmentation datasets (ADE20K, LookIntoPerson, PascalContext) are not part of this package.

Offline use: with NBDT_SYNTHETIC_DATA=1 (or when the files are absent and cannot be downloaded) every dataset serves seeded random
images of its real shape and class list, so that main.py runs end to end without the archives.

`data.device` holds the B200-side pipeline: the whole uint8 training set resident in HBM and one kernel per batch for
crop + flip + normalise (csrc/aux_ops.cu), instead of DataLoader worker processes.
"""
from . import cifar, custom, device, imagenet, transforms  # noqa: F401
from .cifar import *  # noqa: F401,F403
from .custom import *  # noqa: F401,F403
from .imagenet import *  # noqa: F401,F403
