"""Constants, path helpers and the `accepts_*` plugin protocol.

Mirrors the helpers `main.py` and the loaders import from the reference's nbdt/utils.py
(constants utils.py:18-56, path helpers :59-75, generate_kwargs :78-97, Colors :118-139,
progress_bar :187-228, generate_checkpoint_fname :268-335, coerce/uncoerce :338-347).
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

METHODS = ("wordnet", "random", "induced")
DATASET_TO_NUM_CLASSES = {
    "CIFAR10": 10,
    "CIFAR100": 100,
    "TinyImagenet200": 200,
    "Imagenet1000": 1000,
    "Cityscapes": 19,
    "PascalContext": 59,
    "LookIntoPerson": 20,
    "ADE20K": 150,
}
DATASETS = tuple(DATASET_TO_NUM_CLASSES)
DATASET_TO_CLASSES = {
    "CIFAR10": ["airplane", "automobile", "bird", "cat", "deer", "dog", "frog", "horse", "ship", "truck"]
}


def fwd():
    """Directory of the installed package (where hierarchies/ and wnids/ live)."""
    return Path(__file__).parent.absolute()


def hierarchy_to_path_graph(dataset, hierarchy):
    return os.path.join(fwd(), f"hierarchies/{dataset}/graph-{hierarchy}.json")


def dataset_to_default_path_graph(dataset):
    return hierarchy_to_path_graph(dataset, "induced")


def dataset_to_default_path_wnids(dataset):
    return os.path.join(fwd(), f"wnids/{dataset}.txt")


def get_directory(dataset, root="./nbdt/hierarchies"):
    return os.path.join(root, dataset)


def makeparentdirs(path):
    os.makedirs(Path(path).parent, exist_ok=True)


class Colors:
    RED, GREEN, CYAN, BOLD, ENDC = "\x1b[31m", "\x1b[32m", "\x1b[36m", "\033[1m", "\033[0m"

    @classmethod
    def _emit(cls, code, args):
        print(code + str(args[0]), *args[1:], cls.ENDC)

    @classmethod
    def red(cls, *args):
        cls._emit(cls.RED, args)

    @classmethod
    def green(cls, *args):
        cls._emit(cls.GREEN, args)

    @classmethod
    def cyan(cls, *args):
        cls._emit(cls.CYAN, args)

    @classmethod
    def bold(cls, *args):
        cls._emit(cls.BOLD, args)


def generate_kwargs(args, object, name="Dataset", globals={}, kwargs=None):
    """Collect constructor kwargs for a plugin class from argparse `args`.

    A class opts in per key with an `accepts_<key>` attribute: `True` copies `args.<key>`
    when it is set, a callable is invoked with the caller's locals (`globals`).
    """
    kwargs = kwargs or {}
    for attr in dir(object):
        if not attr.startswith("accepts_"):
            continue
        accepts = getattr(object, attr, False)
        if not accepts:
            continue
        key = attr[len("accepts_"):]
        assert key in args or callable(accepts)
        value = getattr(args, key, None)
        if callable(accepts):
            kwargs[key] = accepts(**globals)
            Colors.cyan(f"{key}:\t(callable)")
        elif value is not None:
            kwargs[key] = value
            Colors.cyan(f"{key}:\t{value}")
    return kwargs


def set_np_printoptions():
    np.set_printoptions(formatter={"float": lambda x: "{0:0.3f}".format(x)})


# ---- console progress bar (cosmetic; main.py imports it) -------------------------------------
try:
    term_width = int(os.popen("stty size 2>/dev/null", "r").read().split()[1])
except Exception:
    term_width = 80

_BAR = 65
_t_last = _t_begin = time.time()


def format_time(seconds):
    units = (("D", 86400), ("h", 3600), ("m", 60), ("s", 1))
    out, shown = "", 0
    for suffix, size in units:
        n = int(seconds // size)
        seconds -= n * size
        if n > 0 and shown < 2:
            out += f"{n}{suffix}"
            shown += 1
    ms = int(seconds * 1000)
    if ms > 0 and shown < 2:
        out += f"{ms}ms"
    return out or "0ms"


def progress_bar(current, total, msg=None):
    global _t_last, _t_begin
    now = time.time()
    if current == 0:
        _t_begin = now
    done = int(_BAR * current / total)
    bar = "=" * done + ">" + "." * (_BAR - done - 1)
    text = f"  Step: {format_time(now - _t_last)} | Tot: {format_time(now - _t_begin)}"
    if msg:
        text += " | " + msg
    _t_last = now
    sys.stdout.write(f" [{bar}]{text} {current + 1}/{total} ")
    sys.stdout.write("\r" if current < total - 1 else "\n")
    sys.stdout.flush()


def generate_checkpoint_fname(
    dataset, arch, path_graph, wnid=None, name="", trainset=None, include_labels=(), exclude_labels=(),
    include_classes=(), num_samples=0, tree_supervision_weight=1, fine_tune=False, loss="CrossEntropyLoss",
    lr=0.1, tree_supervision_weight_end=None, tree_supervision_weight_power=1, xent_weight=1,
    xent_weight_end=None, xent_weight_power=1, tree_start_epochs=None, tree_update_every_epochs=None,
    tree_update_end_epochs=None, **kwargs,
):
    """Checkpoint file stem derived from the hyper-parameters (same naming scheme as the
    reference so released checkpoints resolve): ckpt-<dataset>-<arch>[-lr..][-name][-hierarchy]..."""
    parts = ["ckpt", dataset, arch]
    if lr != 0.1:
        parts.append(f"lr{lr}")
    if name:
        parts.append(name)
    if path_graph and "TreeSupLoss" in loss:
        parts.append(Path(path_graph).stem.replace("graph-", "", 1))
    for tag, labels in (("incl", include_labels), ("excl", exclude_labels), ("incc", include_classes)):
        if labels:
            parts.append(tag + ",".join(map(str, labels)))
    if num_samples:
        parts.append(f"samples{num_samples}")
    if len(loss) > 1 or loss[0] != "CrossEntropyLoss":
        parts.append(",".join(loss))
        for tag, value, skip in (
            ("tsw", tree_supervision_weight, (None, 1)),
            ("tswe", tree_supervision_weight_end, (tree_supervision_weight, None)),
            ("tswp", tree_supervision_weight_power, (None, 1)),
            ("xw", xent_weight, (None, 1)),
            ("xwe", xent_weight_end, (xent_weight, None)),
            ("xwp", xent_weight_power, (None, 1)),
        ):
            if value not in skip:
                parts.append(f"{tag}{value}")
    if "SoftTreeLoss" in loss:
        for tag, value in (("tse", tree_start_epochs), ("tueve", tree_update_every_epochs), ("tuene", tree_update_end_epochs)):
            if value is not None:
                parts.append(f"{tag}{value}")
    return "-".join(parts)


def coerce_tensor(x, is_label=False):
    """(N,C,H,W) segmentation logits -> (N*H*W, C) rows (labels -> a column)."""
    if is_label:
        return x.reshape(-1, 1)
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])


def uncoerce_tensor(x, original_shape):
    n, c, h, w = original_shape
    return x.reshape(n, h, w, c).permute(0, 3, 1, 2)


def load_image_from_path(path):
    """PIL image from a local path or an http(s) URL (reference: nbdt/utils.py:100-110)."""
    import io
    from urllib.request import Request, urlopen

    from PIL import Image

    if "http" in path:
        request = Request(path, headers={"User-Agent": "Mozilla/5.0 (nbdt-b200)"})
        return Image.open(io.BytesIO(urlopen(request).read()))
    return Image.open(path)


def get_mean_and_std(dataset):
    """Per-channel mean and std of a dataset of (image tensor, label) pairs (reference: nbdt/utils.py:139-156:
    the mean over the samples of each image's per-channel mean / std)."""
    import torch

    loader = torch.utils.data.DataLoader(dataset, batch_size=1, shuffle=False, num_workers=0)
    mean, std, n = torch.zeros(3), torch.zeros(3), 0
    for inputs, _ in loader:
        for i in range(3):
            mean[i] += inputs[:, i, :, :].mean()
            std[i] += inputs[:, i, :, :].std()
        n += 1
    return mean / max(n, 1), std / max(n, 1)


def init_params(net):
    """Kaiming / constant / small-normal initialisation of conv / batch-norm / linear layers
    (reference: nbdt/utils.py:159-172)."""
    import torch.nn as nn
    import torch.nn.init as init

    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            init.kaiming_normal_(m.weight, mode="fan_out")
            if m.bias is not None:
                init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            init.constant_(m.weight, 1)
            init.constant_(m.bias, 0)
        elif isinstance(m, nn.Linear):
            init.normal_(m.weight, std=1e-3)
            if m.bias is not None:
                init.constant_(m.bias, 0)
