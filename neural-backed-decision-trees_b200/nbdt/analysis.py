"""Per-epoch analyzers that main.py hooks around its train / test loops (reference: nbdt/analysis.py).

The decision-rule analyzers (reference analysis.py:204-252) run the fused tree head ONCE per batch (nbdt.model.*EmbeddedDecisionRules
over the CUDA kernels) and count top-k hits on the device (nbdt.metrics); the reference walks the hierarchy in Python per node
and syncs per sample.  Analyzers that need WordNet (Superclass*) or render HTML (Visualize*) are not part of this package.
"""
import functools
import os
import time
from pathlib import Path

import numpy as np
import torch

from nbdt import metrics
from nbdt.model import HardEmbeddedDecisionRules as HardRules
from nbdt.model import SoftEmbeddedDecisionRules as SoftRules
from nbdt.utils import set_np_printoptions

__all__ = names = (
    "Noop",
    "ConfusionMatrix",
    "IgnoredSamples",
    "HardEmbeddedDecisionRules",
    "SoftEmbeddedDecisionRules",
    "Entropy",
    "TopEntropy",
    "TopDifference",
)


def add_arguments(parser):
    parser.add_argument("--superclass-wnids", nargs="*", type=str)
    parser.add_argument("--save-k", type=int, default=20)
    parser.add_argument("--visualize-decision-node-wnid", "--vdnw", type=str)


def _bracket(obj, phase):
    """Decorator running obj.start_<phase>(epoch) / obj.end_<phase>(epoch) around f(epoch, ...)."""
    start, end = getattr(obj, f"start_{phase}"), getattr(obj, f"end_{phase}")

    def decorator(f):
        @functools.wraps(f)
        def wrapper(epoch, *args, **kwargs):
            start(epoch)
            f(epoch, *args, **kwargs)
            end(epoch)

        return wrapper

    return decorator


class _PhaseContext:
    def __init__(self, obj, phase):
        self.obj, self.phase, self.epoch = obj, phase, 0

    def __call__(self, epoch):
        self.epoch = epoch
        return self

    def __enter__(self):
        return getattr(self.obj, f"start_{self.phase}")(self.epoch)

    def __exit__(self, *exc):
        getattr(self.obj, f"end_{self.phase}")(self.epoch)


class Noop:
    accepts_classes = lambda testset, **kwargs: testset.classes  # noqa: E731

    def __init__(self, classes=()):
        set_np_printoptions()
        self.classes = classes
        self.num_classes = len(classes)
        self.epoch = None

    epoch_function = property(lambda self: _bracket(self, "epoch"))
    train_function = property(lambda self: _bracket(self, "train"))
    test_function = property(lambda self: _bracket(self, "test"))
    epoch_context = property(lambda self: _PhaseContext(self, "epoch"))

    def start_epoch(self, epoch):
        self.epoch = epoch

    def start_train(self, epoch):
        assert epoch == self.epoch

    def update_batch(self, outputs, targets, images):
        return self._update_batch(outputs, targets)

    def _update_batch(self, outputs, targets):
        pass

    def end_train(self, epoch):
        assert epoch == self.epoch

    def start_test(self, epoch):
        assert epoch == self.epoch

    def end_test(self, epoch):
        assert epoch == self.epoch

    def end_epoch(self, epoch):
        assert epoch == self.epoch


class ConfusionMatrix(Noop):
    """k x k counts, rows = labels, columns = predictions; accumulated on the tensors' device with one bincount per batch."""

    def __init__(self, classes):
        super().__init__(classes)
        self.k = len(classes)
        self.m = None
        self._acc = None

    def start_train(self, epoch):
        super().start_train(epoch)
        raise NotImplementedError()

    def start_test(self, epoch):
        super().start_test(epoch)
        self.m = np.zeros((self.k, self.k))
        self._acc = None

    def _update_batch(self, outputs, targets):
        predicted = outputs.max(1)[1]
        if predicted.dim() != 1:
            return
        flat = targets.reshape(-1).long() * self.k + predicted.reshape(-1).long()
        counts = torch.bincount(flat, minlength=self.k * self.k)
        self._acc = counts if self._acc is None else self._acc + counts

    def _sync(self):
        if self._acc is not None:
            self.m = self.m + self._acc.reshape(self.k, self.k).double().cpu().numpy()
            self._acc = None

    def end_test(self, epoch):
        super().end_test(epoch)
        recall = self.recall()
        for row, cls in zip(recall, self.classes):
            print(row, cls)
        print(recall.diagonal(), "(diagonal)")

    @staticmethod
    def update(confusion_matrix, preds, labels):
        np.add.at(confusion_matrix, (np.asarray(tuple(labels)), np.asarray(tuple(preds))), 1)

    @staticmethod
    def normalize(confusion_matrix, axis):
        total = confusion_matrix.astype(np.float64).sum(axis=axis)
        return confusion_matrix / (total[:, None] if axis == 1 else total[None])

    def recall(self):
        self._sync()
        return ConfusionMatrix.normalize(self.m, 1)

    def precision(self):
        self._sync()
        return ConfusionMatrix.normalize(self.m, 0)


class IgnoredSamples(Noop):
    """Counts the samples the hard rules could not place (outputs[:, 0] == -1)."""

    def __init__(self, classes=()):
        super().__init__(classes)
        self.ignored = None

    def start_test(self, epoch):
        super().start_test(epoch)
        self.ignored = 0

    def _update_batch(self, outputs, targets):
        self.ignored += outputs[:, 0].eq(-1).sum().item()
        return self.ignored

    def end_test(self, epoch):
        super().end_test(epoch)
        print("Ignored Samples: {}".format(self.ignored))


class DecisionRules(Noop):
    """Accuracy of the network's outputs passed through embedded decision rules (reference analysis.py:204-237)."""

    accepts_tree = lambda tree, **kwargs: tree  # noqa: E731
    accepts_dataset = lambda trainset, **kwargs: trainset.__class__.__name__  # noqa: E731
    accepts_path_graph = True
    accepts_path_wnids = True
    accepts_metric = True

    name = "NBDT"

    def __init__(self, *args, Rules=HardRules, tree=None, metric="top1", **kwargs):
        self.rules = Rules(*args, **kwargs, tree=tree)
        super().__init__(self.rules.tree.classes)
        self.metric = getattr(metrics, metric)()
        self.best_accuracy = 0

    def start_test(self, epoch):
        self.metric.clear()

    def _update_batch(self, outputs, targets):
        outputs = self.rules.forward(outputs)      # one launch of the fused tree head
        self.metric.forward(outputs, targets)      # one launch of the counter kernel, no sync
        return round(self.metric.correct / float(self.metric.total or 1), 4) * 100  # main.py prints it per batch: one sync

    def end_test(self, epoch):
        super().end_test(epoch)
        accuracy = round(self.metric.correct / (self.metric.total or 1) * 100.0, 2)
        self.best_accuracy = max(accuracy, self.best_accuracy)
        print(f"[{self.name}] Accuracy: {accuracy}%, {self.metric.correct}/{self.metric.total} | {self.name} Best Accuracy: {self.best_accuracy}%")


class HardEmbeddedDecisionRules(DecisionRules):
    name = "NBDT-Hard"


class SoftEmbeddedDecisionRules(DecisionRules):
    name = "NBDT-Soft"

    def __init__(self, *args, Rules=None, **kwargs):
        super().__init__(*args, Rules=SoftRules, **kwargs)


class ScoreSave(Noop):
    """Scores every test sample and keeps the k highest / lowest scoring images (selection with torch.topk on the device)."""

    def __init__(self, *args, classes=(), save_k=20, path="out/score-{epoch}-{time}/image-{suffix}-{i}-{score:.2e}.jpg", **kwargs):
        super().__init__(*args, classes=classes, **kwargs)
        self.k = save_k
        self.path = Path(path)
        self.time = int(time.time())
        self.reset()

    accepts_save_k = True

    def start_test(self, epoch):
        super().start_test(epoch)
        self.reset()

    def reset(self):
        self.max = []  # (image, score) pairs, best first
        self.min = []

    def score(self, outputs, targets, images):
        raise NotImplementedError()

    def _keep(self, kept, scores, images, largest):
        if kept:
            scores = torch.cat([torch.stack([s for _, s in kept]).to(scores.device), scores])
            images = torch.cat([torch.stack([im for im, _ in kept]).to(images.device), images])
        k = min(self.k, scores.numel())
        top = torch.topk(scores, k, largest=largest).indices
        return [(images[i], scores[i]) for i in top]

    def update_batch(self, outputs, targets, images):
        super().update_batch(outputs, targets, images)
        scores = self.score(outputs, targets, images).reshape(-1).float()
        self.max = self._keep(self.max, scores, images, True)
        self.min = self._keep(self.min, scores, images, False)

    def end_test(self, epoch):
        super().end_test(epoch)
        from PIL import Image

        directory = str(self.path.parent).format(time=self.time, epoch=self.epoch)
        os.makedirs(directory, exist_ok=True)
        for name, suffix, lst in (("highest", "max", self.max), ("lowest", "min", self.min)):
            print(f"==> Saving {self.k} {name} scored images in {directory}")
            for i, (image, score) in enumerate(lst):
                arr = (image.permute(1, 2, 0) * 255).clamp(0, 255).detach().cpu().numpy().astype(np.uint8)
                Image.fromarray(arr).save(str(self.path).format(epoch=self.epoch, i=i, suffix=suffix, score=float(score), time=self.time))


def _entropy(probs):
    return -(probs * torch.log(probs.clamp_min(1e-30))).sum(1)


class Entropy(ScoreSave):
    """Entropy of the softmax of the outputs: mean / std over the test set, most and least certain samples saved."""

    def __init__(self, *args, path="out/entropy-{epoch}-{time}/image-{suffix}-{i}-{score:.2e}.jpg", **kwargs):
        super().__init__(*args, path=path, **kwargs)

    def reset(self):
        super().reset()
        self.sum = self.sumsq = None
        self.count = 0

    def score(self, outputs, targets, images):
        return _entropy(torch.softmax(outputs.float(), dim=1))

    def update_batch(self, outputs, targets, images):
        super().update_batch(outputs, targets, images)
        s = self.score(outputs, targets, images).double()
        self.sum = s.sum() if self.sum is None else self.sum + s.sum()
        self.sumsq = (s * s).sum() if self.sumsq is None else self.sumsq + (s * s).sum()
        self.count += s.numel()

    def end_test(self, epoch):
        super().end_test(epoch)
        if self.count:
            mean = float(self.sum) / self.count
            var = max(float(self.sumsq) / self.count - mean * mean, 0.0)
            print(f"[Entropy] avg {mean:.2e}, std {var ** 0.5:.2e}")


class TopEntropy(Entropy):
    """Entropy of the two largest softmax probabilities, renormalised."""

    def score(self, outputs, targets, images):
        top2 = torch.softmax(outputs.float(), dim=1).topk(2, dim=1).values
        return _entropy(top2 / top2.sum(1, keepdim=True))


class TopDifference(ScoreSave):
    """Gap between the two largest softmax probabilities."""

    def score(self, outputs, targets, images):
        top2 = torch.softmax(outputs.float(), dim=1).topk(2, dim=1).values
        return top2[:, 0] - top2[:, 1]
