"""Tree-supervision losses.

Same public surface as the reference's nbdt/loss.py: module-level `names`, `add_arguments`,
`set_default_values`, `CrossEntropyLoss` (loss.py:19-94) and the classes `TreeSupLoss`,
`HardTreeSupLoss`, `SoftTreeSupLoss`, `SoftTreeLoss`, `SoftSegTreeSupLoss` (loss.py:97-327)
with the same constructor arguments and `accepts_*` plugin attributes.

When the wrapped criterion is a default `nn.CrossEntropyLoss` (the only criterion main.py and
the ClassyVision adapter ever pass) the whole of `TreeSupLoss.forward` -- plain cross-entropy,
decision rules, tree cross-entropy, the weighted sum AND the gradient w.r.t. the logits -- is
ONE kernel launch (nbdt_soft_loss / nbdt_hard_loss in include/nbdt_b200.h).  Any other
criterion is applied on top of the differentiable soft-rules kernel.
"""
import os

import torch
import torch.nn as nn

from nbdt._b200 import ops
from nbdt.model import HardEmbeddedDecisionRules, SoftEmbeddedDecisionRules
from nbdt.tree import Node, Tree  # noqa: F401
from nbdt.utils import (
    coerce_tensor,
    dataset_to_default_path_graph,
    dataset_to_default_path_wnids,
    hierarchy_to_path_graph,
    uncoerce_tensor,
)

__all__ = names = ("HardTreeSupLoss", "SoftTreeSupLoss", "SoftTreeLoss", "CrossEntropyLoss")


def add_arguments(parser):
    parser.add_argument("--xent-weight", "--xw", type=float, help="Weight for cross entropy term")
    parser.add_argument("--xent-weight-end", "--xwe", type=float,
                        help="Weight for cross entropy term at end of training. If not set, set to xw")
    parser.add_argument("--xent-weight-power", "--xwp", type=float, help="Raise progress to this power.")
    parser.add_argument("--tree-supervision-weight", "--tsw", type=float, default=1,
                        help="Weight assigned to tree supervision losses")
    parser.add_argument("--tree-supervision-weight-end", "--tswe", type=float,
                        help="Weight assigned to tree supervision losses at end of training. "
                        "If not set, this is equal to tsw")
    parser.add_argument("--tree-supervision-weight-power", "--tswp", type=float,
                        help="Raise progress to this power. > 1 to trend towards tswe more slowly. "
                        "< 1 to trend more quickly")
    parser.add_argument("--tree-start-epochs", "--tse", type=int,
                        help="epoch count to start tree supervision loss from (generate tree at that pt)")
    parser.add_argument("--tree-update-end-epochs", "--tuene", type=int,
                        help="epoch count to stop generating new trees at")
    parser.add_argument("--tree-update-every-epochs", "--tueve", type=int,
                        help="Recompute tree from weights every (this many) epochs")


def set_default_values(args):
    assert not (args.hierarchy and args.path_graph), "Only one, between --hierarchy and --path-graph can be provided."
    if args.hierarchy and not args.path_graph:
        args.path_graph = hierarchy_to_path_graph(args.dataset, args.hierarchy)
    if not args.path_graph:
        args.path_graph = dataset_to_default_path_graph(args.dataset)
    if not args.path_wnids:
        args.path_wnids = dataset_to_default_path_wnids(args.dataset)


CrossEntropyLoss = nn.CrossEntropyLoss


def _is_plain_cross_entropy(criterion):
    """True when `criterion` is the mean-reduced, unweighted, unsmoothed CE the fused kernels compute."""
    return (
        type(criterion) is nn.CrossEntropyLoss
        and criterion.reduction == "mean"
        and criterion.weight is None
        and criterion.ignore_index == -100
        and getattr(criterion, "label_smoothing", 0.0) == 0.0
    )


class TreeSupLoss(nn.Module):

    accepts_tree = lambda tree, **kwargs: tree
    accepts_criterion = lambda criterion, **kwargs: criterion
    accepts_dataset = lambda trainset, **kwargs: trainset.__class__.__name__
    accepts_path_graph = True
    accepts_path_wnids = True
    accepts_tree_supervision_weight = True
    accepts_classes = lambda trainset, **kwargs: trainset.classes
    accepts_hierarchy = True
    accepts_tree_supervision_weight_end = True
    accepts_tree_supervision_weight_power = True
    accepts_xent_weight = True
    accepts_xent_weight_end = True
    accepts_xent_weight_power = True

    fused_kind = None  # "soft" / "hard" in the subclasses that have a fused kernel

    def __init__(self, dataset, criterion, path_graph=None, path_wnids=None, classes=None, hierarchy=None,
                 Rules=HardEmbeddedDecisionRules, tree=None, tree_supervision_weight=1.0,
                 tree_supervision_weight_end=None, tree_supervision_weight_power=1, xent_weight=1,
                 xent_weight_end=None, xent_weight_power=1):
        super().__init__()
        if not tree:
            tree = Tree(dataset, path_graph, path_wnids, classes, hierarchy=hierarchy)
        self.num_classes = len(tree.classes)
        self.tree = tree
        self.rules = Rules(tree=tree)
        self.tree_supervision_weight = tree_supervision_weight
        self.tree_supervision_weight_end = (
            tree_supervision_weight if tree_supervision_weight_end is None else tree_supervision_weight_end
        )
        self.tree_supervision_weight_power = tree_supervision_weight_power
        self.xent_weight = xent_weight
        self.xent_weight_end = xent_weight if xent_weight_end is None else xent_weight_end
        self.xent_weight_power = xent_weight_power
        self.criterion = criterion
        self.progress = 1
        self.epochs = 0
        self.batch_rows = 0  # rows the mean is taken over; 0 = this call's rows (set for sharded batches)

    @staticmethod
    def assert_output_not_nbdt(outputs):
        """
        >>> x = torch.randn(1, 3, 224, 224)
        >>> TreeSupLoss.assert_output_not_nbdt(x)  # all good!
        >>> x._nbdt_output_flag = True
        >>> TreeSupLoss.assert_output_not_nbdt(x)  #doctest: +ELLIPSIS
        Traceback (most recent call last):
            ...
        AssertionError: ...
        """
        assert getattr(outputs, "_nbdt_output_flag", False) is False, (
            "Uh oh! Looks like you passed an NBDT model's output to an NBDT loss. NBDT losses are "
            "designed to take in the *original* model's outputs, as input. NBDT models are designed "
            "to only be used during validation and inference, not during training. Confused?  Check "
            "out github.com/alvinwan/nbdt#convert-neural-networks-to-decision-trees for examples "
            "and instructions."
        )

    def forward_tree(self, outputs, targets):
        raise NotImplementedError()

    def get_weight(self, start, end, power=1):
        progress = self.progress ** power
        return (1 - progress) * start + progress * end

    def current_weights(self):
        """(xent weight, tree weight) at the current training progress."""
        tree_weight = self.get_weight(
            self.tree_supervision_weight, self.tree_supervision_weight_end, self.tree_supervision_weight_power
        )
        xent_weight = self.get_weight(self.xent_weight, self.xent_weight_end, self.xent_weight_power)
        return xent_weight, tree_weight

    def forward(self, outputs, targets):
        xent_weight, tree_weight = self.current_weights()
        if self.fused_kind and _is_plain_cross_entropy(self.criterion) and outputs.dim() == 2:
            self.assert_output_not_nbdt(outputs)
            return ops.FusedTreeLossFunction.apply(
                outputs, targets, self.tree, self.fused_kind, xent_weight, tree_weight,
                self.tree_supervision_weight, self.batch_rows,
            )
        loss_xent = self.criterion(outputs, targets)
        loss_tree = self.forward_tree(outputs, targets)
        return loss_xent * xent_weight + loss_tree * tree_weight

    def set_epoch(self, cur, total):
        self.epochs = cur
        self.progress = cur / total
        if hasattr(super(), "set_epoch"):
            super().set_epoch(cur, total)


class HardTreeSupLoss(TreeSupLoss):
    """Cross-entropy at every inner node on the target's root-to-leaf path, down-weighted so that
    on average a sample draws half of its loss from the plain cross-entropy and half from the
    nodes (and, like the reference, scaled by `tree_supervision_weight` twice: once here, once
    in `TreeSupLoss.forward`)."""

    fused_kind = "hard"

    def forward_tree(self, outputs, targets):
        self.assert_output_not_nbdt(outputs)
        if not _is_plain_cross_entropy(self.criterion):
            raise NotImplementedError(
                "HardTreeSupLoss on the B200 path supports criterion=nn.CrossEntropyLoss() (default "
                "arguments); other criteria have no kernel and there is no PyTorch fallback."
            )
        # tree term alone: xent weight 0, outer tree weight 1
        return ops.FusedTreeLossFunction.apply(
            outputs, targets, self.tree, "hard", 0.0, 1.0, self.tree_supervision_weight, self.batch_rows
        )


class SoftTreeSupLoss(TreeSupLoss):
    """criterion(class probabilities of the soft decision rules, targets) -- the probabilities are
    handed to the criterion as if they were logits, exactly as the reference does."""

    fused_kind = "soft"

    def __init__(self, *args, Rules=None, **kwargs):
        super().__init__(*args, Rules=SoftEmbeddedDecisionRules, **kwargs)

    def forward_tree(self, outputs, targets):
        self.assert_output_not_nbdt(outputs)
        if _is_plain_cross_entropy(self.criterion) and outputs.dim() == 2:
            return ops.FusedTreeLossFunction.apply(
                outputs, targets, self.tree, "soft", 0.0, 1.0, 1.0, self.batch_rows
            )
        return self.criterion(self.rules(outputs), targets)


class SoftTreeLoss(SoftTreeSupLoss):
    """SoftTreeSupLoss whose hierarchy is re-induced from the live FC weights every few epochs."""

    accepts_tree_start_epochs = True
    accepts_tree_update_every_epochs = True
    accepts_tree_update_end_epochs = True
    accepts_arch = True
    accepts_net = lambda net, **kwargs: net
    accepts_checkpoint_path = lambda checkpoint_path, **kwargs: checkpoint_path

    fused_kind = None  # forward_tree switches between plain CE and the tree term by epoch

    def __init__(self, *args, arch=None, checkpoint_path="./", net=None, tree_start_epochs=67,
                 tree_update_every_epochs=10, tree_update_end_epochs=120, **kwargs):
        super().__init__(*args, **kwargs)
        self.start_epochs = tree_start_epochs
        self.update_every_epochs = tree_update_every_epochs
        self.update_end_epochs = tree_update_end_epochs
        self.net = net
        self.arch = arch
        self.checkpoint_path = checkpoint_path

    def forward_tree(self, outputs, targets):
        if self.epochs < self.start_epochs:
            return self.criterion(outputs, targets)  # regular xent
        return super().forward_tree(outputs, targets)

    def set_epoch(self, *args, **kwargs):
        super().set_epoch(*args, **kwargs)
        offset = self.epochs - self.start_epochs
        if offset >= 0 and offset % self.update_every_epochs == 0 and self.epochs < self.update_end_epochs:
            checkpoint_dir = self.checkpoint_path.replace(".pth", "")
            path_graph = os.path.join(checkpoint_dir, f"graph-epoch{self.epochs}.json")
            self.tree.update_from_model(self.net, self.arch, self.tree.dataset, path_graph=path_graph)


class SoftSegTreeSupLoss(SoftTreeSupLoss):
    """Segmentation variant: (N,C,H,W) logits, per-pixel rules; un-scheduled tree weight."""

    def forward(self, outputs, targets):
        self.assert_output_not_nbdt(outputs)
        if _is_plain_cross_entropy(self.criterion) and outputs.dim() == 4 and outputs.is_cuda:
            # criterion(outputs) + tsw * criterion(bayesian) is the fused soft loss with one row per pixel, on the (N, C, H, W)
            # tensor in place (no coerce / uncoerce copies, no materialised class probabilities)
            return ops.FusedSegTreeLossFunction.apply(outputs, targets, self.tree, 1.0, float(self.tree_supervision_weight),
                                                      self.batch_rows)
        loss = self.criterion(outputs, targets)
        bayesian = uncoerce_tensor(self.rules(coerce_tensor(outputs)), outputs.shape)
        return loss + self.criterion(bayesian, targets) * self.tree_supervision_weight
