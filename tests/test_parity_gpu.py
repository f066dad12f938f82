"""Parity of the B200 path with vectors the EXECUTING reference produced (tests/golden/make_golden.py, make_golden_extra.py):
whole networks (ResNet18 BasicBlock, ResNet50 Bottleneck) through one train step, get_node_logits_filtered, the segmentation
variants.  Backbone activations are bf16 (fp32 accumulation): north_star's bar for bf16 is 1e-2; what is measured on a B200
is printed by scripts/measure_parity.py and recorded in DESIGN.md §2."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def train_step(arch, golden):
    """Same seed, same batch, same loss as tests/golden/make_golden*.py; returns what the golden file holds."""
    from nbdt import models
    from nbdt.loss import SoftTreeSupLoss
    from nbdt.tree import Tree

    g = np.load(os.path.join(GOLDEN, golden))
    torch.manual_seed(0)
    net = getattr(models, arch)(num_classes=10)          # CPU init: the reference's parameter stream
    assert [k for k, _ in net.named_parameters()] == list(g["param_names"])
    net = net.to(DEV)
    crit = SoftTreeSupLoss(dataset=None, criterion=nn.CrossEntropyLoss(), tree=Tree("CIFAR10", hierarchy="induced-ResNet18"))
    x, y = torch.as_tensor(g["x"], device=DEV), torch.as_tensor(g["y"], device=DEV)
    net.train()
    logits = net(x)
    loss = crit(logits, y)
    loss.backward()
    return g, net, logits, loss, x


def autocast_floor(arch, g):
    """The same network in PyTorch eager mode under bf16 autocast (cuDNN / cuBLAS bf16 kernels, fp32 accumulation) on the same
    parameters and batch: what ANY implementation with bf16 activations gets against the fp32 golden -- the yardstick for ours."""
    from oracle import torch_port as P
    from oracle.tree_oracle import OracleTree

    torch.manual_seed(0)
    net = (P.EagerResNet50 if arch == "ResNet50" else P.EagerResNet)(num_classes=10).to(DEV)
    pkg = os.path.join(ROOT, "neural-backed-decision-trees_b200", "nbdt")
    tree = OracleTree(os.path.join(pkg, "hierarchies/CIFAR10/graph-induced-ResNet18.json"), os.path.join(pkg, "wnids/CIFAR10.txt"))
    x, y = torch.as_tensor(g["x"], device=DEV), torch.as_tensor(g["y"], device=DEV)
    net.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = net(x).float()
    loss = P.soft_tree_sup_loss(tree, nn.CrossEntropyLoss(), logits.cpu(), y.cpu())
    return logits.detach(), float(loss.detach())


@pytest.mark.parametrize("arch,golden", [("ResNet18", "backbone_resnet18.npz"), ("ResNet50", "backbone_resnet50.npz")])
def test_network_train_step_matches_the_reference_golden(arch, golden):
    """Direct comparison with the vectors of the executing reference (fp32, CPU).  Bars: north_star's 1e-2 for bf16 where bf16
    itself allows it (loss; gradients of the classifier), otherwise "no worse than 2x what eager bf16 autocast gets on the same
    network" -- batch-statistics BatchNorm over a batch of 8 re-normalises every layer's bf16 rounding, and through 20 (53) random-init
    layers that alone costs 1.8e-2 (2.7e-1) of the logits' scale for cuDNN's kernels too; measured here 2.8e-2 (2.6e-1)
    (profiles/r02_parity_measured.json)."""
    g, net, logits, loss, x = train_step(arch, golden)
    floor_logits, floor_loss = autocast_floor(arch, g)
    e_ours, e_floor = rel(logits.detach().cpu(), g["train_logits"]), rel(floor_logits.cpu(), g["train_logits"])
    assert e_ours <= max(1e-2, 2.0 * e_floor), (e_ours, e_floor)
    l_ours = abs(loss.item() - float(g["train_loss"])) / abs(float(g["train_loss"]))
    l_floor = abs(floor_loss - float(g["train_loss"])) / abs(float(g["train_loss"]))
    assert l_ours <= max(1e-2, 2.0 * l_floor), (l_ours, l_floor)
    norms = np.array([float(p.grad.double().norm()) for p in net.parameters()])
    ratio = norms / np.maximum(g["grad_norms"], 1e-12)
    big = g["grad_norms"] > 1e-3 * g["grad_norms"].max()
    assert abs(np.median(ratio[big]) - 1) <= 3e-2, np.median(ratio[big])
    assert np.all(np.abs(ratio[big] - 1) <= (0.1 if arch == "ResNet18" else 0.3)), (ratio[big].min(), ratio[big].max())
    cos = lambda a, b: float((a.flatten() @ b.flatten()) / (a.norm() * b.norm()))  # noqa: E731
    assert cos(net.linear.weight.grad.double().cpu(), torch.as_tensor(g["grad_linear_weight"]).double()) > 0.99
    if "eval_logits" in g.files:  # eval-mode forward (running statistics: no batch-statistics amplification) with the same parameters
        net.eval()
        with torch.no_grad():
            assert rel(net(x).cpu(), g["eval_logits"]) <= 3e-2


def test_get_node_logits_filtered_values(structure):
    """Values, not only shapes: per inner node the selector, the rows' node logits and the child targets the reference computed."""
    from nbdt.model import HardEmbeddedDecisionRules
    from nbdt.tree import Tree

    g = np.load(os.path.join(GOLDEN, "filtered_cifar100.npz"))
    tree = Tree("CIFAR100", hierarchy="induced")
    assert [n.wnid for n in tree.inodes] == list(g["inodes"])
    x, y = torch.as_tensor(g["logits"], device=DEV), torch.as_tensor(g["targets"], device=DEV)
    for i, node in enumerate(tree.inodes):
        sel, sub, tgt = HardEmbeddedDecisionRules.get_node_logits_filtered(node, x, y)
        assert list(sel) == [bool(v) for v in g[f"sel_{i}"]]
        assert list(tgt) == list(g[f"tgt_{i}"])
        assert tuple(sub.shape) == g[f"sub_{i}"].shape
        if sub.numel():
            # child order under a node is deterministic here and set-ordered in the reference: compare as multisets per row
            np.testing.assert_allclose(np.sort(sub.cpu().numpy(), axis=1), np.sort(g[f"sub_{i}"], axis=1), rtol=1e-5, atol=1e-6)


def test_segmentation_variants_match_the_reference_golden():
    """SoftSegNBDT / HardSegNBDT (reference model.py:376-399) and SoftSegTreeSupLoss (loss.py:318-327: un-scheduled weight, plain +=)
    on (N, C, H, W) logits with the Cityscapes hierarchy."""
    from nbdt.loss import SoftSegTreeSupLoss
    from nbdt.model import HardSegNBDT, SoftSegNBDT
    from nbdt.tree import Tree

    g = np.load(os.path.join(GOLDEN, "seg_cityscapes.npz"))
    pkg = os.path.join(ROOT, "neural-backed-decision-trees_b200", "nbdt")
    pg = os.path.join(pkg, "hierarchies", "Cityscapes", "graph-induced-HRNet-w18-v1.json")
    pw = os.path.join(GOLDEN, "wnids_cityscapes.txt")
    classes = [f"c{i}" for i in range(19)]
    x = torch.as_tensor(g["logits"], device=DEV)
    y = torch.as_tensor(g["targets"], device=DEV)
    kw = dict(path_graph=pg, path_wnids=pw, classes=classes)
    soft = SoftSegNBDT("Cityscapes", nn.Identity(), **kw)(x)
    np.testing.assert_allclose(soft.cpu().numpy(), g["soft_seg"], rtol=2e-5, atol=1e-8)
    hard = HardSegNBDT("Cityscapes", nn.Identity(), **kw)(x)
    assert np.array_equal(hard.cpu().numpy(), g["hard_seg"])  # one-hot per pixel: bit-exact
    tree = Tree("Cityscapes", path_graph=pg, path_wnids=pw, classes=classes)
    for tsw in (1.0, 0.5):
        xx = x.clone().requires_grad_(True)
        loss = SoftSegTreeSupLoss(dataset=None, criterion=nn.CrossEntropyLoss(), tree=tree, tree_supervision_weight=tsw)(xx, y)
        loss.backward()
        assert abs(loss.item() - float(g[f"seg_loss_tsw{tsw}"])) <= 2e-5 * abs(float(g[f"seg_loss_tsw{tsw}"]))
        want = g[f"seg_grad_tsw{tsw}"]
        np.testing.assert_allclose(xx.grad.cpu().numpy(), want, rtol=2e-4, atol=2e-6 * np.abs(want).max())


@pytest.mark.parametrize("shape", [(2, 19, 8, 8), (3, 19, 5, 7), (1, 19, 32, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_class_major_kernels_equal_the_permuted_row_kernels(shape, dtype):
    """The segmentation wrappers address (N, C, H, W) tensors in place (nbdt_*_nchw): same numbers as the reference's route --
    permute + reshape to rows, rules / loss on the rows, reshape + permute back -- through the row-major kernels; the run-time
    specialised class-major kernel is the one that runs when H*W is a multiple of 32, the table-driven one otherwise."""
    from nbdt._b200 import lib as L
    from nbdt._b200 import ops
    from nbdt.tree import Tree
    from nbdt.utils import coerce_tensor, uncoerce_tensor

    pkg = os.path.join(ROOT, "neural-backed-decision-trees_b200", "nbdt")
    pg = os.path.join(pkg, "hierarchies", "Cityscapes", "graph-induced-HRNet-w18-v1.json")
    pw = os.path.join(GOLDEN, "wnids_cityscapes.txt")
    tree = Tree("Cityscapes", path_graph=pg, path_wnids=pw, classes=[f"c{i}" for i in range(19)])
    N, C, H, W = shape
    torch.manual_seed(N * H + W)
    x = (2 * torch.randn(shape, device=DEV)).to(dtype)
    y = torch.randint(C, (N, H, W), device=DEV)
    rows = coerce_tensor(x).contiguous()
    lp = dtype != torch.float32

    soft = ops.soft_forward_nchw(tree, x)
    name = L.load().nbdt_tree_kernel_variant(tree.device_tree(torch.device(DEV)).handle).decode()
    assert ("_nchw" in name) == ((H * W) % 32 == 0), name
    want = uncoerce_tensor(ops.soft_forward(tree, rows), x.shape)
    torch.testing.assert_close(soft, want, rtol=2e-5, atol=1e-7)

    pred = ops.hard_forward_nchw(tree, x)
    want_pred, _ = ops.hard_forward(tree, rows)
    assert torch.equal(pred.reshape(-1), want_pred)

    for w_tree in (1.0, 0.5):
        loss, grad = ops.tree_loss_nchw(tree, x, y, 1.0, w_tree)
        name = L.load().nbdt_tree_kernel_variant(tree.device_tree(torch.device(DEV)).handle).decode()
        assert ("_nchw" in name) == ((H * W) % 32 == 0), name
        want_l, want_g = ops.tree_loss(tree, rows, y.reshape(-1), "soft", 1.0, w_tree)
        assert abs(loss.item() - want_l.item()) <= 2e-5 * abs(want_l.item())
        want_g = uncoerce_tensor(want_g, x.shape).float()
        torch.testing.assert_close(grad.float(), want_g, rtol=1e-2 if lp else 2e-4, atol=(8e-3 if lp else 2e-6) * want_g.abs().max().item())
