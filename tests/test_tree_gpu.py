"""Parity of the fused tree-head / tree-loss CUDA kernels (through the C ABI) with
 (a) the golden vectors produced by the executing reference, and
 (b) the pinned oracle (oracle/tree_oracle.py) on seeded inputs,
plus size-independent properties at BASELINE.json's full sizes.

Tolerances (north_star): indices bit-exact; fp32 logits/loss within 1e-3 rel (we assert 2e-5 or
tighter), bf16 within 1e-2.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import TREE_CASES, case_paths, load_golden, write_random_hierarchy
from oracle import tree_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
RTOL, ATOL = 2e-5, 2e-7          # fp32 inputs
RTOL_BF16, ATOL_BF16 = 1e-2, 1e-4


def make_tree(dataset=None, hierarchy=None, path_graph=None, path_wnids=None, classes=None, group=None):
    """Fresh product Tree; `group` forces the lanes-per-row variant of the kernel (read at tree upload)."""
    from nbdt.tree import Tree

    if group is None:
        os.environ.pop("NBDT_B200_TREE_G", None)
    else:
        os.environ["NBDT_B200_TREE_G"] = str(group)
    try:
        tree = Tree(dataset, path_graph=path_graph, path_wnids=path_wnids, classes=classes, hierarchy=hierarchy)
        tree.device_tree(torch.device(DEV))  # upload now, while the env var is set
    finally:
        os.environ.pop("NBDT_B200_TREE_G", None)
    return tree


def case_tree(structure, case, group=None):
    s = structure[case]
    return make_tree(s["dataset"], s["hierarchy"], group=group)


def t(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), device=DEV).to(dtype)


def close(got, want, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(got.detach().double().cpu().numpy(), want, rtol=rtol, atol=atol)


GROUPS = {"cifar10_induced": [None, 1, 2, 4, 32], "cifar100_wordnet": [None, 1, 8, 16], "tiny200_induced_resnet18": [None, 4, 32],
          "imagenet1000_induced": [None, 16]}


def case_groups():
    out = []
    for case in TREE_CASES:
        for g in GROUPS.get(case, [None]):
            out.append((case, g))
    return out


@pytest.mark.parametrize("case,group", case_groups())
def test_golden_soft_and_hard_forward(case, group, structure):
    from nbdt._b200 import ops

    g = load_golden(case)
    tree = case_tree(structure, case, group)
    x = t(g["logits"])
    close(ops.soft_forward(tree, x), g["soft_probs_f64in"], rtol=1e-5, atol=1e-9)
    close(ops.soft_forward(tree, x), g["soft_probs"], rtol=3e-5, atol=1e-8)

    pred, paths = ops.hard_forward(tree, x, with_paths=True)
    assert torch.equal(pred.cpu(), torch.as_tensor(g["hard_pred"]))           # bit-exact indices
    d = g["hard_path_child"].shape[1]
    assert torch.equal(paths[0][:, :d].cpu().long(), torch.as_tensor(g["hard_path_child"]))
    assert torch.equal(paths[1][:, :d].cpu().long(), torch.as_tensor(g["hard_path_inode"]))
    assert (paths[0][:, d:] == -1).all()
    close(paths[2][:, :d], g["hard_path_prob"], rtol=1e-5, atol=1e-6)
    close(paths[3][:, :d], g["hard_path_entropy"], rtol=1e-4, atol=5e-6)
    pred2, none = ops.hard_forward(tree, x, with_paths=False)
    assert none is None and torch.equal(pred, pred2)


@pytest.mark.parametrize("case", TREE_CASES)
def test_golden_node_outputs(case, structure):
    from nbdt._b200 import ops

    g = load_golden(case)
    tree = case_tree(structure, case)
    nl, pr, preds, ent, dt = ops.node_outputs(tree, t(g["logits"]))
    assert torch.equal(preds.cpu(), torch.as_tensor(g["node_preds"]))
    for i in range(dt.num_inodes):
        b, e = int(dt.child_ptr[i]), int(dt.child_ptr[i + 1])
        close(nl[:, b:e], g["node_logits_f64"][:, i, : e - b], rtol=1e-5, atol=2e-6)
        close(pr[:, b:e], g["node_probs_f64"][:, i, : e - b], rtol=1e-5, atol=1e-7)
    close(ent, g["node_entropy_f64"], rtol=1e-4, atol=5e-6)


LOSS_TAGS = [("default", 1.0, 1.0), ("tsw10", 1.0, 10.0), ("xw05_tsw2", 0.5, 2.0)]


@pytest.mark.parametrize("case,group", case_groups())
@pytest.mark.parametrize("tag,xw,tsw", LOSS_TAGS)
def test_golden_losses(case, group, tag, xw, tsw, structure):
    from nbdt.loss import HardTreeSupLoss, SoftTreeSupLoss

    g = load_golden(case)
    tree = case_tree(structure, case, group)
    y = t(g["targets"], torch.int64)
    for name, cls in (("soft", SoftTreeSupLoss), ("hard", HardTreeSupLoss)):
        crit = cls(dataset=None, criterion=nn.CrossEntropyLoss(), tree=tree, tree_supervision_weight=tsw, xent_weight=xw)
        x = t(g["logits"]).requires_grad_(True)
        loss = crit(x, y)
        loss.backward()
        want_l, want_g = g[f"{name}_loss_{tag}_f64"], g[f"{name}_grad_{tag}_f64"]
        assert abs(loss.item() - want_l) <= 1e-5 * abs(want_l)
        scale = np.abs(want_g).max()
        close(x.grad, want_g, rtol=2e-4, atol=2e-6 * scale)
        # forward only (no grad buffer) gives the same loss
        with torch.no_grad():
            assert abs(crit(t(g["logits"]), y).item() - want_l) <= 1e-5 * abs(want_l)


def test_golden_weight_schedule(structure):
    from nbdt.loss import SoftTreeSupLoss

    for case in ("cifar10_induced", "cifar100_wordnet"):
        g = load_golden(case)
        crit = SoftTreeSupLoss(dataset=None, criterion=nn.CrossEntropyLoss(), tree=case_tree(structure, case),
                               tree_supervision_weight=1.0, tree_supervision_weight_end=5.0,
                               tree_supervision_weight_power=2, xent_weight=1.0, xent_weight_end=0.25)
        crit.set_epoch(3, 10)
        x = t(g["logits"]).requires_grad_(True)
        loss = crit(x, t(g["targets"], torch.int64))
        loss.backward()
        assert abs(loss.item() - g["soft_loss_sched_f64"]) <= 1e-5 * abs(g["soft_loss_sched_f64"])
        close(x.grad, g["soft_grad_sched_f64"], rtol=2e-4, atol=2e-6 * np.abs(g["soft_grad_sched_f64"]).max())


@pytest.mark.parametrize("case", ["cifar10_induced", "cifar100_wordnet", "imagenet1000_induced"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_low_precision_logits(case, dtype, structure):
    """bf16/fp16 logits: arithmetic is fp32, so the result must match the oracle evaluated on the
    ROUNDED inputs; only the gradient is rounded back to the input dtype."""
    from nbdt._b200 import ops

    g = load_golden(case)
    tree = case_tree(structure, case)
    ot = O.OracleTree(*case_paths(structure, case))
    xl = t(g["logits"], dtype)
    xr = xl.float().cpu().double().numpy()
    y = g["targets"]
    close(ops.soft_forward(tree, xl), O.soft_probs(ot, xr), rtol=1e-5, atol=1e-9)
    pred, _ = ops.hard_forward(tree, xl)
    assert torch.equal(pred.cpu(), torch.as_tensor(O.hard_decisions(ot, xr)[0]))
    for kind, fn in (("soft", O.soft_tree_sup_loss), ("hard", O.hard_tree_sup_loss)):
        loss, grad = ops.tree_loss(tree, xl, t(y, torch.int64), kind, 1.0, 1.0, 1.0)
        want_l, want_g = fn(ot, xr, y)
        assert abs(loss.item() - want_l) <= 1e-5 * abs(want_l)
        assert grad.dtype == dtype
        close(grad, want_g, rtol=RTOL_BF16, atol=RTOL_BF16 * np.abs(want_g).max() * 0.5)


RANDOM_TREES = [  # (leaves, max children, chain bias, rows, group)
    (2, 2, 0.0, 7, None), (3, 3, 0.0, 33, None), (5, 2, 1.0, 1, None), (17, 4, 0.5, 65, None), (17, 4, 0.5, 65, 1),
    (64, 13, 0.2, 40, None), (64, 2, 0.9, 40, 2), (257, 5, 0.3, 19, None), (257, 2, 0.95, 19, 32), (1000, 8, 0.5, 9, None),
]


@pytest.mark.parametrize("leaves,kmax,chain,rows,group", RANDOM_TREES)
def test_random_hierarchies_match_oracle(leaves, kmax, chain, rows, group, tmp_path):
    """Variable arity (up to 13), deep unbalanced chains, ragged row counts."""
    from nbdt._b200 import ops

    pg, pw = write_random_hierarchy(tmp_path, leaves, kmax, seed=leaves * 31 + kmax, chain_bias=chain)
    classes = [f"c{i}" for i in range(leaves)]
    tree = make_tree(None, path_graph=pg, path_wnids=pw, classes=classes, group=group)
    ot = O.OracleTree(pg, pw)
    assert [n.wnid for n in tree.inodes] == ot.inodes
    rng = np.random.default_rng(leaves)
    x = (3 * rng.standard_normal((rows, leaves))).astype(np.float32)
    y = rng.integers(0, leaves, rows)
    xd = x.astype(np.float64)
    close(ops.soft_forward(tree, t(x)), O.soft_probs(ot, xd), rtol=2e-5, atol=1e-9)
    pred, paths = ops.hard_forward(tree, t(x), with_paths=True)
    want_pred, want_paths = O.hard_decisions(ot, xd)
    nodes = O.node_outputs(ot, xd)
    margin = min(np.sort(nodes[w]["logits"], axis=1)[b, -1] - np.sort(nodes[w]["logits"], axis=1)[b, -2]
                 for b, path in enumerate(want_paths) for (w, _, _, _) in path)
    assert margin > 1e-5, "test input has a near-tie; pick another seed"
    assert torch.equal(pred.cpu(), torch.as_tensor(want_pred))
    for b, path in enumerate(want_paths):
        assert paths[0][b, : len(path)].tolist() == [k for _, k, _, _ in path]
        assert (paths[0][b, len(path):] == -1).all()
    for kind, fn in (("soft", O.soft_tree_sup_loss), ("hard", O.hard_tree_sup_loss)):
        loss, grad = ops.tree_loss(tree, t(x), t(y, torch.int64), kind, 0.7, 1.3, 2.0)
        kw = {"tree_supervision_weight": 2.0} if kind == "hard" else {}
        want_l, want_g = fn(ot, xd, y, xent_weight=0.7, tree_weight=1.3, **kw)
        assert abs(loss.item() - want_l) <= 2e-5 * abs(want_l)
        close(grad, want_g, rtol=2e-4, atol=2e-6 * np.abs(want_g).max())


def test_edge_cases(structure):
    from nbdt._b200 import ops

    tree = case_tree(structure, "cifar100_induced")
    ot = O.OracleTree(*case_paths(structure, "cifar100_induced"))
    # empty batch
    x0 = torch.empty((0, 100), device=DEV)
    assert ops.soft_forward(tree, x0).shape == (0, 100)
    assert ops.hard_forward(tree, x0)[0].shape == (0,)
    loss, grad = ops.tree_loss(tree, x0, torch.empty((0,), dtype=torch.int64, device=DEV), "soft", 1, 1)
    assert loss.item() == 0.0 and grad.shape == (0, 100)
    # single row / row stride larger than C (a column slice of a wider matrix) / odd row counts
    rng = np.random.default_rng(5)
    wide = t(3 * rng.standard_normal((37, 128)))
    view = wide[:, 11:111]
    assert view.stride(0) == 128
    xd = view.double().cpu().numpy()
    y = rng.integers(0, 100, 37)
    close(ops.soft_forward(tree, view), O.soft_probs(ot, xd), rtol=2e-5, atol=1e-9)
    loss, grad = ops.tree_loss(tree, view, t(y, torch.int64), "soft", 1, 1)
    want_l, want_g = O.soft_tree_sup_loss(ot, xd, y)
    assert abs(loss.item() - want_l) <= 2e-5 * abs(want_l)
    close(grad, want_g, rtol=2e-4, atol=2e-6 * np.abs(want_g).max())
    close(ops.soft_forward(tree, view[:1]), O.soft_probs(ot, xd[:1]), rtol=2e-5, atol=1e-9)
    # extreme logits: no NaN/inf, probabilities still a distribution
    big = t(np.where(rng.random((8, 100)) < 0.5, 80.0, -80.0))
    p = ops.soft_forward(tree, big)
    assert torch.isfinite(p).all() and torch.allclose(p.sum(1), torch.ones(8, device=DEV), atol=1e-5)
    loss, grad = ops.tree_loss(tree, big, t(rng.integers(0, 100, 8), torch.int64), "soft", 1, 1)
    assert torch.isfinite(loss) and torch.isfinite(grad).all()
    # wrong class count / CPU tensor are rejected
    with pytest.raises(ValueError):
        ops.soft_forward(tree, torch.zeros((2, 99), device=DEV))


FULL_SIZES = [  # BASELINE.json configs 2-5 head shapes + a segmentation-sized row sweep
    ("cifar10_induced", 512, torch.bfloat16), ("cifar100_induced", 512, torch.bfloat16),
    ("tiny200_induced_resnet18", 1024, torch.float32), ("imagenet1000_induced", 2048, torch.bfloat16),
    ("cifar10_induced", 1 << 20, torch.float32), ("lip_induced", 262144 + 3, torch.bfloat16),
]


@pytest.mark.parametrize("case,rows,dtype", FULL_SIZES)
def test_full_size_properties(case, rows, dtype, structure):
    """Size-independent properties at full size + the oracle on a random subset of rows
    (rows are independent; only the 1/B of the mean couples them)."""
    from nbdt._b200 import ops

    tree = case_tree(structure, case)
    ot = O.OracleTree(*case_paths(structure, case))
    C = ot.num_classes
    gen = torch.Generator(device=DEV).manual_seed(rows + C)
    x = (3 * torch.randn((rows, C), device=DEV, generator=gen)).to(dtype)
    y = torch.randint(C, (rows,), device=DEV, generator=gen)
    probs = ops.soft_forward(tree, x)
    assert torch.allclose(probs.sum(1), torch.ones(rows, device=DEV), atol=2e-5)      # a distribution
    assert (probs >= 0).all()
    # hard walk == walking the node-output kernel's per-node argmax on the host tables
    pred, paths = ops.hard_forward(tree, x, with_paths=True)
    sub = torch.randperm(rows, device=DEV, generator=gen)[:64]
    _, _, npreds, _, dt = ops.node_outputs(tree, x[sub])
    npreds = npreds.cpu().numpy()
    for j, b in enumerate(sub.tolist()):
        node = dt.num_classes + dt.root
        while node >= dt.num_classes:
            i = node - dt.num_classes
            node = int(dt.child_node[int(dt.child_ptr[i]) + int(npreds[j, i])])
        assert node == int(pred[b])
    # oracle on the subset
    xs = x[sub].float().double().cpu().numpy()
    ys = y[sub].cpu().numpy()
    close(probs[sub], O.soft_probs(ot, xs), rtol=2e-5, atol=1e-9)
    assert torch.equal(pred[sub].cpu(), torch.as_tensor(O.hard_decisions(ot, xs)[0]))
    lp = dtype != torch.float32
    for kind, fn in (("soft", O.soft_tree_sup_loss), ("hard", O.hard_tree_sup_loss)):
        loss, grad = ops.tree_loss(tree, x, y, kind, 1.0, 1.0, 1.0)
        _, want_g = fn(ot, xs, ys)
        want_g = want_g * (len(ys) / rows)                    # mean over `rows`, not over the subset
        tol = RTOL_BF16 if lp else 2e-4
        close(grad[sub], want_g, rtol=tol, atol=(0.5 * tol if lp else 2e-6) * np.abs(want_g).max())
        # linearity in the weights, and shards add up (multi-GPU convention: batch_rows = global rows)
        loss2, grad2 = ops.tree_loss(tree, x, y, kind, 2.0, 2.0, 1.0)
        assert abs(loss2.item() - 2 * loss.item()) <= 1e-5 * abs(loss2.item())
        h = rows // 2
        la, ga = ops.tree_loss(tree, x[:h], y[:h], kind, 1.0, 1.0, 1.0, batch_rows=rows)
        lb, gb = ops.tree_loss(tree, x[h:], y[h:], kind, 1.0, 1.0, 1.0, batch_rows=rows)
        assert abs(la.item() + lb.item() - loss.item()) <= 2e-5 * abs(loss.item())
        assert torch.equal(torch.cat([ga, gb]), grad)
    # fused soft loss == composition: CE(x) + CE(soft probs) with the differentiable rules kernel
    xr = x.float().detach().requires_grad_(True)
    comp = nn.functional.cross_entropy(xr, y) + nn.functional.cross_entropy(ops.SoftRulesFunction.apply(xr, tree), y)
    comp.backward()
    loss, grad = ops.tree_loss(tree, x.float(), y, "soft", 1.0, 1.0)
    assert abs(comp.item() - loss.item()) <= 2e-5 * abs(loss.item())
    assert torch.allclose(xr.grad, grad, rtol=1e-3, atol=2e-6 * grad.abs().max().item())


def test_generic_criterion_uses_differentiable_rules(structure):
    """A non-default criterion goes through SoftRulesFunction (kernel fwd + kernel bwd)."""
    from nbdt.loss import SoftTreeSupLoss

    g = load_golden("cifar100_wordnet")
    tree = case_tree(structure, "cifar100_wordnet")
    ot = O.OracleTree(*case_paths(structure, "cifar100_wordnet"))
    crit = SoftTreeSupLoss(dataset=None, criterion=nn.CrossEntropyLoss(reduction="sum"), tree=tree)
    x = t(g["logits"]).requires_grad_(True)
    y = g["targets"]
    loss = crit(x, t(y, torch.int64))
    loss.backward()
    B = len(y)
    want_l, want_g = O.soft_tree_sup_loss(ot, g["logits"].astype(np.float64), y)
    assert abs(loss.item() - B * want_l) <= 2e-5 * abs(B * want_l)
    close(x.grad, B * want_g, rtol=2e-4, atol=2e-6 * B * np.abs(want_g).max())


@pytest.mark.parametrize("case", ["cifar10_induced", "cifar100_wordnet", "tiny200_induced_resnet18", "imagenet1000_induced"])
def test_out_of_range_targets_are_ignored_rows(case, structure):
    """A target outside [0, C) (nn.CrossEntropyLoss's ignore_index = -100, or a corrupt label) must not index memory: the row
    contributes nothing to the loss and gets a zero gradient row; with batch_rows = the number of valid rows the result is the
    oracle's on the valid rows alone.  Every kernel family: specialised (AOT + run-time compiled) and table-driven."""
    from nbdt._b200 import ops

    g = load_golden(case)
    ot = O.OracleTree(*case_paths(structure, case))
    x = g["logits"][:16].astype(np.float32)
    n, C = x.shape   # the ImageNet fixture holds 8 rows
    y = g["targets"][:n].copy()
    bad = y.copy()
    rows = [0, n // 2, n - 1]
    bad[rows] = [-100, C, 1 << 40]
    keep = np.array([i for i in range(n) if i not in rows])
    for group in (None, 4):  # None: specialised kernel where there is one; 4: the table-driven kernel
        tree = case_tree(structure, case, group)
        for kind, fn in (("soft", O.soft_tree_sup_loss), ("hard", O.hard_tree_sup_loss)):
            want_l, want_g = fn(ot, x[keep].astype(np.float64), y[keep])
            loss, grad = ops.tree_loss(tree, t(x), t(bad, torch.int64), kind, 1.0, 1.0, 1.0, batch_rows=len(keep))
            assert abs(loss.item() - want_l) <= 2e-5 * abs(want_l)
            assert (grad[rows] == 0).all()
            close(grad[keep], want_g, rtol=2e-4, atol=2e-6 * np.abs(want_g).max())


@pytest.mark.parametrize("case", ["cifar100_induced", "tiny200_wordnet", "ade20k_induced", "imagenet1000_induced"])
def test_specialised_kernel_is_the_one_that_runs(case, structure):
    """The loss / soft-forward launches of a hierarchy with up to 1536 leaves go to its run-time specialised kernel (from the
    cubin cache the build populated, or compiled on the spot), not to the table-driven fallback."""
    from nbdt._b200 import lib as L
    from nbdt._b200 import ops

    g = load_golden(case)
    tree = case_tree(structure, case)
    ops.tree_loss(tree, t(g["logits"]), t(g["targets"], torch.int64), "soft", 1.0, 1.0)
    name = L.load().nbdt_tree_kernel_variant(tree.device_tree(torch.device(DEV)).handle).decode()
    assert name.startswith("ka") and "_m0_t0_" in name, name   # "ka<generator version>_<hierarchy hash>_m<mode>_t<dtype>_...": csrc/tree_jit.cu


def test_unseen_hierarchy_is_compiled_at_run_time(tmp_path):
    """A hierarchy nobody shipped (what SoftTreeLoss's re-induction produces) gets its own kernel: compile, cache, reuse."""
    from nbdt._b200 import lib as L
    from nbdt._b200 import ops

    pg, pw = write_random_hierarchy(tmp_path, 37, 3, seed=4242, chain_bias=0.4)
    ot = O.OracleTree(pg, pw)
    rng = np.random.default_rng(1)
    x = (3 * rng.standard_normal((50, 37))).astype(np.float32)
    y = rng.integers(0, 37, 50)
    for attempt in range(2):  # second tree object: same hash -> served from the on-disk cache
        tree = make_tree(None, path_graph=pg, path_wnids=pw, classes=[f"c{i}" for i in range(37)])
        loss, grad = ops.tree_loss(tree, t(x), t(y, torch.int64), "soft", 1.0, 1.0)
        want_l, want_g = O.soft_tree_sup_loss(ot, x.astype(np.float64), y)
        assert abs(loss.item() - want_l) <= 2e-5 * abs(want_l)
        close(grad, want_g, rtol=2e-4, atol=2e-6 * np.abs(want_g).max())
        assert L.load().nbdt_tree_kernel_variant(tree.device_tree(torch.device(DEV)).handle).decode().startswith("ka")
