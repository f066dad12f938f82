"""The oracle (oracle/tree_oracle.py) pinned against vectors produced by the executing reference
(tests/golden/make_golden.py), and the product's host-side Tree tables against the reference's."""
import numpy as np
import pytest

from conftest import TREE_CASES, case_paths, load_golden
from oracle import tree_oracle as O


@pytest.fixture(scope="module")
def oracle_trees(structure):
    return {case: O.OracleTree(*case_paths(structure, case)) for case in TREE_CASES}


@pytest.mark.parametrize("case", TREE_CASES)
def test_oracle_structure_matches_reference(case, structure, oracle_trees):
    s, t = structure[case], oracle_trees[case]
    assert t.wnids == s["wnids_leaves"]
    assert t.inodes == s["inodes"]
    assert t.nodes == s["nodes"]
    assert t.root == s["root"]
    for w in t.inodes:
        assert t.succ[w] == s["children"][w]
        assert t.child_classes[w] == s["child_index_to_class_index"][w]


@pytest.mark.parametrize("case", TREE_CASES)
def test_product_tree_matches_reference(case, structure):
    from nbdt.tree import Tree

    s = structure[case]
    tree = Tree(s["dataset"], hierarchy=s["hierarchy"])
    assert tree.wnids_leaves == s["wnids_leaves"]
    assert [n.wnid for n in tree.inodes] == s["inodes"]
    assert [n.wnid for n in tree.nodes] == s["nodes"]
    assert tree.root.wnid == s["root"]
    for node in tree.inodes:
        assert [c.wnid for c in node.children] == s["children"][node.wnid]
        got = [sorted(node.child_index_to_class_index[k]) for k in range(node.num_classes)]
        assert got == s["child_index_to_class_index"][node.wnid]
        for k, idx in enumerate(got):
            for c in idx:
                assert node.class_index_to_child_index[c] == [k]
    steps = tree.get_leaf_to_steps()
    for leaf, ref in s["leaf_to_steps"].items():
        assert [[st["node"].wnid, st["next_index"]] for st in steps[leaf]] == ref


@pytest.mark.parametrize("case", TREE_CASES)
def test_oracle_node_outputs_and_soft_probs(case, oracle_trees):
    g, t = load_golden(case), oracle_trees[case]
    x = g["logits"].astype(np.float64)
    nodes = O.node_outputs(t, x)
    for i, w in enumerate(t.inodes):
        k = len(t.succ[w])
        np.testing.assert_allclose(nodes[w]["logits"], g["node_logits_f64"][:, i, :k], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(nodes[w]["probs"], g["node_probs_f64"][:, i, :k], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(nodes[w]["entropy"], g["node_entropy_f64"][:, i], rtol=1e-9, atol=1e-12)
        assert (nodes[w]["preds"] == g["node_preds"][:, i]).all()
    cp = O.soft_probs(t, x, nodes)
    # the reference accumulates into a float32 buffer even for float64 inputs (model.py:226)
    np.testing.assert_allclose(cp, g["soft_probs_f64in"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(cp, g["soft_probs"], rtol=2e-5, atol=1e-8)
    np.testing.assert_allclose(cp.sum(1), 1.0, rtol=1e-12)


@pytest.mark.parametrize("case", TREE_CASES)
def test_oracle_hard_decisions(case, oracle_trees):
    g, t = load_golden(case), oracle_trees[case]
    pred, paths = O.hard_decisions(t, g["logits"].astype(np.float64))
    assert (pred == g["hard_pred"]).all()
    np.testing.assert_array_equal(O.hard_onehot(t, g["logits"]), g["hard_onehot"])
    index = {w: i for i, w in enumerate(t.inodes)}
    for b, path in enumerate(paths):
        d = len(path)
        assert (g["hard_path_child"][b, d:] == -1).all()
        for j, (w, k, p, e) in enumerate(path):
            assert g["hard_path_child"][b, j] == k
            assert g["hard_path_inode"][b, j] == index[w]
            # reference values were computed from float32 logits
            assert abs(p - g["hard_path_prob"][b, j]) < 2e-6
            assert abs(e - g["hard_path_entropy"][b, j]) < 5e-6


@pytest.mark.parametrize("case", TREE_CASES)
@pytest.mark.parametrize("tag,kw", [("default", {}), ("tsw10", {"tsw": 10.0}), ("xw05_tsw2", {"xw": 0.5, "tsw": 2.0})])
def test_oracle_losses(case, tag, kw, oracle_trees):
    g, t = load_golden(case), oracle_trees[case]
    x, y = g["logits"].astype(np.float64), g["targets"]
    xw, tsw = kw.get("xw", 1.0), kw.get("tsw", 1.0)
    # default schedule: progress = 1 -> tree weight = tsw (end defaults to start)
    ls, gs = O.soft_tree_sup_loss(t, x, y, xent_weight=xw, tree_weight=tsw)
    np.testing.assert_allclose(ls, g[f"soft_loss_{tag}_f64"], rtol=1e-6)
    np.testing.assert_allclose(gs, g[f"soft_grad_{tag}_f64"], rtol=2e-5, atol=2e-8)
    lh, gh = O.hard_tree_sup_loss(t, x, y, xent_weight=xw, tree_weight=tsw, tree_supervision_weight=tsw)
    np.testing.assert_allclose(lh, g[f"hard_loss_{tag}_f64"], rtol=1e-10)
    np.testing.assert_allclose(gh, g[f"hard_grad_{tag}_f64"], rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("case", ["cifar10_induced", "cifar100_wordnet"])
def test_oracle_weight_schedule(case, oracle_trees):
    g, t = load_golden(case), oracle_trees[case]
    wt = O.schedule_weight(1.0, 5.0, 2, 3 / 10)
    wx = O.schedule_weight(1.0, 0.25, 1, 3 / 10)
    loss, grad = O.soft_tree_sup_loss(t, g["logits"].astype(np.float64), g["targets"], xent_weight=wx, tree_weight=wt)
    np.testing.assert_allclose(loss, g["soft_loss_sched_f64"], rtol=1e-6)
    np.testing.assert_allclose(grad, g["soft_grad_sched_f64"], rtol=2e-5, atol=2e-8)


# ---- oracle/tree_oracle.c (plain-C restatement; bench_tree.py's cpu_baseline) pinned to the same golden vectors ----
@pytest.mark.parametrize("case", TREE_CASES)
def test_c_oracle_pinned_to_reference(case, oracle_trees):
    from oracle import tree_oracle_c as OC

    g, t = load_golden(case), oracle_trees[case]
    ct = OC.CTree(t)
    x, y = g["logits"], g["targets"]  # float32 logits, as the reference saw them
    cp = OC.soft_probs(ct, x)
    np.testing.assert_allclose(cp, g["soft_probs_f64in"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(cp, O.soft_probs(t, x), rtol=1e-12, atol=1e-15)
    for tag, xw, tsw in (("default", 1.0, 1.0), ("tsw10", 1.0, 10.0), ("xw05_tsw2", 0.5, 2.0)):
        ls, gs = OC.soft_tree_sup_loss(ct, x, y, xw, tsw)
        np.testing.assert_allclose(ls, g[f"soft_loss_{tag}_f64"], rtol=1e-6)
        np.testing.assert_allclose(gs, g[f"soft_grad_{tag}_f64"], rtol=2e-5, atol=2e-8)
        lh, gh = OC.hard_tree_sup_loss(ct, x, y, xw, tsw, tsw)
        np.testing.assert_allclose(lh, g[f"hard_loss_{tag}_f64"], rtol=1e-10)
        np.testing.assert_allclose(gh, g[f"hard_grad_{tag}_f64"], rtol=1e-9, atol=1e-13)
    # loss only (no gradient buffer)
    assert OC.soft_tree_sup_loss(ct, x, y, want_grad=False)[1] is None


@pytest.mark.parametrize("leaves,kmax,seed,chain", [(7, 2, 0, 0.0), (33, 5, 1, 0.5), (100, 13, 2, 0.0), (257, 3, 3, 0.9)])
def test_c_oracle_equals_numpy_oracle_on_random_hierarchies(tmp_path, leaves, kmax, seed, chain):
    """Ragged / deep / wide hierarchies the golden set does not contain: the two restatements must agree to round-off."""
    from conftest import write_random_hierarchy
    from oracle import tree_oracle_c as OC

    t = O.OracleTree(*write_random_hierarchy(tmp_path, leaves, kmax, seed, chain_bias=chain))
    ct = OC.CTree(t)
    rng = np.random.default_rng(seed)
    x = (4 * rng.standard_normal((37, leaves))).astype(np.float32)
    y = rng.integers(0, leaves, 37)
    np.testing.assert_allclose(OC.soft_probs(ct, x), O.soft_probs(t, x), rtol=1e-11, atol=1e-300)
    for fc, fn, kw in ((OC.soft_tree_sup_loss, O.soft_tree_sup_loss, {}),
                       (OC.hard_tree_sup_loss, O.hard_tree_sup_loss, {"tree_supervision_weight": 3.0})):
        lc, gc = fc(ct, x, y, 0.7, 3.0, *kw.values())
        ln, gn = fn(t, x, y, xent_weight=0.7, tree_weight=3.0, **kw)
        np.testing.assert_allclose(lc, ln, rtol=1e-12)
        np.testing.assert_allclose(gc, gn, rtol=1e-9, atol=1e-15)


# ---- oracle/torch_port.py (the CPU baseline bench.py times) pinned to the same golden vectors ----
@pytest.mark.parametrize("case", ["cifar10_induced", "cifar100_wordnet", "tiny200_induced_resnet18"])
def test_torch_port_tree_losses(case, oracle_trees):
    import torch
    import torch.nn as nn
    from oracle import torch_port as P

    g, t = load_golden(case), oracle_trees[case]
    y = torch.as_tensor(g["targets"])
    np.testing.assert_allclose(P.soft_rules(t, torch.as_tensor(g["logits"])).numpy(), g["soft_probs"], rtol=1e-6, atol=1e-9)
    for name, fn, kw in (("soft", P.soft_tree_sup_loss, {}), ("hard", P.hard_tree_sup_loss, {"tsw": 10.0})):
        x = torch.as_tensor(g["logits"]).double().requires_grad_(True)
        loss = fn(t, nn.CrossEntropyLoss(), x, y, 1.0, 10.0, **kw)
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"{name}_loss_tsw10_f64"], rtol=1e-9)
        np.testing.assert_allclose(x.grad.numpy(), g[f"{name}_grad_tsw10_f64"], rtol=1e-7, atol=1e-12)


def test_torch_port_resnet18_train_step(oracle_trees):
    """Seeded eager ResNet18 + SoftTreeSupLoss + SGD reproduces the executing reference's step."""
    import os
    import torch
    import torch.nn as nn
    from conftest import GOLDEN
    from oracle import torch_port as P

    g = np.load(os.path.join(GOLDEN, "backbone_resnet18.npz"))
    torch.manual_seed(0)
    net = P.EagerResNet(num_classes=10)
    assert [k for k, _ in net.named_parameters()] == list(g["param_names"])
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    tree = oracle_trees["cifar10_induced_resnet18"]
    x, y = torch.as_tensor(g["x"]), torch.as_tensor(g["y"])
    net.train()
    opt.zero_grad()
    logits = net(x)
    loss = P.soft_tree_sup_loss(tree, nn.CrossEntropyLoss(), logits, y)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), g["train_logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(loss.item(), g["train_loss"], rtol=1e-5)
    norms = np.array([float(p.grad.double().norm()) for p in net.parameters()])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-3)
    opt.step()
    net.eval()
    with torch.no_grad():
        np.testing.assert_allclose(net(x).numpy(), g["eval_logits_after_step"], rtol=1e-3, atol=1e-4)


def test_eager_resnet50_pinned_to_reference_golden(oracle_trees):
    """oracle/torch_port.EagerResNet50 (Bottleneck restatement) == the reference's ResNet50 on the seeded golden train step."""
    import os
    import torch
    import torch.nn as nn
    from conftest import GOLDEN
    from oracle import torch_port as P

    g = np.load(os.path.join(GOLDEN, "backbone_resnet50.npz"))
    torch.manual_seed(0)
    net = P.EagerResNet50(num_classes=10)
    assert [k for k, _ in net.named_parameters()] == list(g["param_names"])
    tree = oracle_trees["cifar10_induced_resnet18"]
    x, y = torch.as_tensor(g["x"]), torch.as_tensor(g["y"])
    net.train()
    logits = net(x)
    loss = P.soft_tree_sup_loss(tree, nn.CrossEntropyLoss(), logits, y)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), g["train_logits"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(loss.item(), g["train_loss"], rtol=1e-5)
    norms = np.array([float(p.grad.double().norm()) for p in net.parameters()])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=5e-3)
