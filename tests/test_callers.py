"""CPU checks of the modules either side of the hot path: hierarchy induction against graphs the
EXECUTING reference built (tests/golden/induced_graphs.json, tests/golden/make_golden_induced.py), the analyzers' bookkeeping,
the datasets' interface, and that the reference's own main.py drives this package unchanged up to the first CUDA-only call."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-backed-decision-trees_b200")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def golden():
    with open(os.path.join(ROOT, "tests", "golden", "induced_graphs.json")) as f:
        return json.load(f)


def weights(C, D, seed):  # same construction as tests/golden/make_golden_induced.py
    g = torch.Generator().manual_seed(seed)
    groups = torch.randn(max(2, C // 7), D, generator=g)
    return (groups[torch.randint(len(groups), (C,), generator=g)] + 0.5 * torch.randn(C, D, generator=g)).float()


@pytest.mark.parametrize("case", ["cifar10", "cifar100"])
def test_induced_graph_equals_the_reference_graph(case, monkeypatch):
    """Same nodes (order, generated ids, labels) and same edges (order) as nbdt.graph.build_induced_graph of the reference."""
    from networkx.readwrite.json_graph import node_link_data

    from nbdt.graph import build_induced_graph
    from nbdt.thirdparty.wn import get_wnids_from_dataset

    monkeypatch.chdir(PKG)
    g = golden()[case]
    W = weights(g["C"], g["D"], g["seed"])
    G = build_induced_graph(get_wnids_from_dataset(g["dataset"]), checkpoint=None, state_dict={"fc.weight": W}, dataset=g["dataset"])
    got, want = node_link_data(G, edges="links"), g["graph"]
    assert [n["id"] for n in got["nodes"]] == [n["id"] for n in want["nodes"]]
    assert [(e["source"], e["target"]) for e in got["links"]] == [(e["source"], e["target"]) for e in want["links"]]
    assert [n.get("label") for n in got["nodes"]] == [n.get("label") for n in want["nodes"]]


def test_generate_hierarchy_writes_a_loadable_tree(tmp_path, monkeypatch):
    """Tree.update_from_model's path: live model -> induced graph file -> Tree (what SoftTreeLoss.set_epoch does, loss.py:303-315)."""
    from nbdt.hierarchy import generate_hierarchy, test_hierarchy
    from nbdt.tree import Tree

    monkeypatch.chdir(PKG)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = torch.nn.Linear(32, 10)

    torch.manual_seed(0)
    path = generate_hierarchy(dataset="CIFAR10", method="induced", arch="ResNet18", model=Net(), path=str(tmp_path / "graph-induced-test.json"))
    test_hierarchy(path=path, dataset="CIFAR10")
    tree = Tree("CIFAR10", path_graph=path)
    assert len(tree.inodes) == 9 and all(len(n.children) == 2 for n in tree.inodes)
    tree2 = Tree("CIFAR10")
    tree2.update_from_model(Net(), arch="ResNet18", dataset="CIFAR10", path_graph=str(tmp_path / "graph-induced-test2.json"))
    assert len(tree2.inodes) == 9


def test_graph_file_names_follow_the_shipped_convention():
    from nbdt.graph import generate_graph_fname

    assert generate_graph_fname("induced", arch="ResNet18") == "graph-induced-ResNet18"
    assert generate_graph_fname("induced", checkpoint="checkpoint/ckpt-CIFAR10-wrn28_10_cifar10.pth") == "graph-induced-wrn28_10_cifar10"
    assert generate_graph_fname("random", seed=3, branching_factor=4, no_prune=True) == "graph-random-seed3-branch4-noprune"
    assert generate_graph_fname("wordnet", multi_path=True) == "graph-wordnet-multi"
    assert generate_graph_fname("induced", path="/x/y/graph-custom.json") == "graph-custom"


def test_random_graph_is_a_tree_over_the_classes(monkeypatch):
    from nbdt.graph import build_random_graph, prune_single_successor_nodes
    from nbdt.thirdparty.nx import get_leaves, get_roots
    from nbdt.thirdparty.wn import get_wnids_from_dataset

    monkeypatch.chdir(PKG)
    wnids = get_wnids_from_dataset("CIFAR100")
    G = prune_single_successor_nodes(build_random_graph(wnids, seed=1, branching_factor=3))
    assert len(list(get_roots(G))) == 1 and set(get_leaves(G)) == set(wnids)
    assert all(len(G.pred[n]) <= 1 for n in G.nodes) and max(len(G.succ[n]) for n in G.nodes) <= 3


def test_topk_metric_host_path_matches_the_reference_rule():
    """CPU tensors (the reference's own tests run there): same counts as `target in topk(outputs, k)`."""
    from nbdt import metrics

    g = torch.Generator().manual_seed(0)
    x = torch.randn(257, 37, generator=g)
    y = torch.randint(37, (257,), generator=g)
    for k in (1, 2, 5, 10):
        m = metrics.TopK(k)
        m.forward(x[:100], y[:100])
        m.forward(x[100:], y[100:])
        want = sum(int(t in p) for p, t in zip(x.topk(k).indices.tolist(), y.tolist()))
        assert (m.correct, m.total) == (want, 257)
        assert abs(m.report() - want / 257) < 1e-12


def test_analyzers_bracket_the_loops_like_the_reference():
    from nbdt import analysis

    a = analysis.ConfusionMatrix(classes=("a", "b", "c"))
    calls = []

    @a.test_function
    def test(epoch):
        calls.append(epoch)
        a.update_batch(torch.tensor([[2.0, 1, 0], [0, 3, 1], [0, 3, 1], [0, 0, 9]]), torch.tensor([0, 1, 2, 2]), None)

    with a.epoch_context(4):
        test(4)
    assert calls == [4]
    np.testing.assert_array_equal(a.m + 0, [[1, 0, 0], [0, 1, 0], [0, 1, 1]])
    np.testing.assert_allclose(a.recall().diagonal(), [1, 1, 0.5])
    assert set(analysis.names) >= {"Noop", "ConfusionMatrix", "HardEmbeddedDecisionRules", "SoftEmbeddedDecisionRules"}
    with pytest.raises(NotImplementedError):
        a.start_epoch(5), a.start_train(5)


def test_datasets_serve_the_reference_interface(monkeypatch, tmp_path):
    from nbdt import data

    monkeypatch.setenv("NBDT_SYNTHETIC_DATA", "1")
    monkeypatch.setenv("NBDT_SYNTHETIC_SAMPLES", "40")
    assert data.cifar.names == ("CIFAR10", "CIFAR100") and data.imagenet.names == ("TinyImagenet200", "Imagenet1000") and len(data.custom.names) == 12
    ds = data.CIFAR10(root=str(tmp_path), train=True, download=True, transform=data.CIFAR10.transform_train())
    x, y = ds[3]
    assert x.shape == (3, 32, 32) and 0 <= y < 10 and len(ds.classes) == 10 and type(ds).__name__ == "CIFAR10"
    tiny = data.TinyImagenet200(root=str(tmp_path), train=False, transform=data.TinyImagenet200.transform_val())
    assert tiny[0][0].shape == (3, 64, 64) and len(tiny.classes) == 200
    inc = data.CIFAR10IncludeLabels(root=str(tmp_path), train=True, include_labels=(1, 3))
    assert {inc[i][1] for i in range(len(inc))} <= {1, 3} and type(inc).__name__ == "CIFAR10IncludeLabels"
    inv = ds.transform_val_inverse()
    back = inv(data.CIFAR10.transform_val()(__import__("PIL.Image", fromlist=["x"]).fromarray(ds.data[0]))[None])
    np.testing.assert_allclose(back[0].permute(1, 2, 0).numpy() * 255, ds.data[0], atol=1e-3)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "main.py")), reason="reference not vendored (oracle/make_ref.py)")
def test_reference_main_py_drives_this_package_unchanged(tmp_path):
    """The reference's main.py (byte-for-byte copy) imports `data, analysis, loss, models, metrics, tree` from THIS package, parses
    its arguments, builds datasets / model / tree / SoftTreeSupLoss / analyzer and enters the train loop; on a box without a GPU
    the first forward stops at the CUDA-only guard (no CPU fallback by design).  The GPU test runs the loop itself."""
    shutil.copy(os.path.join(ROOT, "oracle", "_ref", "main.py"), tmp_path / "main.py")
    env = dict(os.environ, NBDT_SYNTHETIC_DATA="1", NBDT_SYNTHETIC_SAMPLES="64", PYTHONPATH=PKG)
    res = subprocess.run([sys.executable, "main.py", "--dataset", "CIFAR10", "--arch", "ResNet18", "--loss", "SoftTreeSupLoss", "--analysis",
                          "SoftEmbeddedDecisionRules", "--epochs", "1", "--batch-size", "32"], cwd=tmp_path, env=env, capture_output=True, text=True,
                         timeout=600)
    out = res.stdout + res.stderr
    assert "==> Building model.." in out and "Epoch: 0" in out, out[-1500:]
    if not torch.cuda.is_available():
        assert "expected a CUDA tensor" in out, out[-1500:]
    else:
        assert res.returncode == 0 and "Best accuracy" in out, out[-1500:]
