"""GPU checks of the kernels either side of the hot path (csrc/aux_ops.cu) and of the reference's main.py running on this package."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from test_callers import PKG, ROOT, golden, weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("case", ["cifar10", "cifar100", "tiny200", "imagenet1000"])
def test_ward_kernel_matches_sklearn_children(case):
    """nbdt_ward_linkage == sklearn AgglomerativeClustering(linkage="ward").children_ (what the reference clusters with), merge for
    merge, on seeded clustered classifier matrices up to 1000 x 1280; heights to fp64 round-off."""
    from nbdt.graph import ward_children

    g = golden()[case]
    W = weights(g["C"], g["D"], g["seed"]).to(DEV)
    children, heights = ward_children(W)
    assert children == [tuple(c) for c in g["children"]]
    np.testing.assert_allclose(heights, g["heights"], rtol=1e-9)


def test_induced_graph_on_device_equals_the_reference_graph(monkeypatch):
    from networkx.readwrite.json_graph import node_link_data

    from nbdt.graph import build_induced_graph
    from nbdt.thirdparty.wn import get_wnids_from_dataset

    monkeypatch.chdir(PKG)
    g = golden()["tiny200"]
    W = weights(g["C"], g["D"], g["seed"]).to(DEV)
    G = build_induced_graph(get_wnids_from_dataset(g["dataset"]), checkpoint=None, state_dict={"linear.weight": W}, dataset=g["dataset"])
    got, want = node_link_data(G, edges="links"), g["graph"]
    assert [n["id"] for n in got["nodes"]] == [n["id"] for n in want["nodes"]]
    assert [(e["source"], e["target"]) for e in got["links"]] == [(e["source"], e["target"]) for e in want["links"]]


def test_soft_tree_loss_reinduces_and_lands_on_a_specialised_kernel(monkeypatch, tmp_path):
    """SoftTreeLoss.set_epoch (reference loss.py:303-315): rebuild the hierarchy from the live model, keep training -- the new tree
    gets its own run-time compiled kernel."""
    import torch.nn as nn

    from nbdt._b200 import lib as L
    from nbdt.loss import SoftTreeLoss
    from nbdt.models import ResNet18

    monkeypatch.chdir(PKG)
    torch.manual_seed(0)
    net = ResNet18(num_classes=10).to(DEV)
    crit = SoftTreeLoss(dataset="CIFAR10", criterion=nn.CrossEntropyLoss(), hierarchy="induced-ResNet18", arch="ResNet18", net=net,
                        checkpoint_path=str(tmp_path / "ckpt.pth"), tree_start_epochs=1, tree_update_every_epochs=1, tree_update_end_epochs=3)
    x = torch.randn(16, 3, 32, 32, device=DEV)
    y = torch.randint(10, (16,), device=DEV)
    crit.set_epoch(0, 4)                 # before tree_start_epochs: plain cross-entropy, shipped hierarchy untouched
    l0 = crit(net(x), y)
    shipped = crit.tree.path_graph
    crit.set_epoch(1, 4)                 # re-induction from the live (random) classifier -> <checkpoint dir>/graph-epoch1.json
    l1 = crit(net(x), y)
    l1.backward()
    assert torch.isfinite(l0) and torch.isfinite(l1)
    assert crit.tree.path_graph != shipped and os.path.exists(tmp_path / "ckpt" / "graph-epoch1.json")
    assert len(crit.tree.inodes) == 9 and all(p.grad is not None for p in net.parameters())
    name = L.load().nbdt_tree_kernel_variant(crit.tree.device_tree(torch.device(DEV)).handle).decode()
    assert name.startswith("ka"), name


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(513, 10), (77, 1000), (1, 3)])
def test_topk_counters_match_torch_topk(dtype, shape):
    from nbdt import metrics

    g = torch.Generator(device=DEV).manual_seed(shape[0])
    x = torch.randn(shape, device=DEV, generator=g).to(dtype)
    x[:, 0] = x[:, -1]  # exact ties
    y = torch.randint(shape[1], (shape[0],), device=DEV, generator=g)
    for k in (1, 2, 5):
        if k > shape[1]:
            continue
        m = metrics.TopK(k)
        h = shape[0] // 2
        m.forward(x[:h], y[:h])
        m.forward(x[h:], y[h:])          # counters accumulate on the device across batches
        xf = x.float()
        xt = xf.gather(1, y[:, None])
        idx = torch.arange(shape[1], device=DEV)[None]
        want = int((((xf > xt) | ((xf == xt) & (idx < y[:, None]))).sum(1) < k).sum())
        assert (m.correct, m.total) == (want, shape[0])
        # and against torch.topk itself where there are no ties
        xn = torch.randn(shape, device=DEV, generator=g).to(dtype)
        m2 = metrics.TopK(k)
        m2.forward(xn, y)
        assert abs(m2.correct - int((xn.float().topk(k).indices == y[:, None]).any(1).sum())) <= 1


def test_decision_rule_analyzers_count_like_the_reference(structure):
    """analysis.Hard/SoftEmbeddedDecisionRules: accuracy of rules(outputs) -- equals counting the oracle's predictions by hand."""
    from conftest import case_paths, load_golden
    from nbdt import analysis
    from nbdt.tree import Tree
    from oracle import tree_oracle as O

    s = structure["cifar100_induced"]
    g = load_golden("cifar100_induced")
    ot = O.OracleTree(*case_paths(structure, "cifar100_induced"))
    x = torch.as_tensor(g["logits"], device=DEV)
    y = torch.as_tensor(g["targets"], device=DEV)
    for cls, pred in ((analysis.HardEmbeddedDecisionRules, O.hard_decisions(ot, g["logits"].astype(np.float64))[0]),
                      (analysis.SoftEmbeddedDecisionRules, O.soft_probs(ot, g["logits"].astype(np.float64)).argmax(1))):
        a = cls(dataset=s["dataset"], hierarchy=s["hierarchy"], tree=Tree(s["dataset"], hierarchy=s["hierarchy"]))
        a.start_epoch(0)
        a.start_test(0)
        a.update_batch(x, y, None)
        a.end_test(0)
        assert a.metric.total == len(y) and a.metric.correct == int((np.asarray(pred) == g["targets"]).sum())


def test_augment_kernel_matches_torchvision_transforms():
    """Same pixels as RandomCrop(32, padding=4) + RandomHorizontalFlip + ToTensor + Normalize for given (dx, dy, flip)."""
    import torchvision.transforms.functional as F

    from nbdt.data import cifar
    from nbdt.data.device import DeviceLoader

    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(37, 32, 32, 3), dtype=np.uint8)
    labels = rng.integers(0, 10, size=37)
    loader = DeviceLoader(imgs, labels, 16, cifar.MEAN, cifar.STD, pad=4, train=True, device=DEV, seed=1)
    params = torch.stack([torch.randint(0, 9, (37,)), torch.randint(0, 9, (37,)), torch.randint(0, 2, (37,))], 1).int()
    got = loader.augment(torch.arange(37, device=DEV), params).cpu()
    for n in range(37):
        dx, dy, flip = params[n].tolist()
        t = torch.from_numpy(imgs[n]).permute(2, 0, 1)
        t = F.crop(F.pad(t, 4), dy, dx, 32, 32)
        if flip:
            t = F.hflip(t)
        want = F.normalize(t.float() / 255, cifar.MEAN, cifar.STD)
        torch.testing.assert_close(got[n], want, rtol=1e-5, atol=1e-5)
    # the loader: every sample exactly once per epoch, labels follow their images
    seen = []
    for xb, yb in loader:
        assert xb.shape[1:] == (3, 32, 32) and xb.dtype == torch.float32 and xb.is_cuda
        seen += yb.tolist()
    assert sorted(seen) == sorted(labels.tolist())
    ev = DeviceLoader(imgs, labels, 16, cifar.MEAN, cifar.STD, train=False, device=DEV)
    xb, yb = next(iter(ev))
    torch.testing.assert_close(xb.cpu()[0], F.normalize(torch.from_numpy(imgs[0]).permute(2, 0, 1).float() / 255, cifar.MEAN, cifar.STD), rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "main.py")), reason="reference not vendored (oracle/make_ref.py)")
@pytest.mark.parametrize("loss,analysis", [("SoftTreeSupLoss", "SoftEmbeddedDecisionRules"), ("HardTreeSupLoss", "HardEmbeddedDecisionRules")])
def test_reference_main_py_trains_on_this_package(loss, analysis, tmp_path):
    """The reference's main.py, unchanged, end to end on the GPU: synthetic CIFAR10, ResNet18 on the tcgen05 kernels, the fused
    tree-supervision loss, the decision-rule analyzer, a checkpoint written."""
    shutil.copy(os.path.join(ROOT, "oracle", "_ref", "main.py"), tmp_path / "main.py")
    env = dict(os.environ, NBDT_SYNTHETIC_DATA="1", NBDT_SYNTHETIC_SAMPLES="128", PYTHONPATH=PKG)
    res = subprocess.run([sys.executable, "main.py", "--dataset", "CIFAR10", "--arch", "ResNet18", "--loss", loss, "--analysis", analysis,
                          "--epochs", "1", "--batch-size", "64"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-3000:]
    assert "Best accuracy" in out and "NBDT-" in out, out[-1500:]
    assert any(f.endswith(".pth") for f in os.listdir(tmp_path / "checkpoint"))
