"""The reference's own smoke tests (tests/test_inference.py:6-45, tests/test_train.py:9-49) ported to
the CUDA device (this package has no CPU path), plus checks of the drop-in API contract the
reference tests never assert: output flags, decision lists, autograd through the fused loss."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture
def criterion():
    return nn.CrossEntropyLoss()


def resnet18(num_classes):
    from nbdt.models import ResNet18

    return ResNet18(num_classes=num_classes).to(DEV)


@pytest.mark.parametrize("dataset,classes,size", [("CIFAR10", 10, 32), ("CIFAR100", 100, 32), ("TinyImagenet200", 200, 64)])
@pytest.mark.parametrize("kind", ["soft", "hard"])
def test_nbdt_inference(dataset, classes, size, kind):
    from nbdt.model import HardNBDT, SoftNBDT

    cls = SoftNBDT if kind == "soft" else HardNBDT
    model = cls(dataset=dataset, model=resnet18(classes), hierarchy="induced")
    out = model(torch.randn(1, 3, size, size, device=DEV))
    assert out.shape == (1, classes) and out.dtype == torch.float32 and out.device.type == "cuda"
    assert getattr(out, "_nbdt_output_flag", False) is True
    assert abs(out.sum().item() - 1.0) < 1e-5  # soft: distribution; hard: one-hot


@pytest.mark.parametrize("dataset,classes", [("CIFAR10", 10), ("CIFAR100", 100), ("TinyImagenet200", 200)])
def test_criterion(dataset, classes, criterion):
    from nbdt.loss import SoftTreeSupLoss

    crit = SoftTreeSupLoss(dataset=dataset, criterion=criterion, hierarchy="induced")
    loss = crit(torch.randn((1, classes), device=DEV), torch.randint(classes, (1,), device=DEV))
    assert loss.dim() == 0 and torch.isfinite(loss)


@pytest.mark.parametrize("kind", ["hard", "soft"])
def test_nbdt_gradient(kind, criterion):
    from nbdt.loss import HardTreeSupLoss, SoftTreeSupLoss

    net = resnet18(10)
    out = net(torch.randn(1, 3, 32, 32, device=DEV))
    assert out.requires_grad
    cls = HardTreeSupLoss if kind == "hard" else SoftTreeSupLoss
    loss = cls(dataset="CIFAR10", criterion=criterion, hierarchy="induced")(out, torch.randint(10, (1,), device=DEV))
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_loss_rejects_nbdt_output(criterion):
    from nbdt.loss import SoftTreeSupLoss
    from nbdt.model import SoftNBDT

    model = SoftNBDT(dataset="CIFAR10", model=resnet18(10), hierarchy="induced")
    y = model(torch.randn(2, 3, 32, 32, device=DEV))
    with pytest.raises(AssertionError):
        SoftTreeSupLoss(dataset="CIFAR10", criterion=criterion, hierarchy="induced")(y, torch.zeros(2, dtype=torch.long, device=DEV))


def test_hard_decisions_match_reference_golden():
    from nbdt.model import HardEmbeddedDecisionRules
    from nbdt.tree import Tree

    g = load_golden("cifar100_wordnet")
    tree = Tree("CIFAR100", hierarchy="wordnet")
    rules = HardEmbeddedDecisionRules(tree=tree)
    x = torch.as_tensor(g["logits"], device=DEV)
    onehot, decisions = rules.forward_with_decisions(x)
    assert np.array_equal(onehot.cpu().numpy(), g["hard_onehot"])
    assert onehot._nbdt_output_flag is True
    for b, steps in enumerate(decisions):
        assert steps[0] == {"node": tree.root, "name": "root", "prob": 1, "entropy": 0}
        assert [s["next_index"] for s in steps[1:]] == [k for k in g["hard_path_child"][b] if k >= 0]
        assert steps[-1]["node"].is_leaf()
        assert tree.wnid_to_class_index[steps[-1]["node"].wnid] == g["hard_pred"][b]
        for j, s in enumerate(steps[1:]):
            assert abs(s["prob"] - g["hard_path_prob"][b, j]) < 1e-5
            assert abs(s["entropy"] - g["hard_path_entropy"][b, j]) < 1e-5
    # compatibility entry points: node-output dict -> traverse_tree agrees with the fused walk
    w2o = rules.forward_nodes(x)
    assert set(w2o) == {n.wnid for n in tree.inodes}
    preds, dec2 = HardEmbeddedDecisionRules.traverse_tree(w2o, tree)
    assert np.array_equal(preds.cpu().numpy(), g["hard_pred"])
    assert [[s.get("next_index") for s in d] for d in dec2] == [[s.get("next_index") for s in d] for d in decisions]


def test_soft_decisions_match_reference_golden():
    from nbdt.model import SoftEmbeddedDecisionRules
    from nbdt.tree import Tree

    for case, dataset, hierarchy in (("cifar10_induced", "CIFAR10", "induced"), ("tiny200_wordnet", "TinyImagenet200", "wordnet")):
        g = load_golden(case)
        rules = SoftEmbeddedDecisionRules(tree=Tree(dataset, hierarchy=hierarchy))
        x = torch.as_tensor(g["logits"], device=DEV)
        probs, decisions = rules.forward_with_decisions(x)
        np.testing.assert_allclose(probs.cpu().numpy(), g["soft_probs"], rtol=3e-5, atol=1e-8)
        for b, steps in enumerate(decisions):
            n = len(steps)
            assert [s["next_index"] for s in steps] == list(g["soft_dec_next_index"][b, :n])
            np.testing.assert_allclose([s["prob"] for s in steps], g["soft_dec_prob"][b, :n], rtol=2e-5, atol=1e-7)
            np.testing.assert_allclose([s["entropy"] for s in steps], g["soft_dec_entropy"][b, :n], rtol=1e-4, atol=5e-6)
        # node-dict route (traverse_tree from forward_nodes) gives the same probabilities
        again = rules(x, rules.forward_nodes(x))
        np.testing.assert_allclose(again.cpu().numpy(), g["soft_probs"], rtol=3e-5, atol=1e-8)


def test_get_node_logits_variants():
    from nbdt.model import HardEmbeddedDecisionRules as Rules
    from nbdt.tree import Tree

    g = load_golden("cifar10_wordnet")
    tree = Tree("CIFAR10", hierarchy="wordnet")
    x = torch.as_tensor(g["logits"], device=DEV)
    for i, node in enumerate(tree.inodes):
        k = node.num_classes
        want = g["node_logits_f64"][:, i, :k]
        np.testing.assert_allclose(Rules.get_node_logits(x, node).cpu().numpy(), want, rtol=1e-5, atol=2e-6)
        alt = Rules.get_node_logits(x, new_to_old_classes=node.child_index_to_class_index, num_classes=k)
        np.testing.assert_allclose(alt.cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    node = tree.inodes[1]
    sel, sub, tgt = Rules.get_node_logits_filtered(node, x, g["targets"].tolist())
    assert sum(sel) == sub.shape[0] == len(tgt)


@pytest.mark.parametrize("dataset,hierarchy", [("CIFAR10", "induced-ResNet18"), ("CIFAR100", "wordnet-single"), ("TinyImagenet200", "induced")])
def test_node_outputs_carry_gradients(dataset, hierarchy):
    """Criteria written on the per-node entry points (`forward_nodes`, `get_node_logits`) are differentiable in the reference
    (autograd ops, nbdt/model.py:83-123); here the forward is the fused kernel and the backward an autograd.Function: the
    gradient of a made-up criterion over node logits, child probabilities and entropies equals that of the same criterion written
    in plain torch ops on a double-precision copy."""
    from nbdt.model import HardEmbeddedDecisionRules
    from nbdt.tree import Tree

    tree = Tree(dataset, hierarchy=hierarchy)
    rules = HardEmbeddedDecisionRules(tree=tree)
    C = len(tree.classes)
    torch.manual_seed(3)
    x = (2 * torch.randn(5, C, device=DEV)).requires_grad_(True)
    out = rules.forward_nodes(x)
    loss = sum((o["logits"].float() ** 2).sum() * 0.1 + (o["probs"] ** 2).sum() + 0.5 * o["entropy"].sum() for o in out.values())
    node = tree.inodes[len(tree.inodes) // 2]
    loss = loss + rules.get_node_logits(x, node).float().sum()
    loss.backward()

    xr = x.detach().double().requires_grad_(True)
    want = 0
    for n in tree.inodes:
        nl = torch.stack([xr[:, list(n.child_index_to_class_index[k])].mean(1) for k in range(n.num_classes)], 1)
        p = torch.softmax(nl, 1)
        want = want + (nl ** 2).sum() * 0.1 + (p ** 2).sum() - 0.5 * (p * torch.log(p)).sum()
        if n is node:
            want = want + nl.sum()
    want.backward()
    assert abs(loss.item() - want.item()) <= 1e-4 * abs(want.item())
    np.testing.assert_allclose(x.grad.double().cpu().numpy(), xr.grad.cpu().numpy(), rtol=2e-4, atol=2e-5 * xr.grad.abs().max().item())
    assert not out[tree.inodes[0].wnid]["preds"].requires_grad


def test_seg_nbdt_shapes():
    from nbdt.model import SoftSegNBDT

    class Tiny(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(3, 20, 1)

        def forward(self, x):
            return self.conv(x)

    model = SoftSegNBDT(dataset="LookIntoPerson", model=Tiny().to(DEV), hierarchy="induced-HRNet-w48-cls20")
    out = model(torch.randn(2, 3, 5, 7, device=DEV))
    assert out.shape == (2, 20, 5, 7)
    assert torch.allclose(out.sum(1), torch.ones(2, 5, 7, device=DEV), atol=1e-5)
