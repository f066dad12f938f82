"""More golden vectors from the EXECUTING reference (see make_golden.py for the shims):

    tests/golden/backbone_resnet50.npz   seeded reference ResNet50 (Bottleneck, nbdt/models/resnet.py:77-112) train step
    tests/golden/seg_cityscapes.npz      SoftSegNBDT / HardSegNBDT outputs (nbdt/model.py:376-399) and SoftSegTreeSupLoss
                                          (nbdt/loss.py:318-327) loss + gradient on (N, C, H, W) logits, Cityscapes hierarchy
    tests/golden/filtered_cifar100.npz   EmbeddedDecisionRules.get_node_logits_filtered (nbdt/model.py:127-143) for every inner node
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if os.environ.get("PYTHONHASHSEED") != "0":
    os.execve(sys.executable, [sys.executable] + sys.argv, dict(os.environ, PYTHONHASHSEED="0"))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from nbdt.loss import SoftSegTreeSupLoss, SoftTreeSupLoss  # noqa: E402
from nbdt.model import HardEmbeddedDecisionRules, HardSegNBDT, SoftSegNBDT  # noqa: E402
from nbdt.models import ResNet50  # noqa: E402
from nbdt.tree import Tree  # noqa: E402

os.chdir(ref_shims.REFERENCE_ROOT)


def resnet50():
    torch.manual_seed(0)
    net = ResNet50(num_classes=10)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 3, 32, 32, generator=g)
    y = torch.randint(10, (8,), generator=g)
    crit = SoftTreeSupLoss(dataset=None, criterion=nn.CrossEntropyLoss(), tree=Tree("CIFAR10", hierarchy="induced-ResNet18"))
    net.train()
    logits = net(x)
    loss = crit(logits, y)
    loss.backward()
    names = [k for k, _ in net.named_parameters()]
    out = {"x": x.numpy(), "y": y.numpy(), "train_logits": logits.detach().numpy(), "train_loss": np.float64(loss.item()),
           "param_names": np.array(names), "grad_norms": np.array([float(p.grad.double().norm()) for p in net.parameters()]),
           "grad_linear_weight": net.linear.weight.grad.numpy().copy(), "grad_conv1_weight": net.conv1.weight.grad.numpy().copy()}
    net.eval()
    with torch.no_grad():
        out["eval_logits"] = net(x).numpy()
    np.savez_compressed(os.path.join(HERE, "backbone_resnet50.npz"), **out)
    print("resnet50 loss", out["train_loss"], len(names), "parameters")


def seg():
    pg = "nbdt/hierarchies/Cityscapes/graph-induced-HRNet-w18-v1.json"
    pw = os.path.join(HERE, "wnids_cityscapes.txt")  # Cityscapes ships no wnids file: the sorted leaves, written once
    import json

    data = json.load(open(pg))
    succ = {n["id"]: [] for n in data["nodes"]}
    for e in data["links"]:
        succ[e["source"]].append(e["target"])
    leaves = sorted(w for w in succ if not succ[w])
    with open(pw, "w") as f:
        f.write("\n".join(leaves))
    classes = [f"c{i}" for i in range(len(leaves))]
    tree = Tree("Cityscapes", path_graph=pg, path_wnids=pw, classes=classes)
    g = torch.Generator().manual_seed(0)
    x = (2 * torch.randn(2, len(leaves), 5, 7, generator=g))
    y = torch.randint(len(leaves), (2, 5, 7), generator=g)
    ident = nn.Identity()
    out = {"logits": x.numpy(), "targets": y.numpy()}
    kw = dict(path_graph=pg, path_wnids=pw, classes=classes)
    out["soft_seg"] = SoftSegNBDT("Cityscapes", ident, **kw)(x).detach().numpy()
    out["hard_seg"] = HardSegNBDT("Cityscapes", ident, **kw)(x).detach().numpy()
    for tsw in (1.0, 0.5):
        xx = x.clone().requires_grad_(True)
        crit = SoftSegTreeSupLoss(dataset=None, criterion=nn.CrossEntropyLoss(), tree=tree, tree_supervision_weight=tsw)
        loss = crit(xx, y)
        loss.backward()
        out[f"seg_loss_tsw{tsw}"] = np.float64(loss.item())
        out[f"seg_grad_tsw{tsw}"] = xx.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "seg_cityscapes.npz"), **out)
    print("seg", out["soft_seg"].shape, out["seg_loss_tsw1.0"])


def filtered():
    tree = Tree("CIFAR100", hierarchy="induced")
    rules = HardEmbeddedDecisionRules("CIFAR100", tree=tree)
    g = torch.Generator().manual_seed(3)
    x = 3 * torch.randn(40, 100, generator=g)
    y = torch.randint(100, (40,), generator=g)
    out = {"logits": x.numpy(), "targets": y.numpy(), "inodes": np.array([n.wnid for n in tree.inodes])}
    for i, node in enumerate(tree.inodes):
        sel, sub, tgt = HardEmbeddedDecisionRules.get_node_logits_filtered(node, x, y)
        out[f"sel_{i}"] = np.array(sel)
        out[f"sub_{i}"] = sub.numpy()
        out[f"tgt_{i}"] = np.array(tgt)
    np.savez_compressed(os.path.join(HERE, "filtered_cifar100.npz"), **out)
    print("filtered", len(tree.inodes), "nodes")


if __name__ == "__main__":
    resnet50()
    seg()
    filtered()
