#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the unmodified reference (/root/reference).

Test infrastructure.  Run in the build container only (the reference tree does
not exist on the GPU box):

    python tests/golden/make_golden.py

It re-executes itself with PYTHONHASHSEED=0 (reference leaf-set ordering, see
oracle/ref_shims.py), imports the reference through the shims, and writes

    tests/golden/tree_<case>.npz      numeric vectors of the tree head / losses
    tests/golden/structure.json       Tree/Node tables as the reference builds them
    tests/golden/backbone_resnet18.npz  seeded ResNet18 forward / train-step vectors

Every array is produced by reference code paths:
  soft probs        nbdt/model.py:268-273   (SoftEmbeddedDecisionRules.forward)
  hard preds/paths  nbdt/model.py:145-199   (HardEmbeddedDecisionRules.forward_with_decisions)
  soft decisions    nbdt/model.py:244-266
  node outputs      nbdt/model.py:101-123
  soft/hard loss    nbdt/loss.py:191-266    (TreeSupLoss.forward + forward_tree)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

if os.environ.get("PYTHONHASHSEED") != "0":
    env = dict(os.environ, PYTHONHASHSEED="0")
    os.execve(sys.executable, [sys.executable] + sys.argv, env)

sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from nbdt.model import HardEmbeddedDecisionRules, SoftEmbeddedDecisionRules  # noqa: E402
from nbdt.loss import HardTreeSupLoss, SoftTreeSupLoss  # noqa: E402
from nbdt.tree import Tree  # noqa: E402
from nbdt.models import ResNet18  # noqa: E402

# (case name, dataset, hierarchy, rows, seed)
CASES = [
    ("cifar10_induced", "CIFAR10", "induced", 32, 0),
    ("cifar10_induced_resnet18", "CIFAR10", "induced-ResNet18", 8, 1),
    ("cifar10_wordnet", "CIFAR10", "wordnet", 16, 2),
    ("cifar100_induced", "CIFAR100", "induced", 16, 3),
    ("cifar100_wordnet", "CIFAR100", "wordnet", 16, 4),
    ("tiny200_induced_resnet18", "TinyImagenet200", "induced-ResNet18", 16, 5),
    ("tiny200_wordnet", "TinyImagenet200", "wordnet", 8, 6),
    ("imagenet1000_induced", "Imagenet1000", "induced", 8, 7),
    ("ade20k_induced", "ADE20K", "induced-HRNet-w48", 8, 8),
    ("lip_induced", "LookIntoPerson", "induced-HRNet-w48-cls20", 8, 9),
]

LOSS_VARIANTS = [  # (tag, kwargs)
    ("default", {}),
    ("tsw10", {"tree_supervision_weight": 10.0}),
    ("xw05_tsw2", {"xent_weight": 0.5, "tree_supervision_weight": 2.0}),
]


def tree_structure(tree):
    """Reference Tree/Node tables in a hash-seed independent form."""
    out = {
        "wnids_leaves": list(tree.wnids_leaves),
        "root": tree.root.wnid,
        "inodes": [n.wnid for n in tree.inodes],
        "nodes": [n.wnid for n in tree.nodes],
        "children": {n.wnid: [c.wnid for c in n.children] for n in tree.inodes},
        # leaf class indices under each child, sorted (the reference order is set order)
        "child_index_to_class_index": {
            n.wnid: [sorted(n.child_index_to_class_index[k]) for k in range(n.num_classes)]
            for n in tree.inodes
        },
    }
    steps = tree.get_leaf_to_steps()
    out["leaf_to_steps"] = {
        leaf: [[s["node"].wnid, int(s["next_index"])] for s in steps[leaf]] for leaf in tree.wnids_leaves
    }
    return out


def run_loss(cls, tree, logits, targets, kwargs, dtype):
    crit = cls(dataset=None, criterion=nn.CrossEntropyLoss(), tree=tree, **kwargs)
    x = logits.to(dtype).clone().requires_grad_(True)
    loss = crit(x, targets)
    loss.backward()
    return float(loss.detach().double()), x.grad.detach().double().numpy()


def tree_case(name, dataset, hierarchy, rows, seed, structure):
    tree = Tree(dataset, hierarchy=hierarchy)
    C = len(tree.wnids_leaves)
    g = torch.Generator().manual_seed(seed)
    logits = (3.0 * torch.randn(rows, C, generator=g)).float()
    targets = torch.randint(C, (rows,), generator=g)
    structure[name] = dict(tree_structure(tree), dataset=dataset, hierarchy=hierarchy)

    inode_index = {n.wnid: i for i, n in enumerate(tree.inodes)}
    out = {"logits": logits.numpy(), "targets": targets.numpy()}

    soft = SoftEmbeddedDecisionRules(tree=tree)
    hard = HardEmbeddedDecisionRules(tree=tree)
    with torch.no_grad():
        out["soft_probs"] = soft(logits).numpy()
        out["soft_probs_f64in"] = soft(logits.double()).double().numpy()

        # per-node outputs in tree.inodes order, children padded to Kmax
        w2o = soft.forward_nodes(logits.double())
        kmax = max(n.num_classes for n in tree.inodes)
        N = len(tree.inodes)
        nl = np.full((rows, N, kmax), np.nan)
        pr = np.full((rows, N, kmax), np.nan)
        ent = np.zeros((rows, N))
        prd = np.zeros((rows, N), dtype=np.int64)
        for i, n in enumerate(tree.inodes):
            o = w2o[n.wnid]
            k = n.num_classes
            nl[:, i, :k] = o["logits"].numpy()
            pr[:, i, :k] = o["probs"].numpy()
            ent[:, i] = o["entropy"].numpy()
            prd[:, i] = o["preds"].numpy()
        out["node_logits_f64"] = nl
        out["node_probs_f64"] = pr
        out["node_entropy_f64"] = ent
        out["node_preds"] = prd
        # decision margin per node: gap between best and second-best child logit
        srt = np.sort(np.where(np.isnan(nl), -np.inf, nl), axis=2)
        out["node_margin_f64"] = srt[:, :, -1] - srt[:, :, -2]

        one_hot, decisions = hard.forward_with_decisions(logits)
        out["hard_onehot"] = one_hot.numpy()
        out["hard_pred"] = one_hot.argmax(1).numpy()
        dmax = max(len(d) - 1 for d in decisions)
        pc = np.full((rows, dmax), -1, dtype=np.int64)
        pn = np.full((rows, dmax), -1, dtype=np.int64)  # inode index the decision was taken AT
        pp = np.zeros((rows, dmax))
        pe = np.zeros((rows, dmax))
        for b, d in enumerate(decisions):
            assert d[0]["name"] == "root" and d[0]["prob"] == 1 and d[0]["entropy"] == 0
            at = tree.root
            for j, step in enumerate(d[1:]):
                pc[b, j] = step["next_index"]
                pn[b, j] = inode_index[at.wnid]
                pp[b, j] = step["prob"]
                pe[b, j] = step["entropy"]
                at = step["node"]
            assert at.is_leaf()
        out["hard_path_child"] = pc
        out["hard_path_inode"] = pn
        out["hard_path_prob"] = pp
        out["hard_path_entropy"] = pe
        # the margin along each taken path (robustness of the index comparison)
        m = np.full((rows,), np.inf)
        for b in range(rows):
            for j in range(dmax):
                if pn[b, j] >= 0:
                    m[b] = min(m[b], out["node_margin_f64"][b, pn[b, j]])
        out["hard_path_min_margin"] = m

        # soft decisions (reference uses sample 0's node probs for every sample: model.py:259)
        _, sdec = soft.forward_with_decisions(logits)
        smax = max(len(d) for d in sdec)
        sp = np.zeros((rows, smax))
        se = np.zeros((rows, smax))
        sn = np.full((rows, smax), -2, dtype=np.int64)
        for b, d in enumerate(sdec):
            for j, step in enumerate(d):
                sp[b, j] = step["prob"]
                se[b, j] = step["entropy"]
                sn[b, j] = step["next_index"]
        # NOTE: the reference mutates shared per-leaf step dicts (model.py:253,262-265), so
        # rows predicting the same leaf alias one list; values are equal anyway because they
        # all come from sample 0.
        out["soft_dec_prob"] = sp
        out["soft_dec_entropy"] = se
        out["soft_dec_next_index"] = sn

    for tag, kw in LOSS_VARIANTS:
        for lname, cls in (("soft", SoftTreeSupLoss), ("hard", HardTreeSupLoss)):
            l32, g32 = run_loss(cls, tree, logits, targets, kw, torch.float32)
            l64, g64 = run_loss(cls, tree, logits, targets, kw, torch.float64)
            out[f"{lname}_loss_{tag}_f32"] = np.float64(l32)
            out[f"{lname}_grad_{tag}_f32"] = g32.astype(np.float32)
            out[f"{lname}_loss_{tag}_f64"] = np.float64(l64)
            out[f"{lname}_grad_{tag}_f64"] = g64
    # weight schedule (loss.py:187-209): epoch 3 of 10, start != end
    crit = SoftTreeSupLoss(
        dataset=None, criterion=nn.CrossEntropyLoss(), tree=tree,
        tree_supervision_weight=1.0, tree_supervision_weight_end=5.0, tree_supervision_weight_power=2,
        xent_weight=1.0, xent_weight_end=0.25, xent_weight_power=1,
    )
    crit.set_epoch(3, 10)
    x = logits.double().clone().requires_grad_(True)
    loss = crit(x, targets)
    loss.backward()
    out["soft_loss_sched_f64"] = np.float64(loss.item())
    out["soft_grad_sched_f64"] = x.grad.numpy()

    # bf16 inputs (reference computes in bf16 then; stored for orientation, tolerance 1e-2)
    # (the reference itself raises for bf16 on K>2 nodes: Categorical's simplex check fails in bf16)
    try:
        with torch.no_grad():
            out["soft_probs_bf16in"] = soft(logits.bfloat16()).float().numpy()
    except ValueError as e:
        print(f"  {name}: reference cannot run bf16 logits ({str(e)[:60]}...)")

    np.savez_compressed(os.path.join(HERE, f"tree_{name}.npz"), **out)
    print(f"{name}: C={C} N={len(tree.inodes)} rows={rows} min path margin={out['hard_path_min_margin'].min():.3e}")


def backbone_case():
    """Seeded reference ResNet18 (nbdt/models/resnet.py:115-179): forward + one train step."""
    torch.manual_seed(0)
    net = ResNet18(num_classes=10)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 3, 32, 32, generator=g)
    y = torch.randint(10, (8,), generator=g)
    tree = Tree("CIFAR10", hierarchy="induced-ResNet18")
    crit = SoftTreeSupLoss(dataset=None, criterion=nn.CrossEntropyLoss(), tree=tree)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    out = {"x": x.numpy(), "y": y.numpy()}
    net.train()
    opt.zero_grad()
    logits = net(x)
    loss = crit(logits, y)
    loss.backward()
    out["train_logits"] = logits.detach().numpy()
    out["train_loss"] = np.float64(loss.item())
    names, gnorm, wnorm = [], [], []
    for k, p in net.named_parameters():
        names.append(k)
        gnorm.append(float(p.grad.double().norm()))
        wnorm.append(float(p.detach().double().norm()))
    out["param_names"] = np.array(names)
    out["grad_norms"] = np.array(gnorm)
    out["weight_norms_before"] = np.array(wnorm)
    out["grad_linear_weight"] = net.linear.weight.grad.numpy().copy()
    out["grad_conv1_weight"] = net.conv1.weight.grad.numpy().copy()
    opt.step()
    out["weight_norms_after"] = np.array([float(p.detach().double().norm()) for _, p in net.named_parameters()])
    out["bn1_running_mean_after"] = net.bn1.running_mean.numpy().copy()
    out["bn1_running_var_after"] = net.bn1.running_var.numpy().copy()
    net.eval()
    with torch.no_grad():
        out["eval_logits_after_step"] = net(x).numpy()
    np.savez_compressed(os.path.join(HERE, "backbone_resnet18.npz"), **out)
    print("backbone_resnet18: loss", out["train_loss"])


def main():
    torch.set_num_threads(8)
    structure = {}
    for case in CASES:
        tree_case(*case, structure)
    with open(os.path.join(HERE, "structure.json"), "w") as f:
        json.dump(structure, f)
    backbone_case()


if __name__ == "__main__":
    main()
