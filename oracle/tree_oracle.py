"""TEST INFRASTRUCTURE ONLY -- CPU (numpy, float64) restatement of the NBDT tree head and
tree-supervision losses.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg as the CHECKER; the product path never imports it.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
vectors produced by executing the unmodified reference (tests/golden/make_golden.py).

Each function follows the reference one node at a time (the reference's own formulation,
not the fused closed form the CUDA kernels use), citing reference file:line.
"""
import json

import numpy as np


class OracleTree:
    """Hierarchy tables as the reference builds them.

    reference: nbdt/thirdparty/nx.py:69-75 (read_graph), nbdt/thirdparty/wn.py:24-31 (get_wnids),
    nbdt/tree.py:160-174 (load_hierarchy: nodes sorted by wnid, inodes = non-leaves),
    nbdt/tree.py:105-125 (build_class_mappings), nbdt/tree.py:202-207 (root).
    """

    def __init__(self, path_graph, path_wnids):
        with open(path_graph) as f:
            data = json.load(f)
        with open(path_wnids) as f:
            self.wnids = [line.strip() for line in f.readlines()]
        ids = [n["id"] for n in data["nodes"]]
        self.succ = {i: [] for i in ids}
        self.pred = {i: [] for i in ids}
        for e in data["links"]:  # child order == edge insertion order (nx DiGraph adjacency)
            self.succ[e["source"]].append(e["target"])
            self.pred[e["target"]].append(e["source"])
        self.class_index = {w: i for i, w in enumerate(self.wnids)}
        self.nodes = sorted(ids)
        self.inodes = [w for w in self.nodes if self.succ[w]]
        roots = [w for w in self.inodes if not self.pred[w]]
        assert len(roots) == 1
        self.root = roots[0]
        # leaves (as class indices) under every child of every inner node
        self.child_classes = {}
        for w in self.inodes:
            self.child_classes[w] = [sorted(self._leaves(c)) for c in self.succ[w]]

    def _leaves(self, w):
        if not self.succ[w]:
            return [self.class_index[w]]
        out = []
        for c in self.succ[w]:
            out.extend(self._leaves(c))
        return out

    @property
    def num_classes(self):
        return len(self.wnids)


def _softmax(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def _log_softmax(z):
    z = z - z.max(axis=1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=1, keepdims=True))


def node_outputs(tree, x):
    """reference: nbdt/model.py:83-120 (get_node_logits + get_all_node_outputs).

    logits[b,k] = mean of the leaf logits under child k; preds = first argmax;
    probs = softmax; entropy = -sum p ln p (torch Categorical.entropy clamps log at eps).
    """
    x = np.asarray(x, dtype=np.float64)
    out = {}
    for w in tree.inodes:
        nl = np.stack([x[:, idx].mean(axis=1) for idx in tree.child_classes[w]], axis=1)
        p = _softmax(nl)
        eps = np.finfo(np.float64).eps
        logp = np.log(np.clip(p, eps, 1 - eps))  # torch: probs_to_logits clamps probs
        out[w] = {
            "logits": nl,
            "preds": nl.argmax(axis=1),  # numpy argmax = first maximum, like torch.max
            "probs": p,
            "entropy": -(p * logp).sum(axis=1),
        }
    return out


def soft_probs(tree, x, nodes=None):
    """reference: nbdt/model.py:207-242 (SoftEmbeddedDecisionRules.traverse_tree).

    class_probs starts at ones and is multiplied, per inner node, by the probability of
    the child each class sits under."""
    x = np.asarray(x, dtype=np.float64)
    nodes = nodes or node_outputs(tree, x)
    cp = np.ones((x.shape[0], tree.num_classes))
    for w in tree.inodes:
        for k, idx in enumerate(tree.child_classes[w]):
            cp[:, idx] *= nodes[w]["probs"][:, k : k + 1]
    return cp


def hard_decisions(tree, x, nodes=None):
    """reference: nbdt/model.py:145-186 (HardEmbeddedDecisionRules.traverse_tree).

    Greedy root->leaf walk.  Returns pred (B,), and per row a list of
    (inode wnid the decision was taken at, child index, prob, entropy)."""
    x = np.asarray(x, dtype=np.float64)
    nodes = nodes or node_outputs(tree, x)
    preds, paths = [], []
    for b in range(x.shape[0]):
        w = tree.root
        path = []
        while tree.succ[w]:
            o = nodes[w]
            k = int(o["preds"][b])
            path.append((w, k, float(o["probs"][b, k]), float(o["entropy"][b])))
            w = tree.succ[w][k]
        preds.append(tree.class_index[w])
        paths.append(path)
    return np.array(preds, dtype=np.int64), paths


def hard_onehot(tree, x):
    """reference: nbdt/model.py:188-203 (predicted_to_logits: rows of eye(C), float32)."""
    pred, _ = hard_decisions(tree, x)
    return np.eye(tree.num_classes, dtype=np.float32)[pred]


def cross_entropy(z, y):
    """nn.CrossEntropyLoss (mean reduction) value and gradient w.r.t. z."""
    z = np.asarray(z, dtype=np.float64)
    B = z.shape[0]
    ls = _log_softmax(z)
    loss = -ls[np.arange(B), y].mean()
    g = np.exp(ls)
    g[np.arange(B), y] -= 1.0
    return loss, g / B


def schedule_weight(start, end, power, progress):
    """reference: nbdt/loss.py:187-189 (get_weight)."""
    end = start if end is None else end
    p = progress ** power
    return (1 - p) * start + p * end


def soft_tree_sup_loss(tree, x, y, xent_weight=1.0, tree_weight=1.0):
    """reference: nbdt/loss.py:191-203 + 264-266.

    L = xent_weight * CE(x, y) + tree_weight * CE(class_probs(x), y), the second CE taking
    the *probabilities* as if they were logits.  Returns (loss, dL/dx) in float64.
    The gradient is propagated node by node (chain rule through the per-node softmax and
    the leaf means), not via the fused closed form."""
    x = np.asarray(x, dtype=np.float64)
    nodes = node_outputs(tree, x)
    cp = soft_probs(tree, x, nodes)
    l1, g1 = cross_entropy(x, y)
    l2, gcp = cross_entropy(cp, y)
    gx = xent_weight * g1
    # d cp[:,c] / d probs_n[:,k] = cp[:,c] / probs_n[:,k] for c under child k
    for w in tree.inodes:
        p = nodes[w]["probs"]
        gp = np.zeros_like(p)
        for k, idx in enumerate(tree.child_classes[w]):
            # cp/p computed as a product over the other nodes to stay finite when p -> 0
            gp[:, k] = (gcp[:, idx] * (cp[:, idx] / p[:, k : k + 1])).sum(axis=1)
        gnl = p * (gp - (gp * p).sum(axis=1, keepdims=True))  # softmax backward
        for k, idx in enumerate(tree.child_classes[w]):
            gx[:, idx] += tree_weight * gnl[:, k : k + 1] / len(idx)  # mean backward
    return xent_weight * l1 + tree_weight * l2, gx


def hard_tree_sup_loss(tree, x, y, xent_weight=1.0, tree_weight=1.0, tree_supervision_weight=1.0):
    """reference: nbdt/loss.py:212-257 + 191-203.

    For every inner node, rows whose target lies under it are selected
    (model.py:127-143), grouped by child count, one CE(mean) per group scaled by
    rows_in_group / (B*N/2) * tree_supervision_weight; the outer forward multiplies by the
    (scheduled) tree weight again.  Returns (loss, dL/dx) in float64."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y)
    B = x.shape[0]
    num_losses = B * len(tree.inodes) / 2.0
    l1, g1 = cross_entropy(x, y)
    gx = xent_weight * g1
    groups = {}
    for w in tree.inodes:
        child_of_class = {}
        for k, idx in enumerate(tree.child_classes[w]):
            for c in idx:
                child_of_class[c] = k
        rows = [b for b in range(B) if int(y[b]) in child_of_class]
        K = len(tree.child_classes[w])
        groups.setdefault(K, []).append((w, rows, [child_of_class[int(y[b])] for b in rows]))
    loss_tree = 0.0
    for K, items in groups.items():
        total = sum(len(rows) for _, rows, _ in items)
        if total == 0:
            continue
        fraction = total / num_losses * tree_supervision_weight
        for w, rows, tgt in items:
            if not rows:
                continue
            rows = np.array(rows)
            tgt = np.array(tgt)
            nl = np.stack([x[rows][:, idx].mean(axis=1) for idx in tree.child_classes[w]], axis=1)
            ls = _log_softmax(nl)
            loss_tree += fraction * (-ls[np.arange(len(rows)), tgt].sum() / total)
            gnl = np.exp(ls)
            gnl[np.arange(len(rows)), tgt] -= 1.0
            gnl *= tree_weight * fraction / total
            for k, idx in enumerate(tree.child_classes[w]):
                gx[np.ix_(rows, idx)] += gnl[:, k : k + 1] / len(idx)
    return xent_weight * l1 + tree_weight * loss_tree, gx
