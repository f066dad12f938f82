"""NBDT inference: embedded decision rules and the model wrappers.

Same public surface as the reference's nbdt/model.py -- `EmbeddedDecisionRules`,
`HardEmbeddedDecisionRules`, `SoftEmbeddedDecisionRules` (model.py:65-273) and `NBDT`,
`HardNBDT`, `SoftNBDT`, `SegNBDT`, `HardSegNBDT`, `SoftSegNBDT` (model.py:281-399) -- but
the per-node Python loops (forward_nodes, traverse_tree) are single fused CUDA kernels
(csrc/tree.cu through include/nbdt_b200.h).  The rules take the *logits* of the wrapped
network; outputs carry `_nbdt_output_flag` so the losses can refuse them (loss.py).
"""
import torch
import torch.nn as nn

from nbdt._b200 import ops
from nbdt.models.utils import coerce_state_dict, load_state_dict_from_key
from nbdt.tree import Node, Tree  # noqa: F401  (re-exported like the reference)
from nbdt.utils import coerce_tensor, uncoerce_tensor

_RELEASE = "https://github.com/alvinwan/neural-backed-decision-trees/releases/download/0.0.1/"
model_urls = {
    ("ResNet18", "CIFAR10"): _RELEASE + "ckpt-CIFAR10-ResNet18-induced-ResNet18-SoftTreeSupLoss.pth",
    ("wrn28_10_cifar10", "CIFAR10"): _RELEASE + "ckpt-CIFAR10-wrn28_10_cifar10-induced-wrn28_10_cifar10-SoftTreeSupLoss.pth",
    ("wrn28_10_cifar10", "CIFAR10", "wordnet"): _RELEASE + "ckpt-CIFAR10-wrn28_10_cifar10-wordnet-SoftTreeSupLoss.pth",
    ("ResNet18", "CIFAR100"): _RELEASE + "ckpt-CIFAR100-ResNet18-induced-ResNet18-SoftTreeSupLoss.pth",
    ("wrn28_10_cifar100", "CIFAR100"): _RELEASE + "ckpt-CIFAR100-wrn28_10_cifar100-induced-wrn28_10_cifar100-SoftTreeSupLoss.pth",
    ("ResNet18", "TinyImagenet200"): _RELEASE + "ckpt-TinyImagenet200-ResNet18-induced-ResNet18-SoftTreeSupLoss-tsw10.0.pth",
    ("wrn28_10", "TinyImagenet200"): _RELEASE + "ckpt-TinyImagenet200-wrn28_10-induced-wrn28_10-SoftTreeSupLoss-tsw10.0.pth",
}


def _flag(t):
    t._nbdt_output_flag = True  # checked by the NBDT losses (loss.py), to prevent mistakes
    return t


#########
# RULES #
#########


class EmbeddedDecisionRules(nn.Module):
    def __init__(self, dataset=None, path_graph=None, path_wnids=None, classes=(), hierarchy=None, tree=None):
        super().__init__()
        if not tree:
            tree = Tree(dataset, path_graph, path_wnids, classes, hierarchy=hierarchy)
        self.tree = tree
        self.correct = 0
        self.total = 0
        self.I = torch.eye(len(self.tree.classes))

    @staticmethod
    def get_node_logits(outputs, node=None, new_to_old_classes=None, num_classes=None):
        """Logits of ONE node: column k is the mean of `outputs` over the classes under child k.

        With `node`, the columns are sliced out of the fused node-outputs kernel; with an explicit
        mapping (no node) they are gathered on the device."""
        assert node or (new_to_old_classes and num_classes), (
            "Either pass node or (new_to_old_classes mapping and num_classes)"
        )
        if node is not None and not new_to_old_classes and not num_classes and not node.has_other:
            nl, _, _, _, dt = ops.node_outputs_autograd(node.tree, outputs)
            i = [n.wnid for n in node.tree.inodes].index(node.wnid)
            return nl[:, int(dt.child_ptr[i]) : int(dt.child_ptr[i + 1])].to(outputs.dtype)
        ops.L.require_cuda(outputs, "get_node_logits")
        new_to_old_classes = new_to_old_classes or node.child_index_to_class_index
        num_classes = num_classes or node.num_classes
        cols = [outputs[:, list(new_to_old_classes[k])].mean(dim=1) for k in range(num_classes)]
        return torch.stack(cols, dim=1)

    @classmethod
    def get_all_node_outputs(cls, outputs, nodes):
        """{wnid: {"logits", "preds", "probs", "entropy"}} for every node in `nodes` (one kernel
        launch for the whole hierarchy; the dict holds views into its packed outputs)."""
        nodes = list(nodes)
        if not nodes:
            return {}
        tree = nodes[0].tree
        nl, pr, preds, ent, dt = ops.node_outputs_autograd(tree, outputs)
        index = {n.wnid: i for i, n in enumerate(tree.inodes)}
        out = {}
        for node in nodes:
            i = index[node.wnid]
            b, e = int(dt.child_ptr[i]), int(dt.child_ptr[i + 1])
            out[node.wnid] = {
                "logits": nl[:, b:e],
                "preds": preds[:, i],
                "probs": pr[:, b:e],
                "entropy": ent[:, i],
            }
        return out

    def forward_nodes(self, outputs):
        return self.get_all_node_outputs(outputs, self.tree.inodes)


class HardEmbeddedDecisionRules(EmbeddedDecisionRules):
    @classmethod
    def get_node_logits_filtered(cls, node, outputs, targets):
        """Rows of `outputs` whose target lies under `node`, their node logits and child-index
        targets: (selector, outputs_sub, targets_sub)."""
        classes = [node.class_index_to_child_index[int(t)] for t in targets]
        selector = [bool(c) for c in classes]
        targets_sub = [c[0] for c in classes if c]
        outputs = outputs[torch.tensor(selector, dtype=torch.bool, device=outputs.device)]
        if outputs.size(0) == 0:
            return selector, outputs[:, : node.num_classes], targets_sub
        return selector, cls.get_node_logits(outputs, node), targets_sub

    @staticmethod
    def _decisions(tree, pred, paths):
        """Host-side list-of-dicts view of the kernel's path arrays (one D2H copy)."""
        child, inode, prob, ent = (t.cpu() for t in paths)
        child, inode, prob, ent = child.tolist(), inode.tolist(), prob.tolist(), ent.tolist()
        inodes = tree.inodes
        kids = [node.children for node in inodes]
        root = tree.root
        decisions = []
        for b in range(len(child)):
            steps = [{"node": root, "name": "root", "prob": 1, "entropy": 0}]
            for k, i, pr, en in zip(child[b], inode[b], prob[b], ent[b]):
                if k < 0:
                    break
                node = kids[i][k]
                steps.append({"node": node, "name": node.name, "prob": pr, "next_index": k, "entropy": en})
            decisions.append(steps)
        return decisions

    @classmethod
    def traverse_tree(cls, wnid_to_outputs, tree):
        """Greedy walk driven by precomputed node outputs (compatibility entry point; the model
        path uses the fused kernel on the logits directly).  Returns (preds, decisions)."""
        example = wnid_to_outputs[tree.inodes[0].wnid]
        device = example["logits"].device
        B = int(example["logits"].size(0))
        host = {
            w: (o["preds"].cpu().tolist(), o["probs"].detach().float().cpu().tolist(), o["entropy"].cpu().tolist())
            for w, o in wnid_to_outputs.items()
        }
        preds, decisions = [], []
        for b in range(B):
            node = tree.root
            steps = [{"node": node, "name": "root", "prob": 1, "entropy": 0}]
            while not node.is_leaf():
                if node.wnid not in host:
                    node = None
                    break
                p, pr, en = host[node.wnid]
                k = p[b]
                node = node.children[k]
                steps.append({"node": node, "name": node.name, "prob": pr[b][k], "next_index": k, "entropy": en[b]})
            preds.append(tree.wnid_to_class_index[node.wnid])
            decisions.append(steps)
        return torch.tensor(preds, dtype=torch.long, device=device), decisions

    def predicted_to_logits(self, predicted):
        """Predicted classes -> one-hot float32 rows."""
        if self.I.device != predicted.device:
            self.I = self.I.to(predicted.device)
        return self.I[predicted]

    def forward_with_decisions(self, outputs):
        predicted, paths = ops.hard_forward(self.tree, outputs, with_paths=True)
        decisions = self._decisions(self.tree, predicted, paths)
        return _flag(self.predicted_to_logits(predicted)), decisions

    def forward(self, outputs):
        predicted, _ = ops.hard_forward(self.tree, outputs, with_paths=False)
        return _flag(self.predicted_to_logits(predicted))


class SoftEmbeddedDecisionRules(EmbeddedDecisionRules):
    @classmethod
    def traverse_tree(cls, wnid_to_outputs, tree):
        """Class probabilities from precomputed node outputs (compatibility entry point): the
        product, over the inner nodes above a class, of the probability of the child it sits
        under.  The model path computes the same thing from the logits in one kernel."""
        example = wnid_to_outputs[tree.inodes[0].wnid]
        probs = example["probs"]
        ops.L.require_cuda(probs, "SoftEmbeddedDecisionRules.traverse_tree")
        B = probs.size(0)
        log_cp = torch.zeros((B, len(tree.classes)), dtype=torch.float32, device=probs.device)
        for node in tree.inodes:
            node_probs = wnid_to_outputs[node.wnid]["probs"].float()
            old, new = [], []
            for k in range(len(node.children)):
                idx = node.child_index_to_class_index[k]
                old.extend(idx)
                new.extend([k] * len(idx))
            assert len(set(old)) == len(old), "All old indices must be unique in order for this operation to be correct."
            log_cp[:, old] += torch.log(node_probs[:, new])
        return torch.exp(log_cp)

    def forward_with_decisions(self, outputs):
        """(class_probs, decisions).  As in the reference (model.py:244-266) the `prob` / `entropy`
        of every step are read from SAMPLE 0's node distribution (model.py:259)."""
        probs = self.forward(outputs)
        predicted = probs.max(1)[1].cpu().tolist()
        nl, pr, _, ent, dt = ops.node_outputs(self.tree, outputs[:1])
        pr0, ent0 = pr[0].cpu().tolist(), ent[0].cpu().tolist()
        index = {n.wnid: i for i, n in enumerate(self.tree.inodes)}
        leaf_to_steps = self.tree.get_leaf_to_steps()
        decisions = []
        for prediction in predicted:
            template = leaf_to_steps[self.tree.wnids_leaves[prediction]]
            steps = [dict(step) for step in template]
            steps[0]["prob"], steps[0]["entropy"] = 1.0, 0.0
            for prev, step in zip(steps[:-1], steps[1:]):
                i = index[prev["node"].wnid]
                step["prob"] = float(pr0[int(dt.child_ptr[i]) + prev["next_index"]])
                step["entropy"] = float(ent0[i])
            decisions.append(steps)
        return probs, decisions

    def forward(self, outputs, wnid_to_outputs=None):
        if wnid_to_outputs:
            return _flag(self.traverse_tree(wnid_to_outputs, self.tree))
        if outputs.requires_grad and torch.is_grad_enabled():
            return _flag(ops.SoftRulesFunction.apply(outputs, self.tree))
        return _flag(ops.soft_forward(self.tree, outputs))


##########
# MODELS #
##########


class NBDT(nn.Module):
    def __init__(self, dataset, model, arch=None, path_graph=None, path_wnids=None, classes=None,
                 hierarchy=None, pretrained=None, **kwargs):
        super().__init__()
        if dataset and not hierarchy and not path_graph:
            assert arch, "Must specify `arch` if no `hierarchy` or `path_graph`"
            hierarchy = f"induced-{arch}"
        if pretrained and not arch:
            raise UserWarning(
                "To load a pretrained NBDT, you need to specify the `arch`. "
                "`arch` is the name of the architecture. e.g., ResNet18"
            )
        if isinstance(model, str):
            raise NotImplementedError("Model must be nn.Module")
        tree = Tree(dataset, path_graph, path_wnids, classes, hierarchy=hierarchy)
        self.init(dataset, model, tree, arch=arch, pretrained=pretrained, hierarchy=hierarchy, **kwargs)

    def init(self, dataset, model, tree, arch=None, pretrained=False, hierarchy=None, eval=True,
             Rules=HardEmbeddedDecisionRules):
        """Second-stage constructor: everything the wrapper finally needs (tree already built)."""
        self.rules = Rules(tree=tree)
        self.model = model
        if pretrained:
            assert arch is not None
            keys = [(arch, dataset), (arch, dataset, hierarchy)]
            self.load_state_dict(load_state_dict_from_key(keys, model_urls, pretrained=True))
        if eval:
            self.eval()

    def load_state_dict(self, state_dict, **kwargs):
        state_dict = coerce_state_dict(state_dict, self.model.state_dict())
        return self.model.load_state_dict(state_dict, **kwargs)

    def state_dict(self, *args, **kwargs):
        return self.model.state_dict(*args, **kwargs)

    def forward(self, x):
        return self.rules(self.model(x))

    def forward_with_decisions(self, x):
        return self.rules.forward_with_decisions(self.model(x))


class HardNBDT(NBDT):
    def __init__(self, *args, **kwargs):
        kwargs.update({"Rules": HardEmbeddedDecisionRules})
        super().__init__(*args, **kwargs)


class SoftNBDT(NBDT):
    def __init__(self, *args, **kwargs):
        kwargs.update({"Rules": SoftEmbeddedDecisionRules})
        super().__init__(*args, **kwargs)


class SegNBDT(NBDT):
    """Per-pixel NBDT: (N,C,H,W) logits are treated as N*H*W rows."""

    def forward(self, x):
        assert len(x.shape) == 4, "Input must be of shape (N,C,H,W) for segmentation"
        x = self.model(x)
        shape = x.shape
        if x.is_cuda and not (x.requires_grad and torch.is_grad_enabled()):
            # the (N, C, H, W) logits are addressed in place: no permute + reshape copies around the rules
            if isinstance(self.rules, SoftEmbeddedDecisionRules):
                return _flag(ops.soft_forward_nchw(self.rules.tree, x))
            if isinstance(self.rules, HardEmbeddedDecisionRules):
                pred = ops.hard_forward_nchw(self.rules.tree, x)
                return _flag(torch.nn.functional.one_hot(pred, shape[1]).permute(0, 3, 1, 2).to(torch.float32))
        return uncoerce_tensor(self.rules.forward(coerce_tensor(x)), shape)


class HardSegNBDT(SegNBDT):
    def __init__(self, *args, **kwargs):
        kwargs.update({"Rules": HardEmbeddedDecisionRules})
        super().__init__(*args, **kwargs)


class SoftSegNBDT(SegNBDT):
    def __init__(self, *args, **kwargs):
        kwargs.update({"Rules": SoftEmbeddedDecisionRules})
        super().__init__(*args, **kwargs)
