"""TEST INFRASTRUCTURE ONLY -- torch-CPU port of the reference's train step, in the reference's own
op structure (one group of tiny tensor ops per inner node, autograd for the backward).

Used (a) by bench.py's `cpu_baseline` / `--impl reference` arm, timed on the GPU box's host cores
(`/root/reference` cannot travel there), and (b) by tests as the fp32 eager reference for the
backbones.  The product never imports this module.

Parity status: the tree functions are pinned by tests/test_oracle_golden.py::test_torch_port_*
against the executing reference's golden vectors; ResNet18 is pinned by the seeded forward/train-step
vectors in tests/golden/backbone_resnet18.npz; WRN-28-10 is a restatement of pytorchcv's
definition and is UNPINNED (pytorchcv is not installed anywhere offline).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Categorical


# ---- tree head / losses, node by node (reference nbdt/model.py, nbdt/loss.py) -------------------
def node_logits(outputs, child_classes):
    """reference: nbdt/model.py:94-99 -- mean of the transposed rows, stacked, transposed back."""
    return torch.stack([outputs.T[idx].mean(dim=0) for idx in child_classes]).T


def all_node_outputs(tree, outputs):
    """reference: nbdt/model.py:101-120."""
    out = {}
    for w in tree.inodes:
        nl = node_logits(outputs, tree.child_classes[w])
        probs = F.softmax(nl, dim=1)
        out[w] = {"logits": nl, "preds": torch.max(nl, dim=1)[1], "probs": probs,
                  "entropy": Categorical(probs=probs).entropy()}
    return out


def soft_rules(tree, outputs):
    """reference: nbdt/model.py:207-242 + 268-273 (float32 buffer of ones, in-place indexed multiply)."""
    w2o = all_node_outputs(tree, outputs)
    class_probs = torch.ones((outputs.size(0), tree.num_classes))
    for w in tree.inodes:
        old, new = [], []
        for k, idx in enumerate(tree.child_classes[w]):
            old.extend(idx)
            new.extend([k] * len(idx))
        class_probs[:, old] *= w2o[w]["probs"][:, new]
    return class_probs


def soft_tree_sup_loss(tree, criterion, outputs, targets, xent_weight=1.0, tree_weight=1.0):
    """reference: nbdt/loss.py:191-203, 264-266."""
    return criterion(outputs, targets) * xent_weight + criterion(soft_rules(tree, outputs), targets) * tree_weight


def hard_tree_sup_loss(tree, criterion, outputs, targets, xent_weight=1.0, tree_weight=1.0, tsw=1.0):
    """reference: nbdt/loss.py:212-257 (rows selected per node, grouped by child count)."""
    num_losses = outputs.size(0) * len(tree.inodes) / 2.0
    targets_int = [int(v) for v in targets.cpu().long()]
    subs, tgts = {}, {}
    for w in tree.inodes:
        child_of = {c: k for k, idx in enumerate(tree.child_classes[w]) for c in idx}
        selector = [v in child_of for v in targets_int]
        sub = outputs[selector]
        key = len(tree.child_classes[w])
        if sub.size(0):
            sub = node_logits(sub, tree.child_classes[w])
        else:
            sub = sub[:, :key]
        subs.setdefault(key, []).append(sub)
        tgts.setdefault(key, []).extend(child_of[v] for v in targets_int if v in child_of)
    loss = 0
    for key in subs:
        sub = torch.cat(subs[key], dim=0)
        if not sub.size(0):
            continue
        fraction = sub.size(0) / float(num_losses) * tsw
        loss = loss + criterion(sub, torch.tensor(tgts[key]).long()) * fraction
    return criterion(outputs, targets) * xent_weight + loss * tree_weight


# ---- eager backbones (parameter names match nbdt.models so state_dicts are interchangeable) ------
class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.shortcut = nn.Sequential()
        if stride != 1 or cin != cout:
            self.shortcut = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + self.shortcut(x))


class _Bottleneck(nn.Module):
    """reference: nbdt/models/resnet.py:77-112 (1x1 -> 3x3 (stride) -> 1x1 x4, projection shortcut when the shape changes)."""

    expansion = 4

    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, 4 * planes, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(4 * planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or cin != 4 * planes:
            self.shortcut = nn.Sequential(nn.Conv2d(cin, 4 * planes, 1, stride, bias=False), nn.BatchNorm2d(4 * planes))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return F.relu(out + self.shortcut(x))


class EagerResNet(nn.Module):
    """reference: nbdt/models/resnet.py:115-149 with BasicBlock (:42-74) or Bottleneck (:77-112)."""

    def __init__(self, depths=(2, 2, 2, 2), num_classes=10, bottleneck=False):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        exp = 4 if bottleneck else 1
        for i, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2))):
            blocks = []
            for s in [stride] + [1] * (depths[i] - 1):
                blocks.append(_Bottleneck(cin, planes, s) if bottleneck else _BasicBlock(cin, planes, s))
                cin = planes * exp
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        self.linear = nn.Linear(512 * exp, num_classes)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        for i in range(4):
            out = getattr(self, f"layer{i + 1}")(out)
        out = F.avg_pool2d(out, out.size()[2:]).flatten(1)
        return self.linear(out)


class _PreConv(nn.Module):
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.bn = nn.BatchNorm2d(cin)
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)


class _PreBody(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = _PreConv(cin, cout, 3, stride)
        self.conv2 = _PreConv(cout, cout, 3, 1)


class _PreUnit(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.resize = cin != cout or stride != 1
        self.body = _PreBody(cin, cout, stride)
        if self.resize:
            self.identity_conv = nn.Conv2d(cin, cout, 1, stride, bias=False)

    def forward(self, x):
        a = F.relu(self.body.conv1.bn(x))
        identity = self.identity_conv(a) if self.resize else x
        y = self.body.conv1.conv(a)
        y = self.body.conv2.conv(F.relu(self.body.conv2.bn(y)))
        return y + identity


class _PostActiv(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.bn = nn.BatchNorm2d(c)


def EagerResNet50(num_classes=10):
    return EagerResNet(depths=(3, 4, 6, 3), num_classes=num_classes, bottleneck=True)


class EagerWRN(nn.Module):
    """pytorchcv CIFARWRN(blocks=28, width_factor=10) restated (see nbdt/models/wideresnet.py docstring)."""

    def __init__(self, num_classes=10, widths=(160, 320, 640), units=4, global_pool=False):
        super().__init__()
        self.features = nn.Sequential()
        self.features.add_module("init_block", nn.Conv2d(3, 16, 3, 1, 1, bias=False))
        cin = 16
        for i, w in enumerate(widths):
            stage = nn.Sequential()
            for j in range(units):
                stage.add_module(f"unit{j + 1}", _PreUnit(cin, w, 2 if (j == 0 and i != 0) else 1))
                cin = w
            self.features.add_module(f"stage{i + 1}", stage)
        self.features.add_module("post_activ", _PostActiv(cin))
        self.features.add_module("final_pool", nn.AdaptiveAvgPool2d(1) if global_pool else nn.AvgPool2d(8, 1))
        self.output = nn.Linear(cin, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight)

    def forward(self, x):
        f = self.features
        out = f.init_block(x)
        for name, stage in f.named_children():
            if name.startswith("stage"):
                out = stage(out)
        out = F.relu(f.post_activ.bn(out))
        return self.output(f.final_pool(out).flatten(1))


def train_step(net, opt, loss_fn, x, y):
    """reference: main.py:235-239."""
    opt.zero_grad()
    out = net(x)
    loss = loss_fn(out, y)
    loss.backward()
    opt.step()
    return loss
