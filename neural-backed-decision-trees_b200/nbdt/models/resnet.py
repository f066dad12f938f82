"""CIFAR-style ResNets (3x3 stem, no max-pool, global average pool, `linear` classifier).

Architecture and parameter names follow the reference's nbdt/models/resnet.py (BasicBlock :42-74,
Bottleneck :77-112, ResNet :115-149, factories :152-223) so checkpoints and induced hierarchies
(FC key `linear.weight`) are interchangeable.  Layers are parameter containers; the arithmetic is
dispatched through nbdt._b200.backbone.
"""
import torch.nn as nn

from nbdt._b200 import backbone as B
from nbdt.models.utils import get_pretrained_model

__all__ = ("ResNet10", "ResNet18", "ResNet34", "ResNet50", "ResNet101", "ResNet152")

_RELEASE = "https://github.com/alvinwan/neural-backed-decision-trees/releases/download/0.0.1/"
model_urls = {
    ("ResNet10", "CIFAR10"): _RELEASE + "ckpt-CIFAR10-ResNet10.pth",
    ("ResNet10", "CIFAR100"): _RELEASE + "ckpt-CIFAR100-ResNet10.pth",
    ("ResNet18", "CIFAR10"): _RELEASE + "ckpt-CIFAR10-ResNet18.pth",
    ("ResNet18", "CIFAR100"): _RELEASE + "ckpt-CIFAR100-ResNet18.pth",
    ("ResNet18", "TinyImagenet200"): _RELEASE + "ckpt-TinyImagenet200-ResNet18.pth",
}


def _conv(cin, cout, k, stride):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)


def _shortcut(cin, cout, stride):
    """Identity, or 1x1 conv + BN when the shape changes (keys `shortcut.0`, `shortcut.1`)."""
    if stride == 1 and cin == cout:
        return nn.Sequential()
    return nn.Sequential(_conv(cin, cout, 1, stride), nn.BatchNorm2d(cout))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = _conv(in_planes, planes, 3, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, 1)
        self.bn2 = nn.BatchNorm2d(planes)
        self.shortcut = _shortcut(in_planes, self.expansion * planes, stride)

    def forward(self, x):
        out = B.conv_bn_act(x, self.conv1, self.bn1, relu=True)
        res = B.conv_bn_act(x, self.shortcut[0], self.shortcut[1], relu=False) if len(self.shortcut) else x
        return B.conv_bn_act(out, self.conv2, self.bn2, relu=True, residual=res)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = _conv(in_planes, planes, 1, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, stride)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = _conv(planes, self.expansion * planes, 1, 1)
        self.bn3 = nn.BatchNorm2d(self.expansion * planes)
        self.shortcut = _shortcut(in_planes, self.expansion * planes, stride)

    def forward(self, x):
        out = B.conv_bn_act(x, self.conv1, self.bn1, relu=True)
        out = B.conv_bn_act(out, self.conv2, self.bn2, relu=True)
        res = B.conv_bn_act(x, self.shortcut[0], self.shortcut[1], relu=False) if len(self.shortcut) else x
        return B.conv_bn_act(out, self.conv3, self.bn3, relu=True, residual=res)


class ResNet(nn.Module):
    def __init__(self, block, num_blocks, num_classes=10):
        super().__init__()
        self.in_planes = 64
        self.conv1 = _conv(3, 64, 3, 1)
        self.bn1 = nn.BatchNorm2d(64)
        for i, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2))):
            setattr(self, f"layer{i + 1}", self._make_layer(block, planes, num_blocks[i], stride))
        self.linear = nn.Linear(512 * block.expansion, num_classes)

    def _make_layer(self, block, planes, num_blocks, stride):
        layers = []
        for s in [stride] + [1] * (num_blocks - 1):
            layers.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*layers)

    def features(self, x):
        out = B.stem(x, self.conv1, self.bn1)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            out = layer(out)
        return B.global_avg_pool(out)

    def forward(self, x):
        return B.linear(self.features(x), self.linear)


def _ResNet(arch, *args, pretrained=False, progress=True, dataset="CIFAR10", **kwargs):
    model = ResNet(*args, **kwargs)
    return get_pretrained_model(arch, dataset, model, model_urls, pretrained=pretrained, progress=progress)


def _factory(arch, block, depths):
    def make(pretrained=False, progress=True, **kwargs):
        return _ResNet(arch, block, depths, pretrained=pretrained, progress=progress, **kwargs)

    make.__name__ = make.__qualname__ = arch
    return make


ResNet10 = _factory("ResNet10", BasicBlock, [1, 1, 1, 1])
ResNet18 = _factory("ResNet18", BasicBlock, [2, 2, 2, 2])
ResNet34 = _factory("ResNet34", BasicBlock, [3, 4, 6, 3])
ResNet50 = _factory("ResNet50", Bottleneck, [3, 4, 6, 3])
ResNet101 = _factory("ResNet101", Bottleneck, [3, 4, 23, 3])
ResNet152 = _factory("ResNet152", Bottleneck, [3, 8, 36, 3])
