"""WideResNet-28-10 for CIFAR (`wrn28_10_cifar10`, `wrn28_10_cifar100`) and the global-pool
variant `wrn28_10` the reference uses for TinyImagenet200 (nbdt/models/wideresnet.py:1-40).

The reference imports the architecture from the un-vendored, unpinned `pytorchcv`
(`pytorchcv.models.wrn_cifar.get_wrn_cifar(blocks=28, width_factor=10)`).  pytorchcv is not
available offline, so this is a RESTATEMENT of its published definition (parity with pytorchcv is
unpinned; anchors that can be checked offline: 36.48 M parameters for 10 classes, FC key
`output.weight` -- reference nbdt/graph.py:391 -- and the shipped `induced-wrn28_10*` hierarchies):

  features.init_block          conv3x3(3 -> 16), no BN/activation
  features.stage{1,2,3}.unit{1..4}   pre-activation residual units at 160 / 320 / 640 channels,
                               stride 2 at unit1 of stages 2 and 3:
      body.conv1: bn -> relu -> conv3x3(stride)      (its bn+relu output also feeds identity_conv)
      body.conv2: bn -> relu -> conv3x3
      identity_conv: conv1x1(stride) on the PRE-ACTIVATED input, when the shape changes
  features.post_activ          bn -> relu
  features.final_pool          8x8 average pool (global pool for `wrn28_10`)
  output                       Linear(640 -> num_classes)

Initialisation as pytorchcv: kaiming_uniform_ on every conv weight.
"""
import torch.nn as nn
import torch.nn.init as init

from nbdt._b200 import backbone as B
from nbdt.models.utils import get_pretrained_model

__all__ = ("wrn28_10", "wrn28_10_cifar10", "wrn28_10_cifar100")

model_urls = {
    ("wrn28_10", "TinyImagenet200"): "https://github.com/alvinwan/neural-backed-decision-trees/releases/download/0.0.1/ckpt-TinyImagenet200-wrn28_10.pth"
}


class PreConvBlock(nn.Module):
    """bn -> relu -> conv (keys `bn.*`, `conv.weight`)."""

    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.bn = nn.BatchNorm2d(cin)
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)


class PreResBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = PreConvBlock(cin, cout, 3, stride)
        self.conv2 = PreConvBlock(cout, cout, 3, 1)


class PreResUnit(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.resize_identity = (cin != cout) or (stride != 1)
        self.body = PreResBlock(cin, cout, stride)
        if self.resize_identity:
            self.identity_conv = nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False)

    def forward(self, x):
        if self.resize_identity:
            a = B.bn_act(x, self.body.conv1.bn, relu=True)  # pre-activation, shared with identity_conv
            identity = B.conv(a, self.identity_conv)
        else:  # the identity branch re-enters through the BN function so its gradient is fused into BN backward
            a, identity = B.bn_act(x, self.body.conv1.bn, relu=True, want_skip=True)
        y = B.conv(a, self.body.conv1.conv, bn_stats=True)   # -> body.conv2.bn
        y = B.bn_act(y, self.body.conv2.bn, relu=True)
        return B.conv(y, self.body.conv2.conv, residual=identity, bn_stats=True)   # -> the next unit's (or the final) BatchNorm


class PreResActivation(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.bn = nn.BatchNorm2d(channels)


class _Features(nn.Sequential):
    pass


class CIFARWRN(nn.Module):
    def __init__(self, channels, init_block_channels, in_channels=3, in_size=(32, 32), num_classes=10):
        super().__init__()
        self.in_size = in_size
        self.num_classes = num_classes
        self.features = _Features()
        self.features.add_module(
            "init_block", nn.Conv2d(in_channels, init_block_channels, kernel_size=3, stride=1, padding=1, bias=False)
        )
        cin = init_block_channels
        for i, stage_channels in enumerate(channels):
            stage = nn.Sequential()
            for j, cout in enumerate(stage_channels):
                stage.add_module(f"unit{j + 1}", PreResUnit(cin, cout, 2 if (j == 0 and i != 0) else 1))
                cin = cout
            self.features.add_module(f"stage{i + 1}", stage)
        self.features.add_module("post_activ", PreResActivation(cin))
        self.features.add_module("final_pool", nn.AvgPool2d(kernel_size=8, stride=1))
        self.output = nn.Linear(cin, num_classes)
        for module in self.modules():
            if isinstance(module, nn.Conv2d):
                init.kaiming_uniform_(module.weight)
                if module.bias is not None:
                    init.constant_(module.bias, 0)

    def forward(self, x):
        f = self.features
        out = B.stem(x, f.init_block, None)
        for name, stage in f.named_children():
            if name.startswith("stage"):
                for unit in stage:
                    out = unit(out)
        out = B.bn_act(out, f.post_activ.bn, relu=True)
        out = B.avg_pool(out, f.final_pool)
        return B.linear(out, self.output)


def get_wrn_cifar(num_classes, blocks, width_factor, model_name=None, pretrained=False, **kwargs):
    assert (blocks - 4) % 6 == 0
    layers = [(blocks - 4) // 6] * 3
    channels = [[c * width_factor] * n for c, n in zip([16, 32, 64], layers)]
    if pretrained:
        raise ValueError("pytorchcv pretrained weights are not available offline")
    return CIFARWRN(channels=channels, init_block_channels=16, num_classes=num_classes, **kwargs)


def wrn28_10_cifar10(num_classes=10, **kwargs):
    return get_wrn_cifar(num_classes=num_classes, blocks=28, width_factor=10, model_name="wrn28_10_cifar10", **kwargs)


def wrn28_10_cifar100(num_classes=100, **kwargs):
    return get_wrn_cifar(num_classes=num_classes, blocks=28, width_factor=10, model_name="wrn28_10_cifar100", **kwargs)


def wrn28_10(pretrained=False, progress=True, dataset="CIFAR10", num_classes=10, **kwargs):
    """wrn28_10 with a global average pool instead of the fixed 8x8 pool (64x64 inputs)."""
    model = get_wrn_cifar(num_classes=num_classes, blocks=28, width_factor=10, model_name="wrn28_10", **kwargs)
    model.features.final_pool = nn.AdaptiveAvgPool2d((1, 1))
    return get_pretrained_model("wrn28_10", dataset, model, model_urls, pretrained=pretrained, progress=progress)
