"""EfficientNet-B0 on the B200 kernels (inference path).

The reference has no EfficientNet code of its own: `nbdt/models/__init__.py:3-4` star-imports
`pytorchcv.models.efficientnet` and `torchvision.models`, and the ImageNet NBDTs wrap those modules
(`nbdt/model.py:281-361`).  This module keeps torchvision's `efficientnet_b0` parameter layout -- same module tree and
state_dict keys (`features.0.0.weight`, `features.2.0.block.1.0.weight`, `classifier.1.weight`, ...), so torchvision
checkpoints load with `load_state_dict(strict=True)` -- and runs the forward pass through nbdt-b200 kernels:

  stem      3x3/s2 conv (csrc/bn_ops.cu: stem_fwd_kernel) -> folded BN + SiLU (bn_apply_kernel)
  MBConv    [1x1 expand conv on tcgen05 (csrc/conv_gemm.cu) -> BN + SiLU]
            -> depthwise k x k conv with BN + SiLU fused (csrc/mbconv_ops.cu: dwconv_fwd_kernel)
            -> squeeze-excitation: global average pool kernel, two tiny FC layers (SGEMM kernel of csrc/linear.cu), scale kernel
            -> 1x1 project conv on tcgen05 -> BN (+ residual) in one pass
  head      1x1 conv 320 -> 1280 -> BN + SiLU -> global average pool -> Linear (SGEMM kernel of csrc/linear.cu)

Activations are NHWC bf16 with channel counts padded to multiples of 64.  Eval mode: folded BatchNorm (running statistics),
depthwise conv + BN + SiLU in one kernel, cached packed weights; StochasticDepth / Dropout are identities.  Training mode:
batch-statistics BatchNorm + SiLU through nn_ops.BatchNormActFunction (SiLU gradient recomputed in the backward kernels),
1x1 convs through nn_ops.ConvFunction (tcgen05 fwd / dgrad / wgrad), depthwise conv and squeeze-excitation through
mbconv_ops.DepthwiseConvFunction / SEScaleFunction, row-mode stochastic depth as a per-image scale, dropout on the pooled
features.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from nbdt._b200 import backbone as B
from nbdt._b200 import conv_ops as C
from nbdt._b200 import mbconv_ops as M
from nbdt._b200.nn_ops import AvgPoolFunction, LinearFunction

__all__ = ("efficientnet_b0", "EfficientNet")

# torchvision.models.efficientnet._efficientnet_conf("efficientnet_b0"): expand, kernel, stride, in, out, layers
_B0 = ((1, 3, 1, 32, 16, 1), (6, 3, 2, 16, 24, 2), (6, 5, 2, 24, 40, 2), (6, 3, 2, 40, 80, 3),
       (6, 5, 1, 80, 112, 3), (6, 5, 2, 112, 192, 4), (6, 3, 1, 192, 320, 1))


class ConvNormAct(nn.Sequential):
    """Conv2d (no bias) + BatchNorm2d (+ SiLU): keys `0.weight`, `1.*` like torchvision's Conv2dNormActivation."""

    def __init__(self, cin, cout, k=3, stride=1, groups=1, act=True):
        layers = [nn.Conv2d(cin, cout, k, stride, padding=(k - 1) // 2, groups=groups, bias=False), nn.BatchNorm2d(cout)]
        if act:
            layers.append(nn.SiLU(inplace=True))
        super().__init__(*layers)
        self.has_act = act


class SqueezeExcitation(nn.Module):
    def __init__(self, channels, squeeze):
        super().__init__()
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, squeeze, 1)
        self.fc2 = nn.Conv2d(squeeze, channels, 1)
        self.activation = nn.SiLU(inplace=True)
        self.scale_activation = nn.Sigmoid()

    def forward(self, x):
        """x: NHWC bf16 (N, H, W, C_phys) -> x * sigmoid(fc2(silu(fc1(mean_hw x)))).  Differentiable: the pooling and the
        scaling are own kernels with backward kernels, the two tiny FC layers run on the small SGEMM of csrc/linear.cu."""
        c = self.fc1.in_channels
        pooled = AvgPoolFunction.apply(x)  # (N, C_phys) fp32
        s = LinearFunction.apply(pooled, self.fc1.weight.flatten(1), self.fc1.bias)
        s = LinearFunction.apply(F.silu(s), self.fc2.weight.flatten(1), self.fc2.bias)
        scale = F.pad(torch.sigmoid(s), (0, pooled.shape[1] - c))  # padding channels stay zero
        return M.SEScaleFunction.apply(x, scale)


class MBConv(nn.Module):
    def __init__(self, expand, k, stride, cin, cout, sd_prob):
        super().__init__()
        self.use_res_connect = stride == 1 and cin == cout
        layers = []
        mid = cin * expand
        if mid != cin:
            layers.append(ConvNormAct(cin, mid, 1))
        layers.append(ConvNormAct(mid, mid, k, stride, groups=mid))
        layers.append(SqueezeExcitation(mid, max(1, cin // 4)))
        layers.append(ConvNormAct(mid, cout, 1, act=False))
        self.block = nn.Sequential(*layers)
        self.stochastic_depth = nn.Identity()  # eval-mode identity (torchvision: StochasticDepth(sd_prob, "row"))
        self.sd_prob = sd_prob

    def forward(self, x):
        y = x
        mods = list(self.block)
        dw, se, proj = mods[-3], mods[-2], mods[-1]
        if not self.training:
            if len(mods) == 4:
                y = M.bn_act_eval(M.conv1x1(y, mods[0][0]), mods[0][1], M.ACT_SILU)
            y = M.dwconv_bn_act(y, dw[0], dw[1], M.ACT_SILU)
            y = se(y)
            y = M.conv1x1(y, proj[0])
            return M.bn_act_eval(y, proj[1], M.ACT_NONE, residual=x if self.use_res_connect else None)
        # training: batch-statistics BatchNorm, autograd functions over the fwd / dgrad / wgrad kernels
        if len(mods) == 4:
            y = B.bn_act(B.conv(y, mods[0][0]), mods[0][1], relu=M.ACT_SILU)
        conv = dw[0]
        y = M.DepthwiseConvFunction.apply(y, conv.weight, conv.stride[0])
        y = B.bn_act(y, dw[1], relu=M.ACT_SILU)
        y = se(y)
        y = B.conv(y, proj[0])
        if not self.use_res_connect:
            return B.bn_act(y, proj[1], relu=M.ACT_NONE)
        if self.sd_prob <= 0.0:
            return B.bn_act(y, proj[1], relu=M.ACT_NONE, residual=x)  # residual add fused into the BN apply pass
        # stochastic depth, "row" mode (torchvision.ops.StochasticDepth): drop the whole branch per image, rescale the rest
        y = B.bn_act(y, proj[1], relu=M.ACT_NONE)
        keep = 1.0 - self.sd_prob
        noise = torch.empty(y.shape[0], device=y.device, dtype=torch.float32).bernoulli_(keep) / keep
        y = M.SEScaleFunction.apply(y, noise[:, None].expand(y.shape[0], y.shape[3]).contiguous())
        return y + x


class EfficientNet(nn.Module):
    def __init__(self, setting=_B0, num_classes=1000, dropout=0.2, stochastic_depth_prob=0.2, last_channel=None):
        super().__init__()
        features = [ConvNormAct(3, setting[0][3], 3, 2)]
        total = sum(s[5] for s in setting)
        block_id = 0
        for expand, k, stride, cin, cout, layers in setting:
            stage = []
            for i in range(layers):
                sd = stochastic_depth_prob * float(block_id) / total
                stage.append(MBConv(expand, k, stride if i == 0 else 1, cin if i == 0 else cout, cout, sd))
                block_id += 1
            features.append(nn.Sequential(*stage))
        last_in = setting[-1][4]
        last_channel = last_channel or 4 * last_in
        features.append(ConvNormAct(last_in, last_channel, 1))
        self.features = nn.Sequential(*features)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Sequential(nn.Dropout(p=dropout, inplace=True), nn.Linear(last_channel, num_classes))
        for m in self.modules():  # torchvision's initialisation
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                r = 1.0 / math.sqrt(m.out_features)
                nn.init.uniform_(m.weight, -r, r)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        f = self.features
        stem = f[0]
        if self.training:
            y = B.stem(x, stem[0], None)  # strided 3x3 conv, own kernels forward and backward
            y = B.bn_act(y, stem[1], relu=M.ACT_SILU)
        else:
            y = M.stem(x, stem[0], C.pad64(stem[0].out_channels))
            y = M.bn_act_eval(y, stem[1], M.ACT_SILU)
        for stage in list(f)[1:-1]:
            for block in stage:
                y = block(y)
        head = f[-1]
        if self.training:
            y = B.bn_act(B.conv(y, head[0]), head[1], relu=M.ACT_SILU)
        else:
            y = M.bn_act_eval(M.conv1x1(y, head[0]), head[1], M.ACT_SILU)
        feat = AvgPoolFunction.apply(y)  # (N, C_phys) fp32
        lin = self.classifier[1]
        feat = feat[:, : lin.in_features]
        if self.training:
            feat = F.dropout(feat, p=self.classifier[0].p, training=True)
        return LinearFunction.apply(feat, lin.weight, lin.bias)


def efficientnet_b0(pretrained=False, progress=True, num_classes=1000, weights=None, **kwargs):
    """torchvision-compatible factory.  Pretrained weights need network access (torchvision's hub URL)."""
    model = EfficientNet(_B0, num_classes=num_classes, **kwargs)
    if pretrained or weights is not None:
        from torchvision.models import EfficientNet_B0_Weights

        w = EfficientNet_B0_Weights.IMAGENET1K_V1 if (pretrained or weights in (True, "DEFAULT", "IMAGENET1K_V1")) else weights
        model.load_state_dict(w.get_state_dict(progress=progress, check_hash=True))
    return model
