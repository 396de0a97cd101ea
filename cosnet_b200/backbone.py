"""Encoders of the `raa` model: dilated ResNet (stride 8) + ASPP.  cuDNN convolutions, OUT of the
accelerated hot path; they exist here so that the drop-in module has the
same sub-module tree -- and therefore the same state_dict keys -- as the reference
(deeplab/residual_net.py:47-172, deeplab/deeplabv3_encoder.py:10-185).

Module / attribute names are part of the checkpoint contract and follow the reference one to one;
everything else (construction, forward plumbing) is written from scratch.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

LEARNABLE_AFFINE = True  # deeplab/config.py:1


def _bn(channels: int) -> nn.BatchNorm2d:
    return nn.BatchNorm2d(channels, affine=LEARNABLE_AFFINE)


def init_reference_style(root: nn.Module) -> None:
    """Conv weights ~ N(0, 0.01), BatchNorm weight 1 / bias 0 (residual_net.py:116-121 and every other ctor)."""
    for m in root.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data.normal_(0, 0.01)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


def _bilinear(x, size, align_corners=None):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=align_corners)


class Bottleneck(nn.Module):
    """1x1 (strided) -> 3x3 (dilated) -> 1x1 (x4) residual block, stride on the first 1x1
    (residual_net.py:47-98)."""

    expansion = 4

    def __init__(self, in_channels, shrank_channels, stride=1, dilation=1, downsample=None):
        super().__init__()
        widths = [(in_channels, shrank_channels), (shrank_channels, shrank_channels),
                  (shrank_channels, shrank_channels * self.expansion)]
        self.conv1 = nn.Conv2d(*widths[0], kernel_size=1, stride=stride, bias=False)
        self.bn1 = _bn(widths[0][1])
        self.conv2 = nn.Conv2d(*widths[1], kernel_size=3, stride=1, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = _bn(widths[1][1])
        self.conv3 = nn.Conv2d(*widths[2], kernel_size=1, bias=False)
        self.bn3 = _bn(widths[2][1])
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        shortcut = x if self.downsample is None else self.downsample(x)
        y += shortcut
        return self.relu(y)


class BasicBlock(nn.Module):
    """Two 3x3 convolutions (residual_net.py:15-44).  The reference cannot actually build a ResNet from it
    (its ctor has no `dilation`, which `_make_layer` always passes); here `dilation` is accepted and ignored."""

    expansion = 1

    def __init__(self, in_channels, out_channels, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = _bn(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = _bn(out_channels)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        y += x if self.downsample is None else self.downsample(x)
        return self.relu(y)


class ResNet(nn.Module):
    """Partial ResNet, output stride 8: conv1/maxpool/layer2 stride 2, layer3/4 dilated 2/4
    (residual_net.py:101-171)."""

    # (planes, stride, dilation) of layer1..layer4
    STAGES = ((64, 1, 1), (128, 2, 1), (256, 1, 2), (512, 1, 4))

    def __init__(self, input_channels, res_block, num_blocks_of_layers: Sequence[int], num_classes):
        super().__init__()
        self.input_channels = input_channels
        self.inner_channels = 64
        self.conv1 = nn.Conv2d(input_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = _bn(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=True)
        for idx, ((planes, stride, dilation), blocks) in enumerate(zip(self.STAGES, num_blocks_of_layers), start=1):
            setattr(self, f"layer{idx}", self._stage(res_block, planes, blocks, stride, dilation))
        init_reference_style(self)

    def _stage(self, block, planes, blocks, stride, dilation):
        out_channels = planes * block.expansion
        project = stride != 1 or self.inner_channels != out_channels or dilation in (2, 4)
        if not project:
            # the reference dereferences `downsample` unconditionally (residual_net.py:132) and would fail here
            raise ValueError("every stage of this ResNet needs a projection shortcut on its first block")
        shortcut_bn = _bn(out_channels)
        for prm in shortcut_bn.parameters():   # frozen affine of the projection BN (residual_net.py:132-133)
            prm.requires_grad = False
        downsample = nn.Sequential(
            nn.Conv2d(self.inner_channels, out_channels, kernel_size=1, stride=stride, bias=False), shortcut_bn)
        layers: List[nn.Module] = [block(self.inner_channels, planes, stride, dilation=dilation, downsample=downsample)]
        self.inner_channels = out_channels
        layers += [block(out_channels, planes, dilation=dilation) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def get_params(self):
        return [self.conv1, self.bn1, self.layer1, self.layer2, self.layer3, self.layer4]

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        for idx in (1, 2, 3, 4):
            x = getattr(self, f"layer{idx}")(x)
        return x


class ASPP(nn.Module):
    """Image pooling + 1x1 + three dilated 3x3 branches -> concat -> 3x3 bottleneck -> BN -> PReLU
    (deeplabv3_encoder.py:10-86)."""

    def __init__(self, input_channels, output_channels, depth, dilation_series, padding_series):
        super().__init__()
        self.mean = nn.AdaptiveAvgPool2d((1, 1))
        self.conv = nn.Conv2d(input_channels, depth, kernel_size=1, stride=1)
        self.bn_x = nn.BatchNorm2d(depth)
        self.relu = nn.ReLU(inplace=True)
        self.conv2d_0 = nn.Conv2d(input_channels, depth, kernel_size=1, stride=1)
        self.bn_0 = nn.BatchNorm2d(depth)
        for i, (dil, pad) in enumerate(zip(dilation_series, padding_series), start=1):
            setattr(self, f"conv2d_{i}", nn.Conv2d(input_channels, depth, kernel_size=3, stride=1, padding=pad, dilation=dil))
            setattr(self, f"bn_{i}", nn.BatchNorm2d(depth))
        self.bottleneck = nn.Conv2d(depth * 5, output_channels, kernel_size=3, padding=1)
        self.bn = nn.BatchNorm2d(output_channels)
        self.prelu = nn.PReLU()
        init_reference_style(self)

    def pre_tail(self, x):
        """Everything up to and including the 3x3 bottleneck conv; `forward` = prelu(bn(pre_tail(x))).  The drop-in model's
        eval path hands this to the fused tail kernel (BN + PReLU + 16-bit operand cast, cosnet_b200.coattention.encoder_tail)."""
        size = x.shape[2:]
        pooled = self.relu(self.bn_x(self.conv(self.mean(x))))
        branches = [_bilinear(pooled, size, align_corners=True)]
        for i in range(4):
            branches.append(self.relu(getattr(self, f"bn_{i}")(getattr(self, f"conv2d_{i}")(x))))
        return self.bottleneck(torch.cat(branches, 1))

    def forward(self, x):
        return self.prelu(self.bn(self.pre_tail(x)))


class Encoder(nn.Module):
    """RGB encoder: ResNet + ASPP(6, 12, 18) + auxiliary 1x1 classifier (deeplabv3_encoder.py:91-143).
    forward returns (features [N,256,H/8,W/8], full-resolution sigmoid annotation)."""

    def __init__(self, input_channels, res_block, num_blocks_of_layers, num_classes):
        super().__init__()
        self.input_channels = input_channels
        self.backbone = ResNet(input_channels, res_block, num_blocks_of_layers, num_classes)
        self.aspp = ASPP(2048, 256, 512, dilation_series=[6, 12, 18], padding_series=[6, 12, 18])
        self.main_classifier = nn.Conv2d(256, num_classes, kernel_size=1)
        self.softmax = nn.Sigmoid()
        init_reference_style(self)

    def get_params(self, level="none"):
        if level == "backbone":
            bb = self.backbone
            return [bb.conv1, bb.bn1, bb.layer1, bb.layer2, bb.layer3, bb.layer4, self.aspp]
        if level == "classifier":
            return [self.main_classifier]
        return []

    def forward(self, x):
        features = self.aspp(self.backbone(x))
        annotation = self.softmax(_bilinear(self.main_classifier(features), x.shape[2:]))
        return features, annotation


class DepthEncoder_ResNetASPP(nn.Module):
    """Depth encoder: single-channel ResNet + ASPP(2, 3, 7) (deeplabv3_encoder.py:149-185)."""

    def __init__(self, output_channels, res_block, num_blocks_of_layers, num_classes):
        super().__init__()
        self.input_channels = 1
        self.backbone = ResNet(1, res_block, num_blocks_of_layers, num_classes)
        self.aspp = ASPP(2048, output_channels, 512, dilation_series=[2, 3, 7], padding_series=[2, 3, 7])
        init_reference_style(self)

    def get_params(self):
        return self.backbone.get_params() + [self.aspp]

    def forward(self, x):
        return self.aspp(self.backbone(x))
