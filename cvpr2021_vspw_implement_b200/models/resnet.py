"""Deep-stem ResNet-18/50/101 parameter containers + tape graphs.

Reference: models/resnet.py (BasicBlock :24-53, Bottleneck :56-92, ResNet :95-158, factories
:160-205).  Construction order, shapes and init (conv ~ N(0, sqrt(2/(k*k*Cout))), BN gamma=1,
beta=0) follow the reference so the same torch seed yields the same state_dict; the math runs
through ``engine`` (implicit-GEMM convs, fused BN/ReLU/residual kernels).
"""
import math
import os

import torch.nn as nn

from .. import engine as E
from .sync_batchnorm import BatchNorm2d

__all__ = ["ResNet", "resnet18", "resnet50", "resnet101"]


def _conv(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=False)


def conv_op(tape, conv, x, bn=None):
    """Run an nn.Conv2d container through the engine, honouring stride/padding/dilation edits
    made by ResnetDilated._nostride_dilate (reference models/models.py:737-750).  `bn`: the BatchNorm module the
    output feeds; in train mode its statistics are taken in the conv epilogue."""
    if conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1] or conv.dilation[0] != conv.dilation[1]:
        raise NotImplementedError("anisotropic conv geometry is not on the VSPW hot path")
    if conv.groups != 1:
        raise NotImplementedError("grouped convolutions are not on the VSPW hot path")
    return E.conv2d(tape, x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], conv.dilation[0],
                    want_stats=bn is not None and bn.training)


def _planes_only_ok(conv, x_shape):
    """The activation feeding `conv` can skip its fp32 copy when that conv is its only reader and runs on tcgen05."""
    return E.conv_will_use_tc(tuple(x_shape), tuple(conv.weight.shape), conv.stride[0], conv.padding[0], conv.dilation[0])


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride)
        self.bn1 = BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv(planes, planes, 3)
        self.bn2 = BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def graph(self, tape, x, out_fp32=True):
        y1 = conv_op(tape, self.conv1, x, self.bn1)
        out = E.batchnorm_act(tape, y1, self.bn1, relu=True, fp32_out=not _planes_only_ok(self.conv2, y1.shape))
        y2 = conv_op(tape, self.conv2, out, self.bn2)
        res = x
        if self.downsample is not None:
            res = E.batchnorm_act(tape, conv_op(tape, self.downsample[0], x, self.downsample[1]), self.downsample[1], relu=False,
                                  planes_out=False)
        return E.batchnorm_act(tape, y2, self.bn2, relu=True, residual=res, fp32_out=out_fp32)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 1)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, stride)
        self.bn2 = BatchNorm2d(planes)
        self.conv3 = _conv(planes, planes * 4, 1)
        self.bn3 = BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def graph(self, tape, x, out_fp32=True):
        # the outputs of bn1 and bn2 are read by the next conv only: bf16 planes suffice when that conv is a tcgen05 one
        y1 = conv_op(tape, self.conv1, x, self.bn1)
        out = E.batchnorm_act(tape, y1, self.bn1, relu=True, fp32_out=not _planes_only_ok(self.conv2, y1.shape))
        y2 = conv_op(tape, self.conv2, out, self.bn2)
        out = E.batchnorm_act(tape, y2, self.bn2, relu=True, fp32_out=not _planes_only_ok(self.conv3, y2.shape))
        y3 = conv_op(tape, self.conv3, out, self.bn3)
        res = x
        if self.downsample is not None:
            res = E.batchnorm_act(tape, conv_op(tape, self.downsample[0], x, self.downsample[1]), self.downsample[1], relu=False,
                                  planes_out=False)  # read by the residual add only: no operand planes
        # bn3 -> (+residual) -> relu fused in one pass
        return E.batchnorm_act(tape, y3, self.bn3, relu=True, residual=res, fp32_out=out_fp32)


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=146):
        super().__init__()
        self.inplanes = 128
        self.conv1 = _conv(3, 64, 3, stride=2)
        self.bn1 = BatchNorm2d(64)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = _conv(64, 64, 3)
        self.bn2 = BatchNorm2d(64)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv3 = _conv(64, 128, 3)
        self.bn3 = BatchNorm2d(128)
        self.relu3 = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        # kept so that construction consumes the RNG exactly like the reference (resnet.py:115-116)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc_1 = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan))
            elif isinstance(m, BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                BatchNorm2d(planes * block.expansion),
            )
        seq = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        seq += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*seq)


def stem_and_layers_graph(tape, net, x):
    """conv1..conv3 (+BN+ReLU) -> maxpool -> layer1..4; returns the four stage outputs
    (reference ResnetDilated.forward, models/models.py:752-767)."""
    x = E.batchnorm_act(tape, conv_op(tape, net.conv1, x, net.bn1), net.bn1, relu=True)
    x = E.batchnorm_act(tape, conv_op(tape, net.conv2, x, net.bn2), net.bn2, relu=True)
    x = E.batchnorm_act(tape, conv_op(tape, net.conv3, x, net.bn3), net.bn3, relu=True)
    x = E.maxpool3x3s2(tape, x)
    outs = []
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        blocks = list(layer)
        for i, blk in enumerate(blocks):
            # An interior block output is read by the NEXT block only: its conv1 and its residual add.  With
            # VSPW_PLANES_ONLY_OUT=1 it exists as bf16 (hi, lo) operand planes only (the residual add reads the planes:
            # vspw_bn_train_fwd residual_hi/lo): 4 fewer bytes written per element.  Measured inside one box (tools/ab.sh):
            # 92.60 ms/step with it, 91.91 without — two 8-byte plane loads cost the latency-bound BN kernel more than the
            # 16-byte fp32 store saves — so the default keeps the fp32 copy and releases it right after the next block has read it.
            nxt = blocks[i + 1] if i + 1 < len(blocks) else None
            planes_only = nxt is not None and _planes_only_next(blk, nxt, x.shape)
            y = blk.graph(tape, x, out_fp32=not planes_only)
            if i > 0 and x.planes is not None and x.data is not None and _planes_only_ok(blk.conv1, x.shape) and blk.downsample is None:
                x.data = None  # backward reads only the planes: the fp32 copy goes back to the allocator now (12 % of the footprint)
            x = y
        outs.append(x)
    return outs


def _planes_only_next(blk, nxt, in_shape):
    """True when `blk`'s output can exist as operand planes only: the next block of the stage reads it through tcgen05 convs."""
    if E.get_precision() == "fp32" or nxt.downsample is not None or os.environ.get("VSPW_PLANES_ONLY_OUT", "0") != "1":
        return False
    n, h, w, _ = in_shape
    # the stage's first block may stride; ResnetDilated rewrites strides in place, so read the conv that carries it
    s = (blk.conv2 if hasattr(blk, "conv3") else blk.conv1).stride[0]
    ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
    cout = (blk.conv3 if hasattr(blk, "conv3") else blk.conv2).weight.shape[0]
    return cout % 64 == 0 and _planes_only_ok(nxt.conv1, (n, ho, wo, cout))


def resnet18(pretrained=False, **kw):
    if pretrained:
        raise NotImplementedError("no network access: pass `weights=` to ModelBuilder.build_encoder instead")
    return ResNet(BasicBlock, [2, 2, 2, 2], **kw)


def resnet50(pretrained=False, **kw):
    if pretrained:
        raise NotImplementedError("no network access: pass `weights=` to ModelBuilder.build_encoder instead")
    return ResNet(Bottleneck, [3, 4, 6, 3], **kw)


def resnet101(pretrained=False, **kw):
    if pretrained:
        raise NotImplementedError("no network access: pass `weights=` to ModelBuilder.build_encoder instead")
    return ResNet(Bottleneck, [3, 4, 23, 3], **kw)
