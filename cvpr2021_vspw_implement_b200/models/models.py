"""Plugin surface of the reference for the image path: ModelBuilder, SegmentationModule,
Resnet / ResnetDilated encoders and the PPMDeepsup decoder.

Reference: models/models.py (SegmentationModule :74-111, ModelBuilder :512-656, Resnet :659-704,
ResnetDilated :707-767, PPMDeepsup :938-995).  Same names, constructor arguments, state_dict keys
and error behaviour; forward passes run on the vspw_b200 tape engine (CUDA only).
"""
from functools import partial

import torch
import torch.nn as nn

from .. import engine as E
from . import resnet
from .resnet import conv_op, stem_and_layers_graph
from .sync_batchnorm import BatchNorm2d


def _labels_of(feed_dict, key="seg_label"):
    lab = feed_dict[key]
    if lab.dim() != 4 or lab.shape[1] != 1:
        raise ValueError(f"{key} must have shape (n, 1, H, W), got {tuple(lab.shape)}")
    lab = lab.contiguous()
    return lab if lab.dtype == torch.float32 else lab.float()


class SegmentationModuleBase(nn.Module):
    def pixel_acc(self, pred, label):
        """Reference formula (models.py:65-71), kept for callers that evaluate saved predictions."""
        _, preds = torch.max(pred, dim=1)
        valid = (label >= 0).long()
        acc_sum = torch.sum(valid * (preds == label).long())
        pixel_sum = torch.sum(valid)
        return acc_sum.float() / (pixel_sum.float() + 1e-10)


def _ignore_index(crit):
    return int(getattr(crit, "ignore_index", -100)) if crit is not None else -100


class SegmentationModule(SegmentationModuleBase):
    """encoder + decoder + criterion wrapper (reference models.py:74-111)."""

    def __init__(self, net_enc, net_dec, crit, deep_sup_scale=None):
        super().__init__()
        self.encoder = net_enc
        self.decoder = net_dec
        self.crit = crit
        self.deep_sup_scale = deep_sup_scale

    def forward(self, feed_dict=None, segSize=None):
        if feed_dict is None:
            raise ValueError("feed_dict is required")
        img = feed_dict["img_data"]
        if segSize is None:
            labels = _labels_of(feed_dict)
            ignore = _ignore_index(self.crit)

            def runner(tape):
                x = E.Var(E.input_from_frames([img]))
                maps = self.encoder.graph(tape, x)
                logits, logits_ds = self.decoder.graph(tape, maps, training=self.training,
                                                       want_deepsup=self.deep_sup_scale is not None)
                E.publish("logits", logits)
                main = E.nll_term(tape, logits, labels, ignore, want_acc=True)
                aux = E.nll_term(tape, logits_ds, labels, ignore, want_acc=False) if logits_ds is not None else None
                loss, acc, gslot = E.loss_combine(tape, main, aux, self.deep_sup_scale or 0.0)
                return (loss, acc), lambda g: gslot.__setitem__("g", g.contiguous())

            loss, acc = E.run_graph(self, runner)
            return loss, acc

        def runner(tape):
            x = E.Var(E.input_from_frames([img]))
            maps = self.encoder.graph(tape, x)
            logits, _ = self.decoder.graph(tape, maps, training=False, want_deepsup=False)
            return (E.up_softmax(logits, int(segSize[0]), int(segSize[1])),), None

        with torch.no_grad():
            (pred,) = E.run_graph(self, runner)
        return pred


class _EncoderBase(nn.Module):
    """Shared forward: NCHW in, list of NCHW stage maps out (what external callers of the encoder see)."""

    def graph(self, tape, x):
        return stem_and_layers_graph(tape, self, x)

    def forward(self, x, return_feature_maps=False):
        def runner(tape):
            xin = E.Var(E.input_from_frames([x]), needs_grad=False)
            maps = self.graph(tape, xin)
            sel = maps if return_feature_maps else maps[-1:]
            outs = tuple(E.nhwc_to_nchw(m.data) for m in sel)
            return outs, None

        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # stand-alone differentiable use of the encoder is not part of the hot path: the TCB modules
            # call `graph` directly.  Be explicit instead of silently detaching.
            with torch.no_grad():
                outs = E.run_graph(self, runner)
            return [o.detach() for o in outs]
        outs = E.run_graph(self, runner)
        return list(outs)


def _take_resnet_parts(dst, orig):
    for name in ("conv1", "bn1", "relu1", "conv2", "bn2", "relu2", "conv3", "bn3", "relu3", "maxpool",
                 "layer1", "layer2", "layer3", "layer4"):
        setattr(dst, name, getattr(orig, name))


class Resnet(_EncoderBase):
    def __init__(self, orig_resnet):
        super().__init__()
        _take_resnet_parts(self, orig_resnet)


class ResnetDilated(_EncoderBase):
    """Output-stride-8 (or 16) dilated ResNet: strides of layer3/4 removed, 3x3 convs dilated
    (reference models.py:707-750)."""

    def __init__(self, orig_resnet, dilate_scale=8):
        super().__init__()
        if dilate_scale == 8:
            orig_resnet.layer3.apply(partial(self._nostride_dilate, dilate=2))
            orig_resnet.layer4.apply(partial(self._nostride_dilate, dilate=4))
        elif dilate_scale == 16:
            orig_resnet.layer4.apply(partial(self._nostride_dilate, dilate=2))
        _take_resnet_parts(self, orig_resnet)

    @staticmethod
    def _nostride_dilate(m, dilate):
        if not isinstance(m, nn.Conv2d):
            return
        three = m.kernel_size == (3, 3)
        if m.stride == (2, 2):  # the strided conv of the stage: stride removed, half dilation
            m.stride = (1, 1)
            if three:
                m.dilation = (dilate // 2, dilate // 2)
                m.padding = (dilate // 2, dilate // 2)
        elif three:
            m.dilation = (dilate, dilate)
            m.padding = (dilate, dilate)


class PPMDeepsup(nn.Module):
    """Pyramid pooling decoder with deep supervision (reference models.py:938-995)."""

    def __init__(self, num_class=150, fc_dim=4096, use_softmax=False, pool_scales=(1, 2, 3, 6)):
        super().__init__()
        self.use_softmax = use_softmax
        self.pool_scales = tuple(pool_scales)
        self.ppm = nn.ModuleList([
            nn.Sequential(nn.AdaptiveAvgPool2d(s), nn.Conv2d(fc_dim, 512, kernel_size=1, bias=False), BatchNorm2d(512),
                          nn.ReLU(inplace=True)) for s in pool_scales])
        self.cbr_deepsup = nn.Sequential(nn.Conv2d(fc_dim // 2, fc_dim // 4, kernel_size=3, stride=1, padding=1, bias=False),
                                         BatchNorm2d(fc_dim // 4), nn.ReLU(inplace=True))
        self.conv_last_ = nn.Sequential(
            nn.Conv2d(fc_dim + len(pool_scales) * 512, 512, kernel_size=3, padding=1, bias=False), BatchNorm2d(512),
            nn.ReLU(inplace=True), nn.Dropout2d(0.1), nn.Conv2d(512, num_class, kernel_size=1))
        self.conv_last_deepsup_ = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)
        self.dropout_deepsup = nn.Dropout2d(0.1)

    def graph(self, tape, conv_out, training, want_deepsup):
        conv5 = conv_out[-1]
        n, h, w, c = conv5.shape
        # AdaptiveAvgPool2d(s) of a single image = temporal pooling with T=1
        pooled = E.tcb_pool(tape, conv5, 1, n, self.pool_scales)
        pyr = []
        for branch, p in zip(self.ppm, pooled):
            pyr.append(E.batchnorm_act(tape, conv_op(tape, branch[1], p, branch[2]), branch[2], relu=True, training=training))
        conv = self.conv_last_[0]
        if E.ppm_fused_supported(conv5.shape, [p.shape for p in pyr], conv.weight.shape, conv.padding[0], conv.dilation[0]):
            y = E.ppm_conv_fused(tape, conv5, pyr, conv.weight, conv.padding[0], conv.dilation[0], want_stats=bool(training))
        else:
            cat = E.ppm_concat(tape, conv5, pyr)
            y = conv_op(tape, conv, cat, self.conv_last_[1])
        mask = E.dropout2d_mask(self.conv_last_[3].p, n, y.shape[3], y.data.device, training and self.conv_last_[3].training)
        x = E.batchnorm_act(tape, y, self.conv_last_[1], relu=True, chan_scale=mask, training=training)
        logits = conv_op(tape, self.conv_last_[4], x)
        if not want_deepsup:
            return logits, None
        conv4 = conv_out[-2]
        y = conv_op(tape, self.cbr_deepsup[0], conv4, self.cbr_deepsup[1])
        mask = E.dropout2d_mask(self.dropout_deepsup.p, n, y.shape[3], y.data.device, training and self.dropout_deepsup.training)
        d = E.batchnorm_act(tape, y, self.cbr_deepsup[1], relu=True, chan_scale=mask, training=training)
        return logits, conv_op(tape, self.conv_last_deepsup_, d)


def conv3x3_bn_relu(in_planes, out_planes, stride=1):
    """3x3 convolution + BN + relu (reference models.py:658-666)."""
    return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False),
                         BatchNorm2d(out_planes), nn.ReLU(inplace=True))


def _cbr_graph(tape, seq, x, training, chan_scale=None):
    """conv -> BN -> ReLU of a `conv3x3_bn_relu`-style Sequential (conv at [0], BN at [1])."""
    return E.batchnorm_act(tape, conv_op(tape, seq[0], x, seq[1]), seq[1], relu=True, chan_scale=chan_scale, training=training)


class C1(nn.Module):
    """Single 3x3 conv head (reference models.py:862-886)."""

    def __init__(self, num_class=150, fc_dim=2048, use_softmax=False):
        super().__init__()
        self.use_softmax = use_softmax
        self.cbr = conv3x3_bn_relu(fc_dim, fc_dim // 4, 1)
        self.conv_last_1 = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)

    def graph(self, tape, conv_out, training, want_deepsup):
        if want_deepsup:
            raise ValueError("the 'c1' decoder returns one prediction: build the SegmentationModule with deep_sup_scale=None")
        return conv_op(tape, self.conv_last_1, _cbr_graph(tape, self.cbr, conv_out[-1], training)), None


class C1DeepSup(nn.Module):
    """C1 head with deep supervision on the layer3 map (reference models.py:826-858)."""

    def __init__(self, num_class=150, fc_dim=2048, use_softmax=False):
        super().__init__()
        self.use_softmax = use_softmax
        self.cbr = conv3x3_bn_relu(fc_dim, fc_dim // 4, 1)
        self.cbr_deepsup = conv3x3_bn_relu(fc_dim // 2, fc_dim // 4, 1)
        self.conv_last_ = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)
        self.conv_last_deepsup_ = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)

    def graph(self, tape, conv_out, training, want_deepsup):
        logits = conv_op(tape, self.conv_last_, _cbr_graph(tape, self.cbr, conv_out[-1], training))
        if not want_deepsup:
            return logits, None
        return logits, conv_op(tape, self.conv_last_deepsup_, _cbr_graph(tape, self.cbr_deepsup, conv_out[-2], training))


def _ppm_head(tape, branches, conv_idx, conv5, pool_scales, conv_last, bn_last, training):
    """Pyramid pooling -> per-scale 1x1 conv + BN + ReLU -> (up-sample, concat, 3x3 conv) as ONE fused op when the tensor-core
    path can take it (engine.ppm_conv_fused: no up-sampled maps, no concat), else the explicit concat."""
    n = conv5.shape[0]
    pooled = E.tcb_pool(tape, conv5, 1, n, pool_scales)  # AdaptiveAvgPool2d(s) of single images = temporal pooling with T=1
    pyr = [E.batchnorm_act(tape, conv_op(tape, br[conv_idx], p, br[conv_idx + 1]), br[conv_idx + 1], relu=True, training=training)
           for br, p in zip(branches, pooled)]
    if E.ppm_fused_supported(conv5.shape, [p.shape for p in pyr], conv_last.weight.shape, conv_last.padding[0], conv_last.dilation[0]):
        return E.ppm_conv_fused(tape, conv5, pyr, conv_last.weight, conv_last.padding[0], conv_last.dilation[0], want_stats=bool(training))
    return conv_op(tape, conv_last, E.ppm_concat(tape, conv5, pyr), bn_last)


class PPM(nn.Module):
    """Pyramid pooling decoder without deep supervision (reference models.py:889-935)."""

    def __init__(self, num_class=150, fc_dim=4096, use_softmax=False, pool_scales=(1, 2, 3, 6)):
        super().__init__()
        self.use_softmax = use_softmax
        self.pool_scales = tuple(pool_scales)
        self.ppm = nn.ModuleList([
            nn.Sequential(nn.AdaptiveAvgPool2d(s), nn.Conv2d(fc_dim, 512, kernel_size=1, bias=False), BatchNorm2d(512),
                          nn.ReLU(inplace=True)) for s in pool_scales])
        self.conv_last = nn.Sequential(
            nn.Conv2d(fc_dim + len(pool_scales) * 512, 512, kernel_size=3, padding=1, bias=False), BatchNorm2d(512),
            nn.ReLU(inplace=True), nn.Dropout2d(0.1), nn.Conv2d(512, num_class, kernel_size=1))

    def graph(self, tape, conv_out, training, want_deepsup):
        if want_deepsup:
            raise ValueError("the 'ppm' decoder returns one prediction: build the SegmentationModule with deep_sup_scale=None")
        conv5 = conv_out[-1]
        y = _ppm_head(tape, self.ppm, 1, conv5, self.pool_scales, self.conv_last[0], self.conv_last[1], training)
        mask = E.dropout2d_mask(self.conv_last[3].p, y.shape[0], y.shape[3], y.data.device, training and self.conv_last[3].training)
        x = E.batchnorm_act(tape, y, self.conv_last[1], relu=True, chan_scale=mask, training=training)
        return conv_op(tape, self.conv_last[4], x), None


class UPerNet(nn.Module):
    """PPM on the top stage + FPN over the four stages (reference models.py:1085-1175)."""

    def __init__(self, num_class=150, fc_dim=4096, use_softmax=False, pool_scales=(1, 2, 3, 6),
                 fpn_inplanes=(256, 512, 1024, 2048), fpn_dim=256):
        super().__init__()
        self.use_softmax = use_softmax
        self.pool_scales = tuple(pool_scales)
        self.ppm_pooling = nn.ModuleList([nn.AdaptiveAvgPool2d(s) for s in pool_scales])
        self.ppm_conv = nn.ModuleList([
            nn.Sequential(nn.Conv2d(fc_dim, 512, kernel_size=1, bias=False), BatchNorm2d(512), nn.ReLU(inplace=True))
            for _ in pool_scales])
        self.ppm_last_conv = conv3x3_bn_relu(fc_dim + len(pool_scales) * 512, fpn_dim, 1)
        self.fpn_in = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, fpn_dim, kernel_size=1, bias=False), BatchNorm2d(fpn_dim), nn.ReLU(inplace=True))
            for c in fpn_inplanes[:-1]])
        self.fpn_out = nn.ModuleList([nn.Sequential(conv3x3_bn_relu(fpn_dim, fpn_dim, 1)) for _ in range(len(fpn_inplanes) - 1)])
        self.conv_last_ = nn.Sequential(conv3x3_bn_relu(len(fpn_inplanes) * fpn_dim, fpn_dim, 1),
                                        nn.Conv2d(fpn_dim, num_class, kernel_size=1))

    def graph(self, tape, conv_out, training, want_deepsup):
        if want_deepsup:
            raise ValueError("the 'upernet' decoders return one prediction: build the SegmentationModule with deep_sup_scale=None")
        conv5 = conv_out[-1]
        n, h, w, _ = conv5.shape
        # NB UPerNet applies the 1x1 conv + BN AFTER the up-sampling (:1131-1134), so its BN statistics are taken over the h x w
        # map: the branches cannot be evaluated in bin space, the up-sampled maps are materialised as in the reference
        pooled = E.tcb_pool(tape, conv5, 1, n, self.pool_scales)
        parts = [conv5] + [_cbr_graph(tape, br, E.upsample_bilinear(tape, p, h, w), training) for br, p in zip(self.ppm_conv, pooled)]
        f = _cbr_graph(tape, self.ppm_last_conv, E.concat_channels(tape, parts), training)
        feats = [f]
        for i in reversed(range(len(conv_out) - 1)):
            lat = _cbr_graph(tape, self.fpn_in[i], conv_out[i], training)
            f = E.add_vars(tape, lat, E.upsample_bilinear(tape, f, lat.shape[1], lat.shape[2]))
            feats.append(_cbr_graph(tape, self.fpn_out[i][0], f, training))
        feats.reverse()
        H, W = feats[0].shape[1], feats[0].shape[2]
        fusion = E.concat_channels(tape, [feats[0]] + [E.upsample_bilinear(tape, t, H, W) for t in feats[1:]])
        x = _cbr_graph(tape, self.conv_last_[0], fusion, training)
        return conv_op(tape, self.conv_last_[1], x), None


_OUT_OF_SCOPE_DECODERS = ("ppm_deepsup_clip", "ppm_clip", "deeplab", "nonlocal2d")
_OUT_OF_SCOPE_ENCODERS = ("mobilenetv2dilated", "resnext101", "hrnetv2", "hrnetv2_clip", "hrnetv2_clip2")


class ModelBuilder:
    """Factory with the reference's static-method API (models.py:512-656)."""

    @staticmethod
    def weights_init(m):
        classname = m.__class__.__name__
        if classname.find("Conv") != -1:
            nn.init.kaiming_normal_(m.weight.data)
        elif classname.find("BatchNorm") != -1:
            m.weight.data.fill_(1.0)
            m.bias.data.fill_(1e-4)

    @staticmethod
    def build_encoder(arch="resnet50dilated", fc_dim=512, weights="", args=None):
        arch = arch.lower()
        table = {
            "resnet18": ("resnet18", False), "resnet18dilated": ("resnet18", True),
            "resnet50": ("resnet50", False), "resnet50dilated": ("resnet50", True),
            "resnet101": ("resnet101", False), "resnet101dilated": ("resnet101", True),
        }
        if arch in ("resnet34", "resnet34dilated"):
            raise NotImplementedError  # as the reference (models.py:539-546)
        if arch in _OUT_OF_SCOPE_ENCODERS:
            raise NotImplementedError(f"encoder '{arch}' is outside the VSPW TCB hot path this engine implements")
        if arch not in table:
            raise Exception("Architecture undefined!")
        ctor, dilated = table[arch]
        orig = resnet.__dict__[ctor](pretrained=False)
        net_encoder = ResnetDilated(orig, dilate_scale=8) if dilated else Resnet(orig)
        if len(weights) > 0:
            print("Loading weights for net_encoder")
            net_encoder.load_state_dict(torch.load(weights, map_location=lambda storage, loc: storage), strict=False)
        return net_encoder

    @staticmethod
    def build_decoder(arch="ppm_deepsup", fc_dim=512, num_class=150, weights="", use_softmax=False):
        arch = arch.lower()
        if arch == "ppm_deepsup":
            net_decoder = PPMDeepsup(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == "c1_deepsup":
            net_decoder = C1DeepSup(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == "c1":
            net_decoder = C1(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == "ppm":
            net_decoder = PPM(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == "upernet_lite":
            net_decoder = UPerNet(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax, fpn_dim=256)
        elif arch == "upernet":
            net_decoder = UPerNet(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax, fpn_dim=512)
        elif arch == "ocrnet_deepsup":
            from .ocrnet import SpatialOCRNet  # image-level OCR head (reference models/ocrnet.py:22-72)
            net_decoder = SpatialOCRNet(num_class=num_class)
        elif arch in _OUT_OF_SCOPE_DECODERS:
            raise NotImplementedError(f"decoder '{arch}' is outside the VSPW TCB hot path this engine implements")
        else:
            raise Exception("Architecture undefined!")
        net_decoder.apply(ModelBuilder.weights_init)
        if len(weights) > 0:
            print("Loading weights for net_decoder")
            net_decoder.load_state_dict(torch.load(weights, map_location=lambda storage, loc: storage), strict=False)
        return net_decoder
