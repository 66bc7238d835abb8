"""SpatialOCRNet — the image-level OCR decoder (`--arch_decoder ocrnet_deepsup`) on the vspw_b200 tape engine.

Reference: models/ocrnet.py:22-72 and SpatialGather_Module (models/ocr_modules/spatial_ocr_block.py:39-68).  Same constructor,
parameter names and outputs as the reference decoder; it shares every kernel with ClipOCRNet: the region gather is the
temporal gather with T = 1 (tcgen05 weight-gradient kernel), the pixel->region attention is the fused tcgen05 kernel of
csrc/ocr_tc.cu.  Used through `SegmentationModule(net_enc, net_dec, crit, deep_sup_scale)`.
"""
import torch.nn as nn

from .. import engine as E
from .ocr_modules.spatial_ocr_block import SpatialOCR_Module
from .resnet import conv_op
from .sync_batchnorm import BatchNorm2d


class SpatialGather_Module(nn.Module):
    """Soft object regions x pixel features of ONE image (reference spatial_ocr_block.py:39-68); no parameters."""

    def __init__(self, cls_num=0, scale=1, use_gt=False):
        super().__init__()
        if use_gt or scale != 1:
            raise NotImplementedError("use_gt / scale != 1 are never enabled by the reference's builders")
        self.cls_num, self.scale, self.use_gt = cls_num, scale, use_gt
        self.relu = nn.ReLU(inplace=True)

    def graph(self, tape, feats, probs):
        return E.region_gather(tape, feats, probs, 1, feats.shape[0])


class SpatialOCRNet(nn.Module):
    def __init__(self, num_class):
        super().__init__()
        self.inplanes = 128
        self.num_classes = num_class
        in_channels = [1024, 2048]
        self.conv_3x3 = nn.Sequential(nn.Conv2d(in_channels[1], 512, kernel_size=3, stride=1, padding=1), BatchNorm2d(512),
                                      nn.ReLU(inplace=True))
        self.spatial_context_head = SpatialGather_Module(self.num_classes)
        self.spatial_ocr_head = SpatialOCR_Module(in_channels=512, key_channels=256, out_channels=512, scale=1, dropout=0.05)
        self.head = nn.Conv2d(512, self.num_classes, kernel_size=1, stride=1, padding=0, bias=True)
        self.dsn_head = nn.Sequential(nn.Conv2d(in_channels[0], 512, kernel_size=3, stride=1, padding=1), BatchNorm2d(512),
                                      nn.ReLU(inplace=True), nn.Dropout2d(0.05),
                                      nn.Conv2d(512, self.num_classes, kernel_size=1, stride=1, padding=0, bias=True))

    def graph(self, tape, conv_out, training, want_deepsup):
        y = conv_op(tape, self.dsn_head[0], conv_out[-2], self.dsn_head[1])
        mask = E.dropout2d_mask(self.dsn_head[3].p, y.shape[0], y.shape[3], y.data.device, training and self.dsn_head[3].training)
        d = E.batchnorm_act(tape, y, self.dsn_head[1], relu=True, chan_scale=mask, training=training)
        x_dsn = conv_op(tape, self.dsn_head[4], d)
        feats = E.batchnorm_act(tape, conv_op(tape, self.conv_3x3[0], conv_out[-1], self.conv_3x3[1]), self.conv_3x3[1], relu=True,
                                training=training)
        context = self.spatial_context_head.graph(tape, feats, x_dsn)
        z = self.spatial_ocr_head.graph(tape, feats, context, training)
        logits = conv_op(tape, self.head, z)
        # the reference's forward always returns (x, x_dsn) in training: SegmentationModule needs deep_sup_scale for it
        return logits, (x_dsn if want_deepsup else None)
