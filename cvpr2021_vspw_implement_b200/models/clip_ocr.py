"""TCB-OCR: ClipOCRNet on the vspw_b200 tape engine.

Reference: models/clip_ocr.py:23-198.  Same constructor, parameter names, LR-group generators and
forward contract as the reference (see clip_psp.py in this package for the shared quirks).
``--clipocr_all True`` is broken in the reference itself (batch mismatch T*n vs n, quirk Q11) and
raises here with that explanation.
"""
import torch
import torch.nn as nn

from .. import engine as E
from .models import _ignore_index, _labels_of
from .ocr_modules.spatial_ocr_block import SpatialOCR_Module, SpatialTemporalGather_Module
from .resnet import conv_op
from .sync_batchnorm import BatchNorm2d
from .clip_psp import Clip_PSP


class ClipOCRNet(nn.Module):
    def __init__(self, net_enc, crit, args, deep_sup_scale=None):
        super().__init__()
        self.args = args
        if self.args.use_memory:
            self.memory = []
        self.crit = crit
        self.deep_sup_scale = deep_sup_scale
        self.encoder = net_enc
        self.inplanes = 128
        self.num_classes = args.num_class
        in_channels = [1024, 2048]
        self.conv_3x3 = nn.Sequential(nn.Conv2d(in_channels[1], 512, kernel_size=3, stride=1, padding=1), BatchNorm2d(512),
                                      nn.ReLU(inplace=True))
        self.spatial_context_head = SpatialTemporalGather_Module(self.num_classes)
        self.spatial_ocr_head = SpatialOCR_Module(in_channels=512, key_channels=256, out_channels=512, scale=1, dropout=0.05)
        self.head = nn.Conv2d(512, self.num_classes, kernel_size=1, stride=1, padding=0, bias=True)
        self.dsn_head = nn.Sequential(
            nn.Conv2d(in_channels[0], 512, kernel_size=3, stride=1, padding=1), BatchNorm2d(512), nn.ReLU(inplace=True),
            nn.Dropout2d(0.05), nn.Conv2d(512, self.num_classes, kernel_size=1, stride=1, padding=0, bias=True))

    _walk = staticmethod(Clip_PSP._walk)

    def _decoder_modules(self):
        return [self.conv_3x3, self.spatial_context_head, self.spatial_ocr_head, self.head, self.dsn_head]

    def get_1x_lr_params(self):
        return self._walk([self.encoder], False)

    def get_10x_lr_params(self):
        return self._walk(self._decoder_modules(), False)

    def get_1x_lr_params_bias(self):
        return self._walk([self.encoder], True)

    def get_10x_lr_params_bias(self):
        return self._walk(self._decoder_modules(), True)

    def pixel_acc(self, pred, label):
        _, preds = torch.max(pred, dim=1)
        valid = (label >= 0).long()
        acc_sum = torch.sum(valid * (preds == label).long())
        return acc_sum.float() / (torch.sum(valid).float() + 1e-10)

    def _logits(self, tape, frames, training, memory=None):
        t_frames = len(frames)
        n = frames[0].shape[0]
        x = E.Var(E.input_from_frames(frames))
        maps = self.encoder.graph(tape, x)
        # dsn head on layer3 output of all N frames (reference :117)
        y = conv_op(tape, self.dsn_head[0], maps[-2], self.dsn_head[1])
        mask = E.dropout2d_mask(self.dsn_head[3].p, y.shape[0], y.shape[3], y.data.device, training and self.dsn_head[3].training)
        d = E.batchnorm_act(tape, y, self.dsn_head[1], relu=True, chan_scale=mask, training=training)
        x_dsn = conv_op(tape, self.dsn_head[4], d)
        feats = E.batchnorm_act(tape, conv_op(tape, self.conv_3x3[0], maps[-1], self.conv_3x3[1]), self.conv_3x3[1], relu=True,
                                training=training)
        # (recorded before the gather: the backward then runs the gather's first, which owns the full gradient of `feats`)
        cur = E.slice_images(tape, feats, (t_frames - 1) * n, t_frames * n)
        if memory is not None:
            context = self.spatial_context_head.graph(tape, feats, x_dsn, t_frames - 1, memory, self.args.memory_num)
        else:
            context = self.spatial_context_head.graph(tape, feats, x_dsn, t_frames - 1)
        E.publish("context", context)
        if self.args.clipocr_all:
            raise NotImplementedError("--clipocr_all True fails inside the reference itself (view of T*n pixels rows "
                                      "against an n-clip context, clip_ocr.py:136-137); the TCB scripts use False")
        z = self.spatial_ocr_head.graph(tape, cur, context, training)
        return conv_op(tape, self.head, z), x_dsn

    def forward(self, feed_dict, segSize=None):
        c_img = feed_dict["img_data"]
        clip_imgs = feed_dict["clipimgs_data"]
        clip_imgs.append(c_img)  # reference :112
        frames = list(clip_imgs)
        training = self.training

        if segSize is not None:
            memory = None
            if self.args.use_memory:
                if feed_dict["is_clean_memory"]:
                    self.memory = []
                memory = self.memory

            def runner(tape):
                logits, _ = self._logits(tape, frames, training, memory)
                E.publish("logits", logits)
                return (E.up_softmax(logits, int(segSize[0]), int(segSize[1])),), None

            (pred,) = E.run_graph(self, runner)
            return pred

        label = _labels_of(feed_dict)
        clip_labels = feed_dict["cliplabels_data"]
        clip_labels.append(feed_dict["seg_label"])  # reference :181
        ignore = _ignore_index(self.crit)
        n = label.shape[0]

        def runner(tape):
            logits, x_dsn = self._logits(tape, frames, training)
            E.publish("logits", logits)
            E.publish("logits_deepsup", x_dsn)
            main = E.nll_term(tape, logits, label, ignore, want_acc=True)
            all_lab = torch.empty((n * len(clip_labels), 1) + tuple(label.shape[2:]), device=label.device, dtype=torch.float32)
            for t, lab in enumerate(clip_labels):
                all_lab[t * n:(t + 1) * n].copy_(lab)
            aux = E.nll_term(tape, x_dsn, all_lab, ignore, want_acc=False)
            # the reference multiplies by deep_sup_scale unconditionally (:195): None would raise there too
            loss, acc, gslot = E.loss_combine(tape, main, aux, float(self.deep_sup_scale))
            return (loss, acc), lambda g: gslot.__setitem__("g", g.contiguous())

        loss, acc = E.run_graph(self, runner)
        return loss, acc
