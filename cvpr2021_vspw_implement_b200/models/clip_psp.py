"""TCB-PSP: Clip_PSP and PPM_conv on the vspw_b200 tape engine.

Reference: models/clip_psp.py (PPM_conv :23-56, Clip_PSP :63-217).  Same constructor signature,
parameter names, LR-group generators and forward(feed_dict, segSize=None) contract:
train -> (loss, acc) 0-d tensors attached to autograd, eval -> (n, num_class, H, W) probabilities.
Reference quirks that are kept on purpose: the current frame is the LAST
image of the batch (Q1); log_softmax is taken before the bilinear up-sampling (Q2); with
``psp_weight`` the softmax-over-T weights are multiplied in and the result is still divided by T,
and weight j multiplies list position j where position 0 is the current frame (Q3); pixel_acc
counts label 255 as valid (Q6); the LR-group generators yield duplicates (Q10); the forward appends
to the caller's ``clipimgs_data`` / ``cliplabels_data`` lists.
"""
import torch
import torch.nn as nn

from .. import engine as E
from .models import _ignore_index, _labels_of
from .resnet import conv_op
from .sync_batchnorm import BatchNorm2d


class PPM_conv(nn.Module):
    def __init__(self, fc_dim=2048, num_class=None, pool_scales=(1, 2, 3, 6)):
        super().__init__()
        self.ppm = nn.ModuleList([
            nn.Sequential(nn.Conv2d(fc_dim, 512, kernel_size=1, bias=False), BatchNorm2d(512), nn.ReLU(inplace=True))
            for _ in pool_scales])
        self.conv_last_ = nn.Sequential(
            nn.Conv2d(fc_dim + len(pool_scales) * 512, 512, kernel_size=3, padding=1, bias=False), BatchNorm2d(512),
            nn.ReLU(inplace=True), nn.Dropout2d(0.1), nn.Conv2d(512, num_class, kernel_size=1))

    def graph(self, tape, x, pooled, training):
        """x: current-frame layer4 map (n,h,w,2048); pooled: per-scale temporal means (n,s,s,2048)."""
        pyr = [E.batchnorm_act(tape, conv_op(tape, br[0], p, br[1]), br[1], relu=True, training=training)
               for br, p in zip(self.ppm, pooled)]
        conv = self.conv_last_[0]
        want_stats = self.conv_last_[1].training if training is None else bool(training)
        if E.ppm_fused_supported(x.shape, [p.shape for p in pyr], conv.weight.shape, conv.padding[0], conv.dilation[0]):
            # no up-sampled maps, no 4096-channel concat: the pyramid branches are evaluated in bin space (csrc/ppm.cu)
            y = E.ppm_conv_fused(tape, x, pyr, conv.weight, conv.padding[0], conv.dilation[0], want_stats=want_stats)
        else:
            cat = E.ppm_concat(tape, x, pyr)
            y = conv_op(tape, conv, cat, self.conv_last_[1])
        mask = E.dropout2d_mask(self.conv_last_[3].p, y.shape[0], y.shape[3], y.data.device, training and self.conv_last_[3].training)
        z = E.batchnorm_act(tape, y, self.conv_last_[1], relu=True, chan_scale=mask, training=training)
        return conv_op(tape, self.conv_last_[4], z)


class Clip_PSP(nn.Module):
    def __init__(self, net_enc, crit, args, pool_scales=(1, 2, 3, 6), deep_sup_scale=None):
        super().__init__()
        self.encoder = net_enc
        self.crit = crit
        self.deep_sup_scale = deep_sup_scale
        self.args = args
        fc_dim = 2048
        self.pool_scales = tuple(pool_scales)
        self.ppm_conv = PPM_conv(fc_dim, args.num_class, pool_scales=pool_scales)
        self.deepsup = nn.Sequential(
            nn.Conv2d(fc_dim // 2, fc_dim // 4, kernel_size=3, stride=1, padding=1, bias=False), BatchNorm2d(fc_dim // 4),
            nn.ReLU(inplace=True), nn.Dropout2d(0.1), nn.Conv2d(fc_dim // 4, args.num_class, 1, 1, 0))
        if self.args.psp_weight:
            self.pspweight_conv = nn.Sequential(nn.Conv2d(fc_dim, 1, kernel_size=1, bias=False), nn.AdaptiveAvgPool2d((1, 1)))
        self.ppm_pool = nn.ModuleList([nn.AdaptiveAvgPool2d(s) for s in pool_scales])

    # -- optimizer hooks (reference :99-135; duplicates are part of the contract, quirk Q10) --------
    @staticmethod
    def _walk(modules, want_bias):
        for mod in modules:
            for _, sub in mod.named_modules():
                for key, p in sub.named_parameters():
                    if p.requires_grad and (("bias" in key) == want_bias):
                        yield p

    def get_1x_lr_params(self):
        return self._walk([self.encoder], False)

    def get_10x_lr_params(self):
        mods = [self.ppm_conv]
        if self.deep_sup_scale is not None:
            mods.append(self.deepsup)
        if self.args.psp_weight:
            mods.append(self.pspweight_conv)
        return self._walk(mods, False)

    def get_1x_lr_params_bias(self):
        return self._walk([self.encoder], True)

    def get_10x_lr_params_bias(self):
        mods = [self.ppm_conv]
        if self.deep_sup_scale is not None:
            mods.append(self.deepsup)
        return self._walk(mods, True)

    def pixel_acc(self, pred, label):
        _, preds = torch.max(pred, dim=1)
        valid = (label >= 0).long()
        acc_sum = torch.sum(valid * (preds == label).long())
        return acc_sum.float() / (torch.sum(valid).float() + 1e-10)

    # -- graph ---------------------------------------------------------------------------------------
    def _frame_weights(self, tape, feat, t_frames, n_clips):
        """psp_weight branch (reference :147-152,184-187): 1x1 conv 2048->1, global average, softmax over
        the T frames; returned as a (T, n) Var laid out for vspw_tcb_pool (frame t uses list slot (t+1)%T)."""
        score = conv_op(tape, self.pspweight_conv[0], feat)  # 1x1 2048 -> 1, bias-free (reference :83)
        return E.frame_weights(tape, score, t_frames, n_clips)

    def _logits(self, tape, frames, training):
        t_frames = len(frames)
        n = frames[0].shape[0]
        x = E.Var(E.input_from_frames(frames))
        maps = self.encoder.graph(tape, x)
        feat = maps[-1]
        fw = self._frame_weights(tape, feat, t_frames, n) if self.args.psp_weight else None
        # (the slice is recorded BEFORE the pooling so that the backward runs the pooling's first: it then owns the full-size
        # gradient of `feat` and the current frame's rows are added into it, instead of a 526 MB zero fill + a full add)
        cur = E.slice_images(tape, feat, (t_frames - 1) * n, t_frames * n)
        pooled = E.tcb_pool(tape, feat, t_frames, n, self.pool_scales, frame_w=fw)
        return self.ppm_conv.graph(tape, cur, pooled, training), maps

    def forward(self, feed_dict, segSize=None):
        c_img = feed_dict["img_data"]
        clip_imgs = feed_dict["clipimgs_data"]
        clip_imgs.append(c_img)  # the reference mutates the caller's list (:142)
        frames = list(clip_imgs)
        training = self.training

        if segSize is not None:
            def runner(tape):
                logits, _ = self._logits(tape, frames, training=False if not training else training)
                E.publish("logits", logits)
                return (E.up_softmax(logits, int(segSize[0]), int(segSize[1])),), None

            (pred,) = E.run_graph(self, runner)
            return pred

        label = _labels_of(feed_dict)
        clip_labels = feed_dict["cliplabels_data"]
        clip_labels.append(feed_dict["seg_label"])  # reference :197
        ignore = _ignore_index(self.crit)
        n = label.shape[0]

        def runner(tape):
            logits, maps = self._logits(tape, frames, training)
            E.publish("logits", logits)
            main = E.nll_term(tape, logits, label, ignore, want_acc=True)
            aux = None
            if self.deep_sup_scale is not None:
                all_lab = torch.empty((n * len(clip_labels), 1) + tuple(label.shape[2:]), device=label.device,
                                      dtype=torch.float32)
                for t, lab in enumerate(clip_labels):
                    all_lab[t * n:(t + 1) * n].copy_(lab)  # D2D copy (torch.cat in the reference, :204)
                conv4 = maps[-2]
                y = conv_op(tape, self.deepsup[0], conv4, self.deepsup[1])
                mask = E.dropout2d_mask(self.deepsup[3].p, y.shape[0], y.shape[3], y.data.device, training and self.deepsup[3].training)
                d = E.batchnorm_act(tape, y, self.deepsup[1], relu=True, chan_scale=mask, training=training)
                lds = conv_op(tape, self.deepsup[4], d)
                E.publish("logits_deepsup", lds)
                aux = E.nll_term(tape, lds, all_lab, ignore, want_acc=False)
            loss, acc, gslot = E.loss_combine(tape, main, aux, self.deep_sup_scale or 0.0)
            return (loss, acc), lambda g: gslot.__setitem__("g", g.contiguous())

        loss, acc = E.run_graph(self, runner)
        return loss, acc
