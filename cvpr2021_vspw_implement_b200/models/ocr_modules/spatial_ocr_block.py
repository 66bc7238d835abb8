"""OCR blocks of TCB-OCR on the vspw_b200 tape engine.

Reference: models/ocr_modules/spatial_ocr_block.py — SpatialTemporalGather_Module (:70-129),
_ObjectAttentionBlock / ObjectAttentionBlock2D (:176-307), SpatialOCR_Module (:310-381).  Only the
configuration the TCB scripts use is built (scale=1, use_gt=False, use_bg=False,
fetch_attention=False); the other switches raise.
"""
import torch.nn as nn

from ... import engine as E
from ..resnet import conv_op
from ..sync_batchnorm import BatchNorm2d


class SpatialTemporalGather_Module(nn.Module):
    """Soft object regions x pixel features, averaged over the T frames of the clip (or over the
    inference memory bank).  Has no parameters."""

    def __init__(self, cls_num=0, scale=1, use_gt=False):
        super().__init__()
        if use_gt:
            raise NotImplementedError("use_gt is never enabled on the TCB path")
        if scale != 1:
            raise NotImplementedError("scale != 1 is never used on the TCB path")
        self.cls_num = cls_num
        self.scale = scale
        self.use_gt = use_gt
        self.relu = nn.ReLU(inplace=True)

    def graph(self, tape, feats, probs, clip_num, memory=None, memory_num=None):
        """feats (N,h,w,512), probs = dsn logits (N,h,w,K), N = (clip_num+1)*n -> context (n,K,1,512).

        memory is None: mean over the clip's frames (reference :97-109) as ONE fused op
        (softmax over hw, K x hw by hw x C contraction per frame, 1/T scale).
        memory is a list: reference :110-125 including its quirk Q9 — once the shared list is non-empty
        the method works on a detached *copy*, so only the very first call of a video grows the bank."""
        t_frames = clip_num + 1
        n = feats.shape[0] // t_frames
        if memory is None:
            return E.region_gather(tape, feats, probs, t_frames, n)
        bank = memory
        if len(bank) > 0:
            bank = [E.Var(m.data) for m in bank]  # detached copy: later appends are lost (Q9)
        for t in range(t_frames):
            f = E.slice_images(tape, feats, t * n, (t + 1) * n)
            p = E.slice_images(tape, probs, t * n, (t + 1) * n)
            ctx = E.region_gather(tape, f, p, 1, n)
            while len(bank) > memory_num:
                bank.pop(0)
            bank.append(ctx)
        return E.mean_over_stack(tape, bank)


def _cbr(cin, cout):
    return [nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0), BatchNorm2d(cout), nn.ReLU(inplace=True)]


def _run_cbr_chain(tape, seq, x, training):
    mods = list(seq)
    for i in range(0, len(mods), 3):
        x = E.batchnorm_act(tape, conv_op(tape, mods[i], x, mods[i + 1]), mods[i + 1], relu=True, training=training)
    return x


class _ObjectAttentionBlock(nn.Module):
    def __init__(self, in_channels, key_channels, scale=1, use_gt=False, use_bg=False, fetch_attention=False):
        super().__init__()
        if scale != 1 or use_gt or use_bg or fetch_attention:
            raise NotImplementedError("only scale=1, use_gt=False, use_bg=False, fetch_attention=False is on the TCB path")
        self.scale = scale
        self.in_channels = in_channels
        self.key_channels = key_channels
        self.use_gt, self.use_bg, self.fetch_attention = use_gt, use_bg, fetch_attention
        self.pool = nn.MaxPool2d(kernel_size=(scale, scale))
        self.f_pixel = nn.Sequential(*(_cbr(in_channels, key_channels) + _cbr(key_channels, key_channels)))
        self.f_object = nn.Sequential(*(_cbr(in_channels, key_channels) + _cbr(key_channels, key_channels)))
        self.f_down = nn.Sequential(*_cbr(in_channels, key_channels))
        self.f_up = nn.Sequential(*_cbr(key_channels, in_channels))

    def graph(self, tape, x, proxy, training):
        query = _run_cbr_chain(tape, self.f_pixel, x, training)      # (n,h,w,kc)
        key = _run_cbr_chain(tape, self.f_object, proxy, training)   # (n,K,1,kc)
        value = _run_cbr_chain(tape, self.f_down, proxy, training)   # (n,K,1,kc)
        ctx = E.object_attention(tape, query, key, value, self.key_channels)
        return _run_cbr_chain(tape, self.f_up, ctx, training)


class ObjectAttentionBlock2D(_ObjectAttentionBlock):
    pass


class SpatialOCR_Module(nn.Module):
    def __init__(self, in_channels, key_channels, out_channels, scale=1, dropout=0.1, use_gt=False, use_bg=False,
                 use_oc=True, fetch_attention=False):
        super().__init__()
        self.use_gt, self.use_bg, self.use_oc, self.fetch_attention = use_gt, use_bg, use_oc, fetch_attention
        self.object_context_block = ObjectAttentionBlock2D(in_channels, key_channels, scale, use_gt, use_bg, fetch_attention)
        self.conv_bn_dropout = nn.Sequential(nn.Conv2d(2 * in_channels, out_channels, kernel_size=1, padding=0),
                                             BatchNorm2d(out_channels), nn.ReLU(inplace=True), nn.Dropout2d(dropout))

    def graph(self, tape, feats, proxy_feats, training):
        context = self.object_context_block.graph(tape, feats, proxy_feats, training)
        cat = E.concat_channels(tape, [context, feats])
        y = conv_op(tape, self.conv_bn_dropout[0], cat, self.conv_bn_dropout[1])
        mask = E.dropout2d_mask(self.conv_bn_dropout[3].p, y.shape[0], y.shape[3], y.data.device, training and self.conv_bn_dropout[3].training)
        return E.batchnorm_act(tape, y, self.conv_bn_dropout[1], relu=True, chan_scale=mask, training=training)
