"""CPU oracle for the VSPW per-clip hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / CPU baseline.  The product package
``cvpr2021_vspw_implement_b200`` never imports it.

What it is: a functional restatement, in plain PyTorch fp32 CPU ops, of the reference's algorithm for
the path (state_dict in, tensors out; no nn.Module of the reference is needed at run time).  The
reference's arithmetic lives in a third-party dependency that is not vendored under /root/reference:
PyTorch/ATen ("Pytorch 1.3.1" in README.md:11-13, no lockfile; this image has torch 2.11.0+cu128).
The call sites restated here, by reference file:line:

  * ResNet / ResnetDilated graph       models/resnet.py:24-158, models/models.py:707-767
  * BN (single device / eval)          models/sync_batchnorm/batchnorm.py:68-73  (F.batch_norm)
  * Clip_PSP / PPM_conv                models/clip_psp.py:23-56, :136-217
  * ClipOCRNet                         models/clip_ocr.py:106-198
  * gather / attention / OCR module    models/ocr_modules/spatial_ocr_block.py:82-129, :247-289, :358-381
  * SegmentationModule + PPMDeepsup    models/models.py:74-111, :938-995
  * Evaluator (mIoU ...)               utils.py:55-107

Pinning: the reference ships no golden vectors for this path, so the pins are
outputs of the reference modules themselves, executed in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/``; ``tests/test_oracle.py`` checks this
restatement against every one of them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------------------
# primitive ops (each is the ATen call the reference makes)

RELU = True  # False: every ReLU of the restated graphs is the identity (the *_norelu fixtures: ReLU-free gradient parity)


def _act(x):
    return F.relu(x) if RELU else x

def conv2d(x, w, b=None, stride=1, pad=0, dil=1):
    return F.conv2d(x, w, b, stride=stride, padding=pad, dilation=dil)


def batch_norm(sd, prefix, x, train):
    """F.batch_norm(input, running_mean, running_var, weight, bias, training, 0.1, 1e-5)
    (sync_batchnorm/batchnorm.py:71-73).  Running stats in ``sd`` are updated in place when train."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], train, BN_MOMENTUM, BN_EPS)


def maxpool(x):
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)  # resnet.py:109


def bilinear(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------------------------
def _layer_blocks(sd, prefix, layer):
    idx = set()
    for k in sd:
        if k.startswith(f"{prefix}{layer}."):
            idx.add(int(k[len(prefix) + len(layer) + 1:].split(".")[0]))
    return sorted(idx)


def resnet_geometry(layer_i, block_i, dilated, bottleneck):
    """(stride, pad, dil) of the strided 3x3 conv, of the other 3x3 convs and the downsample stride for
    block `block_i` of layer{layer_i}; restates _make_layer (resnet.py:126-141) + _nostride_dilate
    (models.py:737-750) with dilate_scale=8: layer3 -> dilate 2, layer4 -> dilate 4."""
    stage_stride = 1 if layer_i == 1 else 2
    dilate = {3: 2, 4: 4}.get(layer_i, 1) if dilated else 1
    first = block_i == 0
    if dilate > 1:
        strided = (1, dilate // 2, dilate // 2) if first else (1, dilate, dilate)
        other = (1, dilate, dilate)
        ds_stride = 1
    else:
        strided = (stage_stride if first else 1, 1, 1)
        other = (1, 1, 1)
        ds_stride = stage_stride
    return strided, other, ds_stride


def resnet_forward(sd, prefix, x, train, dilated=True):
    """ResnetDilated.forward(x, return_feature_maps=True) (models.py:752-767)."""
    p = prefix
    x = _act(batch_norm(sd, p + "bn1", conv2d(x, sd[p + "conv1.weight"], None, 2, 1, 1), train))
    x = _act(batch_norm(sd, p + "bn2", conv2d(x, sd[p + "conv2.weight"], None, 1, 1, 1), train))
    x = _act(batch_norm(sd, p + "bn3", conv2d(x, sd[p + "conv3.weight"], None, 1, 1, 1), train))
    x = maxpool(x)
    outs = []
    for li in (1, 2, 3, 4):
        layer = f"layer{li}"
        for bi in _layer_blocks(sd, p, layer):
            q = f"{p}{layer}.{bi}."
            bottleneck = (q + "conv3.weight") in sd
            strided, other, ds_stride = resnet_geometry(li, bi, dilated, bottleneck)
            res = x
            if bottleneck:  # resnet.py:72-92 (stride lives on the 3x3 conv2)
                o = _act(batch_norm(sd, q + "bn1", conv2d(x, sd[q + "conv1.weight"]), train))
                o = _act(batch_norm(sd, q + "bn2", conv2d(o, sd[q + "conv2.weight"], None, *strided), train))
                o = batch_norm(sd, q + "bn3", conv2d(o, sd[q + "conv3.weight"]), train)
            else:  # BasicBlock resnet.py:37-53 (stride lives on conv1)
                o = _act(batch_norm(sd, q + "bn1", conv2d(x, sd[q + "conv1.weight"], None, *strided), train))
                o = batch_norm(sd, q + "bn2", conv2d(o, sd[q + "conv2.weight"], None, *other), train)
            if (q + "downsample.0.weight") in sd:
                res = batch_norm(sd, q + "downsample.1", conv2d(x, sd[q + "downsample.0.weight"], None, ds_stride), train)
            x = _act(o + res)
        outs.append(x)
    return outs


# --------------------------------------------------------------------------------------------------
def nll_up(logits, labels, ignore_index):
    """log_softmax at h x w, bilinear to label size, NLLLoss(ignore_index), and the up-sampled log-probs
    (clip_psp.py:198-203; quirk Q2: log_softmax BEFORE interpolation)."""
    lab = labels.squeeze(1).long()
    lp = bilinear(F.log_softmax(logits, dim=1), lab.shape[-2:])
    return F.nll_loss(lp, lab, ignore_index=ignore_index), lp, lab


def pixel_acc(pred, label):
    """clip_psp.py:92-98 (label 255 counts as valid, quirk Q6)."""
    preds = torch.max(pred, dim=1)[1]
    valid = (label >= 0).long()
    return torch.sum(valid * (preds == label).long()).float() / (torch.sum(valid).float() + 1e-10)


def tcb_pool(feat_frames, scales, frame_weights=None):
    """Temporal pyramid pooling (clip_psp.py:157-188): list position 0 is the CURRENT frame (last of the
    batch), positions 1.. are the other frames in order; mean over positions (after the optional
    psp_weight product, which is indexed by list position — quirks Q1/Q3)."""
    cur, others = feat_frames[-1], feat_frames[:-1]
    outs = []
    for s in scales:
        stack = [F.adaptive_avg_pool2d(cur, s).unsqueeze(-1)] + [F.adaptive_avg_pool2d(o, s).unsqueeze(-1) for o in others]
        feature = torch.cat(stack, dim=-1)
        if frame_weights is not None:
            feature = feature * frame_weights
        outs.append(torch.mean(feature, dim=-1))
    return outs


def clip_psp_forward(sd, frames, labels=None, args_psp_weight=False, deep_sup_scale=0.4, train=True, seg_size=None,
                     ignore_index=255, pool_scales=(1, 2, 3, 6), dropout_masks=None):
    """Clip_PSP.forward (clip_psp.py:136-217).  `frames`: list of T (n,3,H,W) tensors, current frame LAST;
    `labels`: list of T (n,1,H,W) float tensors in the same order (train only).  Dropout2d is the identity
    unless per-(n,c) scale masks are injected via dropout_masks={'ppm': m, 'deepsup': m}.
    Returns dict(loss, acc, logits, logits_deepsup) or dict(probs, logits)."""
    T = len(frames)
    n = frames[0].shape[0]
    maps = resnet_forward(sd, "encoder.", torch.cat(frames, dim=0), train)
    out_tmp = maps[-1]
    fw = None
    if args_psp_weight:
        pw = F.adaptive_avg_pool2d(conv2d(out_tmp, sd["pspweight_conv.0.weight"]), (1, 1))
        pw = torch.cat([c.unsqueeze(-1) for c in torch.split(pw, n, dim=0)], dim=-1)
        fw = F.softmax(pw, dim=-1)
    per_frame = list(torch.split(out_tmp, n, dim=0))
    c_tmp = per_frame[-1]
    p_fs = tcb_pool(per_frame, pool_scales, fw)
    # PPM_conv.forward (clip_psp.py:45-56)
    ppm_out = [c_tmp]
    for i, pf in enumerate(p_fs):
        z = _act(batch_norm(sd, f"ppm_conv.ppm.{i}.1", conv2d(pf, sd[f"ppm_conv.ppm.{i}.0.weight"]), train))
        ppm_out.append(bilinear(z, c_tmp.shape[-2:]))
    cat = torch.cat(ppm_out, 1)
    z = _act(batch_norm(sd, "ppm_conv.conv_last_.1", conv2d(cat, sd["ppm_conv.conv_last_.0.weight"], None, 1, 1, 1), train))
    if dropout_masks and dropout_masks.get("ppm") is not None:
        z = z * dropout_masks["ppm"][:, :, None, None]
    logits = conv2d(z, sd["ppm_conv.conv_last_.4.weight"], sd["ppm_conv.conv_last_.4.bias"])
    if seg_size is not None:
        return {"logits": logits, "probs": F.softmax(bilinear(logits, seg_size), dim=1)}
    loss, lp, lab = nll_up(logits, labels[-1], ignore_index)
    out = {"logits": logits, "loss_main": loss}
    if deep_sup_scale is not None:
        d = _act(batch_norm(sd, "deepsup.1", conv2d(maps[-2], sd["deepsup.0.weight"], None, 1, 1, 1), train))
        if dropout_masks and dropout_masks.get("deepsup") is not None:
            d = d * dropout_masks["deepsup"][:, :, None, None]
        lds = conv2d(d, sd["deepsup.4.weight"], sd["deepsup.4.bias"])
        loss_ds, _, _ = nll_up(lds, torch.cat(labels, dim=0), ignore_index)
        out["logits_deepsup"] = lds
        loss = loss + loss_ds * deep_sup_scale
    out["loss"] = loss
    out["acc"] = pixel_acc(lp, lab)
    return out


# --------------------------------------------------------------------------------------------------
def non_local3d_forward(sd, frames, labels=None, train=True, seg_size=None, ignore_index=255):
    """Non_local3d.forward (non_local_models.py:19-71) with NLBlockND(mode='dot', dimension=3) (non_local.py:82-151),
    written with the reference's explicit (T h w) x (T h w) affinity: f = theta^T phi, f / P, y = f g.
    `frames` / `labels`: lists of T tensors, every frame supervised.  Returns dict(loss, acc, logits) or dict(probs list)."""
    T, n = len(frames), frames[0].shape[0]
    x = resnet_forward(sd, "encoder.", torch.cat(frames, dim=0), train)[-1]
    emb = conv2d(x, sd["emb.weight"], sd["emb.bias"])
    v = torch.cat([e.unsqueeze(2) for e in torch.split(emb, n, dim=0)], 2)  # (n, 256, T, h, w)
    q = "nonlocalblock."
    c = sd[q + "g.weight"].shape[0]
    g_x = F.conv3d(v, sd[q + "g.weight"], sd[q + "g.bias"]).view(n, c, -1).permute(0, 2, 1)
    theta_x = F.conv3d(v, sd[q + "theta.weight"], sd[q + "theta.bias"]).view(n, c, -1).permute(0, 2, 1)
    phi_x = F.conv3d(v, sd[q + "phi.weight"], sd[q + "phi.bias"]).view(n, c, -1)
    f = torch.matmul(theta_x, phi_x)
    y = torch.matmul(f / f.size(-1), g_x).permute(0, 2, 1).contiguous().view(n, c, *v.shape[2:])
    w_y = F.conv3d(y, sd[q + "W_z.0.weight"], sd[q + "W_z.0.bias"])
    w_y = F.batch_norm(w_y, sd[q + "W_z.1.running_mean"], sd[q + "W_z.1.running_var"], sd[q + "W_z.1.weight"],
                       sd[q + "W_z.1.bias"], train, BN_MOMENTUM, BN_EPS)
    z = w_y + v
    z = torch.cat([t.squeeze(2) for t in torch.split(z, 1, dim=2)], dim=0)
    logits = conv2d(torch.cat((emb, z), dim=1), sd["last_layer.weight"], sd["last_layer.bias"])
    per_frame = torch.split(logits, n, dim=0)
    if seg_size is not None:
        return {"logits": logits, "probs": [F.softmax(bilinear(p, seg_size), dim=1) for p in per_frame]}
    losses, accs = [], []
    for p, lab in zip(per_frame, labels):
        l, lp, lb = nll_up(p, lab, ignore_index)
        losses.append(l)
        accs.append(pixel_acc(lp, lb))
    return {"logits": logits, "loss": sum(losses) / len(losses), "acc": sum(accs) / len(accs)}


# --------------------------------------------------------------------------------------------------
def region_gather(feats, probs, T):
    """SpatialTemporalGather_Module.forward without memory (spatial_ocr_block.py:95-109)."""
    n = feats.shape[0] // T
    ctxs = []
    for pr, ft in zip(torch.split(probs, n, dim=0), torch.split(feats, n, dim=0)):
        b, c = pr.shape[0], pr.shape[1]
        pr = F.softmax(pr.reshape(b, c, -1), dim=2)
        ft = ft.reshape(b, ft.shape[1], -1).permute(0, 2, 1)
        ctxs.append(torch.matmul(pr, ft).permute(0, 2, 1).unsqueeze(3).unsqueeze(0))
    return torch.mean(torch.cat(ctxs, dim=0), dim=0)


def _cbr(sd, prefix, x, train, idx=0):
    x = conv2d(x, sd[f"{prefix}.{idx}.weight"], sd.get(f"{prefix}.{idx}.bias"))
    return _act(batch_norm(sd, f"{prefix}.{idx + 1}", x, train))


def object_attention(sd, prefix, x, proxy, train, key_channels=256):
    """_ObjectAttentionBlock.forward (spatial_ocr_block.py:247-289), scale=1."""
    b = x.shape[0]
    q = _cbr(sd, prefix + ".f_pixel", _cbr(sd, prefix + ".f_pixel", x, train, 0), train, 3)
    k = _cbr(sd, prefix + ".f_object", _cbr(sd, prefix + ".f_object", proxy, train, 0), train, 3)
    v = _cbr(sd, prefix + ".f_down", proxy, train, 0)
    query = q.reshape(b, key_channels, -1).permute(0, 2, 1)
    key = k.reshape(b, key_channels, -1)
    value = v.reshape(b, key_channels, -1).permute(0, 2, 1)
    sim = F.softmax((key_channels ** -0.5) * torch.matmul(query, key), dim=-1)
    ctx = torch.matmul(sim, value).permute(0, 2, 1).contiguous().reshape(b, key_channels, *x.shape[2:])
    return _cbr(sd, prefix + ".f_up", ctx, train, 0)


def clip_ocr_forward(sd, frames, labels=None, deep_sup_scale=0.4, train=True, seg_size=None, ignore_index=255,
                     memory=None, memory_num=8, dropout_masks=None):
    """ClipOCRNet.forward with clipocr_all=False (clip_ocr.py:106-133,166-198).  `memory`: None or the
    shared python list of the inference memory bank (quirk Q9 is reproduced)."""
    T = len(frames)
    n = frames[0].shape[0]
    maps = resnet_forward(sd, "encoder.", torch.cat(frames, dim=0), train)
    d = _act(batch_norm(sd, "dsn_head.1", conv2d(maps[-2], sd["dsn_head.0.weight"], sd["dsn_head.0.bias"], 1, 1, 1), train))
    if dropout_masks and dropout_masks.get("dsn") is not None:
        d = d * dropout_masks["dsn"][:, :, None, None]
    x_dsn = conv2d(d, sd["dsn_head.4.weight"], sd["dsn_head.4.bias"])
    feats = _act(batch_norm(sd, "conv_3x3.1", conv2d(maps[-1], sd["conv_3x3.0.weight"], sd["conv_3x3.0.bias"], 1, 1, 1), train))
    if memory is None:
        context = region_gather(feats, x_dsn, T)
    else:
        bank = memory
        if len(bank) > 0:
            bank = [m.detach() for m in bank]
        for pr, ft in zip(torch.split(x_dsn, n, dim=0), torch.split(feats, n, dim=0)):
            ctx = region_gather(ft, pr, 1).unsqueeze(0)
            while len(bank) > memory_num:
                bank.pop(0)
            bank.append(ctx)
        context = torch.mean(torch.cat(bank, dim=0), dim=0)
    x = torch.split(feats, n, dim=0)[-1]
    ctx = object_attention(sd, "spatial_ocr_head.object_context_block", x, context, train)
    z = conv2d(torch.cat([ctx, x], 1), sd["spatial_ocr_head.conv_bn_dropout.0.weight"], sd["spatial_ocr_head.conv_bn_dropout.0.bias"])
    z = _act(batch_norm(sd, "spatial_ocr_head.conv_bn_dropout.1", z, train))
    if dropout_masks and dropout_masks.get("ocr") is not None:
        z = z * dropout_masks["ocr"][:, :, None, None]
    logits = conv2d(z, sd["head.weight"], sd["head.bias"])
    if seg_size is not None:
        return {"logits": logits, "probs": F.softmax(bilinear(logits, seg_size), dim=1), "context": context}
    loss, lp, lab = nll_up(logits, labels[-1], ignore_index)
    loss_ds, _, _ = nll_up(x_dsn, torch.cat(labels, dim=0), ignore_index)
    return {"logits": logits, "logits_deepsup": x_dsn, "loss_main": loss, "loss": loss + loss_ds * deep_sup_scale,
            "acc": pixel_acc(lp, lab), "context": context}


# --------------------------------------------------------------------------------------------------
def segmentation_module_forward(sd, img, label=None, deep_sup_scale=0.4, train=True, seg_size=None, ignore_index=255,
                                pool_scales=(1, 2, 3, 6), dropout_masks=None):
    """SegmentationModule.forward with a (dilated) ResNet encoder and PPMDeepsup decoder
    (models.py:74-111, :938-995)."""
    maps = resnet_forward(sd, "encoder.", img, train)
    conv5 = maps[-1]
    ppm_out = [conv5]
    for i, s in enumerate(pool_scales):
        z = conv2d(F.adaptive_avg_pool2d(conv5, s), sd[f"decoder.ppm.{i}.1.weight"])
        ppm_out.append(bilinear(_act(batch_norm(sd, f"decoder.ppm.{i}.2", z, train)), conv5.shape[-2:]))
    z = conv2d(torch.cat(ppm_out, 1), sd["decoder.conv_last_.0.weight"], None, 1, 1, 1)
    z = _act(batch_norm(sd, "decoder.conv_last_.1", z, train))
    if dropout_masks and dropout_masks.get("ppm") is not None:
        z = z * dropout_masks["ppm"][:, :, None, None]
    logits = conv2d(z, sd["decoder.conv_last_.4.weight"], sd["decoder.conv_last_.4.bias"])
    if seg_size is not None:
        return {"logits": logits, "probs": F.softmax(bilinear(logits, seg_size), dim=1)}
    loss, lp, lab = nll_up(logits, label, ignore_index)
    out = {"logits": logits, "loss_main": loss}
    if deep_sup_scale is not None:
        d = _act(batch_norm(sd, "decoder.cbr_deepsup.1", conv2d(maps[-2], sd["decoder.cbr_deepsup.0.weight"], None, 1, 1, 1), train))
        if dropout_masks and dropout_masks.get("deepsup") is not None:
            d = d * dropout_masks["deepsup"][:, :, None, None]
        lds = conv2d(d, sd["decoder.conv_last_deepsup_.weight"], sd["decoder.conv_last_deepsup_.bias"])
        loss_ds, _, _ = nll_up(lds, label, ignore_index)
        out["logits_deepsup"] = lds
        loss = loss + loss_ds * deep_sup_scale
    out["loss"] = loss
    out["acc"] = pixel_acc(lp, lab)
    return out


# --------------------------------------------------------------------------------------------------
class Evaluator:
    """utils.py:55-107 (confusion-matrix metrics), numpy on the host."""

    def __init__(self, num_class):
        self.num_class = num_class
        self.confusion_matrix = np.zeros((num_class, num_class))

    def add_batch(self, gt, pred):
        assert gt.shape == pred.shape
        mask = (gt >= 0) & (gt < self.num_class)
        label = self.num_class * gt[mask].astype("int") + pred[mask]
        self.confusion_matrix += np.bincount(label, minlength=self.num_class ** 2).reshape(self.num_class, self.num_class)

    def pixel_accuracy(self):
        return np.diag(self.confusion_matrix).sum() / self.confusion_matrix.sum()

    def mean_iou(self):
        cm = self.confusion_matrix
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))
        isval = cm.sum(axis=1) > 0
        return np.nansum(iou * isval) / isval.sum()

    def fw_iou(self):
        cm = self.confusion_matrix
        freq = cm.sum(axis=1) / cm.sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            iu = np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))
        return (freq[freq > 0] * iu[freq > 0]).sum()


# --------------------------------------------------------------------------------------------------
# shared synthetic-workload recipe: used by the golden generator, the tests and bench.py
def synthetic_clip(T, n, H, W, num_class=124, seed=304, block=32, ignore_frac=0.05, ignore_index=255):
    """T image tensors (n,3,H,W) ~ N(0,1) and T label tensors (n,1,H,W) float: piece-wise constant
    `block` x `block` tiles of uniform classes with `ignore_frac` of the tiles set to ignore_index."""
    g = torch.Generator().manual_seed(seed)
    imgs = [torch.randn(n, 3, H, W, generator=g) for _ in range(T)]
    labs = []
    bh, bw = (H + block - 1) // block, (W + block - 1) // block
    for _ in range(T):
        tiles = torch.randint(0, num_class, (n, 1, bh, bw), generator=g).float()
        drop = torch.rand(n, 1, bh, bw, generator=g) < ignore_frac
        tiles[drop] = float(ignore_index)
        lab = tiles.repeat_interleave(block, dim=2).repeat_interleave(block, dim=3)[:, :, :H, :W].contiguous()
        labs.append(lab)
    return imgs, labs


def grad_sample_indices(numel, k=2048, seed=99):
    """Seeded sample of `k` flat indices of a tensor with `numel` elements (all of them when numel <= k): the mid-size
    golden fixtures pin gradient tensors on such samples (rel-L2 over a uniform sample estimates rel-L2 of the tensor)."""
    if numel <= k:
        return torch.arange(numel)
    g = torch.Generator().manual_seed(seed + numel % 9973)
    return torch.randint(0, numel, (k,), generator=g)


def condition_nonlocal(sd, seed=8):
    """NLBlockND's output BN is initialised to weight = bias = 0 (the block is the identity at init, non_local.py:62-63),
    which would hide the whole affinity path from a parity fixture: give it live values.  In place; returns sd."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if k.endswith("W_z.1.weight"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("W_z.1.bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return sd


def condition_weights(sd, bn3_gamma=0.25, seed=7):
    """Parity conditioning: residual-branch bn3.weight <- 0.25 tames the
    train-mode amplification of the random-init network; running stats are randomised so that folded
    (eval-mode) BN bugs are visible.  In place; returns sd."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if k.endswith("bn3.weight") and ".layer" in k:
            v.fill_(bn3_gamma)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    return sd
