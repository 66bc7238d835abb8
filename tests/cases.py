"""Shared case recipes for the parity tests: the same seeds oracle/make_golden.py used."""
import argparse
import os

import numpy as np
import torch

import tcb_oracle as O
from cvpr2021_vspw_implement_b200 import models as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NUM_CLASS = 124
CASES = {
    "clip_psp": ("Clip_PSP", "resnet50dilated", 3, 2, 49, 65, 11, 304),
    "clip_psp_pspw": ("Clip_PSP_pspw", "resnet50dilated", 3, 2, 49, 65, 12, 305),
    "clip_ocr": ("ClipOCRNet", "resnet50dilated", 3, 2, 49, 65, 13, 306),
    "segmodule_r18": ("SegmentationModule", "resnet18dilated", 1, 2, 49, 65, 14, 307),
    "non_local3d": ("Non_local3d", "resnet50dilated", 3, 2, 49, 65, 15, 308),
    # image-model family on the same kernels: "Seg/<decoder>/<fc_dim>/<deep_sup_scale or none>"
    "seg_ocrnet_r50": ("Seg/ocrnet_deepsup/2048/0.4", "resnet50dilated", 1, 2, 49, 65, 21, 311),
    "seg_upernet_r50": ("Seg/upernet_lite/2048/none", "resnet50", 1, 2, 65, 97, 22, 312),
    "seg_c1ds_r18": ("Seg/c1_deepsup/512/0.4", "resnet18dilated", 1, 2, 49, 65, 23, 313),
    "seg_ppm_r18": ("Seg/ppm/512/none", "resnet18dilated", 1, 2, 49, 65, 24, 314),
}

# mid-size train-mode fixtures for the gradient gates (oracle/make_golden.py MID_CASES)
MID_CASES = {
    # the same two with every ReLU replaced by the identity (no mask flips: gradient parity at kernel-level tolerances)
    "clip_psp_mid_norelu": ("Clip_PSP", "resnet50dilated", 3, 4, 97, 129, 16, 309),
    "clip_ocr_mid_norelu": ("ClipOCRNet", "resnet50dilated", 3, 4, 97, 129, 17, 310),
    "clip_psp_mid": ("Clip_PSP", "resnet50dilated", 3, 4, 97, 129, 16, 309),
    "clip_ocr_mid": ("ClipOCRNet", "resnet50dilated", 3, 4, 97, 129, 17, 310),
}

# cases driven through the img_data / clipimgs_data feed (Non_local3d has its own: every frame is supervised)
CLIP_CASES = [k for k in CASES if k != "non_local3d" and not k.startswith("seg_")]
SEG_CASES = [k for k in CASES if k.startswith("seg_")]


def ns(**kw):
    base = dict(num_class=NUM_CLASS, psp_weight=False, use_memory=False, memory_num=8, clipocr_all=False)
    base.update(kw)
    return argparse.Namespace(**base)


def build(kind, arch, seed, conditioned=True, **kw):
    """Our host mirror, constructed exactly like oracle/make_golden.py constructs the reference."""
    torch.manual_seed(seed)
    crit = torch.nn.NLLLoss(ignore_index=255)
    enc = M.ModelBuilder.build_encoder(arch)
    if kind == "Clip_PSP":
        m = M.Clip_PSP(enc, crit, ns(**kw), deep_sup_scale=0.4)
    elif kind == "Clip_PSP_pspw":
        m = M.Clip_PSP(enc, crit, ns(psp_weight=True, **kw), deep_sup_scale=0.4)
    elif kind == "ClipOCRNet":
        m = M.ClipOCRNet(enc, crit, ns(**kw), deep_sup_scale=0.4)
    elif kind == "Non_local3d":
        m = M.Non_local3d(ns(**kw), enc, crit)
    elif kind.startswith("Seg/"):
        _, dec_arch, fc, ds = kind.split("/")
        dec = M.ModelBuilder.build_decoder(dec_arch, fc_dim=int(fc), num_class=NUM_CLASS)
        m = M.SegmentationModule(enc, dec, crit, deep_sup_scale=None if ds == "none" else float(ds))
    else:
        dec = M.ModelBuilder.build_decoder("ppm_deepsup", fc_dim=512, num_class=NUM_CLASS)
        m = M.SegmentationModule(enc, dec, crit, deep_sup_scale=0.4)
    if conditioned:
        sd = m.state_dict()
        O.condition_weights(sd)
        O.condition_nonlocal(sd)
        m.load_state_dict(sd)
    return m


def no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.eval()
    return m


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def clip_inputs(name):
    kind, arch, T, n, H, W, mseed, dseed = CASES[name]
    return O.synthetic_clip(T, n, H, W, NUM_CLASS, seed=dseed, block=16)


def feed(imgs, labs, train, device=None):
    """Reference convention (train_clip2.py:75-83): frame 0 of the sampled clip is the current frame."""
    mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)
    d = {"img_data": mv(imgs[0]), "seg_label": mv(labs[0]), "clipimgs_data": [mv(i) for i in imgs[1:]], "step": 1}
    if train:
        d["cliplabels_data"] = [mv(l) for l in labs[1:]]
    return d


def oracle_order(imgs, labs):
    """The oracle takes frames with the CURRENT frame LAST (clip_psp.py:142-143)."""
    return list(imgs[1:]) + [imgs[0]], list(labs[1:]) + [labs[0]]


def rel_err(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64).reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
