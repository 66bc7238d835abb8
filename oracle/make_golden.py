"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (read-only at /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Every case is fully described by seeds: weights = the reference constructors under
``torch.manual_seed(seed)`` (our mirror reproduces them bit-for-bit, tests/test_host.py checks that
while the reference is mounted) + ``tcb_oracle.condition_weights``; inputs =
``tcb_oracle.synthetic_clip``.  The fixtures hold only the reference's OUTPUTS, so they stay small.
Dropout2d modules are put in eval mode (their masks come from torch's global RNG and are not part of
the algorithm under test).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("VSPW_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, os.path.join(REF, "RAFT_core"))

import tcb_oracle as O  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
CASES = {
    # name: (builder kind, arch, T, n, H, W, model seed, data seed)
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
# mid-size, well-conditioned train-mode fixtures for the GRADIENT gates (the 49x65 cases above put everything below
# layer2 on 7x9 maps, where a single ReLU mask flip is 4e-3 of a gradient norm: oracle/NOISE_FLOOR.md)
MID_CASES = {
    "clip_psp_mid_norelu": ("Clip_PSP", "resnet50dilated", 3, 4, 97, 129, 16, 309),
    "clip_ocr_mid_norelu": ("ClipOCRNet", "resnet50dilated", 3, 4, 97, 129, 17, 310),
    "clip_psp_mid": ("Clip_PSP", "resnet50dilated", 3, 4, 97, 129, 16, 309),
    "clip_ocr_mid": ("ClipOCRNet", "resnet50dilated", 3, 4, 97, 129, 17, 310),
}
NUM_CLASS = 124


def ns(**kw):
    base = dict(num_class=NUM_CLASS, psp_weight=False, use_memory=False, memory_num=8, clipocr_all=False)
    base.update(kw)
    return argparse.Namespace(**base)


def build(ref, kind, arch, seed, **kw):
    torch.manual_seed(seed)
    crit = torch.nn.NLLLoss(ignore_index=255)
    enc = ref.ModelBuilder.build_encoder(arch)
    if kind == "Clip_PSP":
        m = ref.Clip_PSP(enc, crit, ns(**kw), deep_sup_scale=0.4)
    elif kind == "Clip_PSP_pspw":
        m = ref.Clip_PSP(enc, crit, ns(psp_weight=True, **kw), deep_sup_scale=0.4)
    elif kind == "ClipOCRNet":
        m = ref.ClipOCRNet(enc, crit, ns(**kw), deep_sup_scale=0.4)
    elif kind == "Non_local3d":
        m = ref.Non_local3d(ns(**kw), enc, crit)
    elif kind.startswith("Seg/"):
        _, dec_arch, fc, ds = kind.split("/")
        dec = ref.ModelBuilder.build_decoder(dec_arch, fc_dim=int(fc), num_class=NUM_CLASS)
        m = ref.SegmentationModule(enc, dec, crit, deep_sup_scale=None if ds == "none" else float(ds))
    else:
        dec = ref.ModelBuilder.build_decoder("ppm_deepsup", fc_dim=512, num_class=NUM_CLASS)
        m = ref.SegmentationModule(enc, dec, crit, deep_sup_scale=0.4)
    sd = m.state_dict()
    O.condition_weights(sd)
    O.condition_nonlocal(sd)
    m.load_state_dict(sd)
    return m


def no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.eval()


def grad_summary(model):
    """Per-parameter gradient pins: L2 norm, sum, and the first 64 values (full tensor when small)."""
    out = {}
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().float().reshape(-1)
        out["gnorm/" + name] = np.float64(g.double().norm().item())
        out["gsum/" + name] = np.float64(g.double().sum().item())
        out["ghead/" + name] = g[:64].numpy().copy()
    return out


def feed(imgs, labs, train):
    # reference convention (train_clip2.py:75-83): frame 0 of the sampled clip is "current"
    d = {"img_data": imgs[0], "seg_label": labs[0], "clipimgs_data": list(imgs[1:]), "step": 1}
    if train:
        d["cliplabels_data"] = list(labs[1:])
    return d


def run_nonlocal_case(ref, name, spec):
    """Non_local3d: every frame is supervised, inference returns one probability map per frame."""
    kind, arch, T, n, H, W, mseed, dseed = spec
    imgs, labs = O.synthetic_clip(T, n, H, W, NUM_CLASS, seed=dseed, block=16)
    rec = {"meta": np.array([T, n, H, W, mseed, dseed])}
    for mode in ("train", "fixbn"):
        m = build(ref, kind, arch, mseed)
        m.train(mode == "train")
        captured = {}
        h = m.last_layer.register_forward_hook(lambda mod, i, o: captured.__setitem__("logits", o.detach()))
        loss, acc = m({"clipimgs_data": list(imgs), "cliplabels_data": list(labs)})
        loss.backward()
        h.remove()
        rec[mode + "/loss"] = np.float64(loss.item())
        rec[mode + "/acc"] = np.float64(acc.item())
        rec[mode + "/logits"] = captured["logits"].numpy().copy()
        rec.update({mode + "/" + k: v for k, v in grad_summary(m).items()})
        if mode == "train":
            sd = m.state_dict()
            for k in ("nonlocalblock.W_z.1.running_mean", "nonlocalblock.W_z.1.running_var"):
                rec["train/after/" + k] = sd[k].numpy().copy()
    m = build(ref, kind, arch, mseed)
    m.eval()
    with torch.no_grad():
        probs = m({"clipimgs_data": list(imgs), "cliplabels_data": list(labs)}, segSize=(H, W))
    rec["eval/probs_sub"] = torch.stack([p[:, :, ::4, ::4] for p in probs]).numpy().copy()
    rec["eval/pred"] = torch.stack([p.argmax(1) for p in probs]).numpy().astype(np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: loss={rec['train/loss']:.6f} acc={rec['train/acc']:.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")


def run_case(ref, name, spec):
    kind, arch, T, n, H, W, mseed, dseed = spec
    if kind == "Non_local3d":
        return run_nonlocal_case(ref, name, spec)
    imgs, labs = O.synthetic_clip(T, n, H, W, NUM_CLASS, seed=dseed, block=16)
    rec = {"meta": np.array([T, n, H, W, mseed, dseed])}
    # ---- train step -------------------------------------------------------------------------------
    m = build(ref, kind, arch, mseed)
    m.train()
    no_dropout(m)
    captured = {}
    hooks = []
    if kind.startswith("Clip_PSP"):
        hooks.append(m.ppm_conv.register_forward_hook(lambda mod, i, o: captured.__setitem__("logits", o.detach())))
        hooks.append(m.deepsup.register_forward_hook(lambda mod, i, o: captured.__setitem__("logits_deepsup", o.detach())))
    elif kind == "ClipOCRNet":
        hooks.append(m.head.register_forward_hook(lambda mod, i, o: captured.__setitem__("logits", o.detach())))
        hooks.append(m.dsn_head.register_forward_hook(lambda mod, i, o: captured.__setitem__("logits_deepsup", o.detach())))
        hooks.append(m.spatial_context_head.register_forward_hook(lambda mod, i, o: captured.__setitem__("context", o.detach())))
    elif kind.startswith("Seg/"):
        last = {"ocrnet_deepsup": "head", "upernet_lite": "conv_last_", "upernet": "conv_last_", "c1_deepsup": "conv_last_", "c1": "conv_last_1",
                "ppm": "conv_last"}[kind.split("/")[1]]
        hooks.append(getattr(m.decoder, last).register_forward_hook(lambda mod, i, o: captured.__setitem__("logits", o.detach())))
    else:
        hooks.append(m.decoder.conv_last_.register_forward_hook(lambda mod, i, o: captured.__setitem__("logits", o.detach())))
    single = kind == "SegmentationModule" or kind.startswith("Seg/")
    loss, acc = m(feed(imgs, labs, True)) if not single else m({"img_data": imgs[0], "seg_label": labs[0]})
    loss.backward()
    for h in hooks:
        h.remove()
    rec["train/loss"] = np.float64(loss.item())
    rec["train/acc"] = np.float64(acc.item())
    for k, v in captured.items():
        rec["train/" + k] = v.numpy().copy()
    rec.update({"train/" + k: v for k, v in grad_summary(m).items()})
    sd = m.state_dict()
    for k in ("encoder.bn1.running_mean", "encoder.bn1.running_var", "encoder.layer4.0.bn2.running_mean",
              "encoder.layer4.0.bn2.running_var"):
        rec["train/after/" + k] = sd[k].numpy().copy()
    # ---- frozen-BN step (cfg.TRAIN.fix_bn -> module.train(False), train_clip2.py:33): loss + gradients with
    #      running statistics; the network is not chaotic in this mode, so gradients pin tightly ----------
    m = build(ref, kind, arch, mseed)
    m.eval()
    loss, acc = m(feed(imgs, labs, True)) if not single else m({"img_data": imgs[0], "seg_label": labs[0]})
    loss.backward()
    rec["fixbn/loss"] = np.float64(loss.item())
    rec["fixbn/acc"] = np.float64(acc.item())
    rec.update({"fixbn/" + k: v for k, v in grad_summary(m).items()})
    # ---- eval forward -----------------------------------------------------------------------------
    m = build(ref, kind, arch, mseed)
    m.eval()
    with torch.no_grad():
        if single:
            if kind.startswith("Seg/"):
                m.decoder.use_softmax = True  # test.py builds the decoder with use_softmax=True for inference
            probs = m({"img_data": imgs[0], "seg_label": labs[0]}, segSize=(H, W))
        else:
            probs = m(feed(imgs, labs, False), segSize=(H, W))
    pred = probs.argmax(1)
    rec["eval/probs_sub"] = probs[:, :, ::4, ::4].numpy().copy()
    rec["eval/pred"] = pred.numpy().astype(np.uint8)
    ev = O.Evaluator(NUM_CLASS)
    import utils as ref_utils  # the reference's own Evaluator (utils.py:55-107)
    rev = ref_utils.Evaluator(NUM_CLASS)
    gt = labs[0].squeeze(1).numpy()
    rev.add_batch(gt, pred.numpy())
    ev.add_batch(gt, pred.numpy())
    rec["eval/miou"] = np.float64(rev.Mean_Intersection_over_Union())
    rec["eval/pixacc"] = np.float64(rev.Pixel_Accuracy())
    rec["eval/fwiou"] = np.float64(rev.Frequency_Weighted_Intersection_over_Union())
    assert abs(ev.mean_iou() - rec["eval/miou"]) < 1e-12 and abs(ev.fw_iou() - rec["eval/fwiou"]) < 1e-12
    # ---- OCR inference memory bank (quirk Q9): two consecutive calls of one "video" -----------------
    if kind == "ClipOCRNet":
        m = build(ref, kind, arch, mseed, use_memory=True, memory_num=2)
        m.eval()
        imgs2, _ = O.synthetic_clip(T, n, H, W, NUM_CLASS, seed=dseed + 1000, block=16)
        with torch.no_grad():
            d1 = feed(imgs, labs, False)
            d1["is_clean_memory"] = True
            p1 = m(d1, segSize=(H, W))
            d2 = feed(imgs2, labs, False)
            d2["is_clean_memory"] = False
            p2 = m(d2, segSize=(H, W))
        rec["mem/probs1_sub"] = p1[:, :, ::4, ::4].numpy().copy()
        rec["mem/probs2_sub"] = p2[:, :, ::4, ::4].numpy().copy()
        rec["mem/bank_len"] = np.array([len(m.memory)])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: loss={rec['train/loss']:.6f} acc={rec['train/acc']:.6f} miou={rec['eval/miou']:.6f} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def run_mid_case(ref, name, spec):
    """Train step and frozen-BN step (cfg.TRAIN.fix_bn): loss, acc, logits, and for EVERY parameter the gradient norm plus a
    seeded 2048-element sample of the gradient tensor (tcb_oracle.grad_sample_indices).  Also the reference's OWN fp32 floor:
    the same step with 1 oneDNN thread instead of 8 (summation order only) — stored per tensor (`gfloor`) so the test can
    gate the CUDA path relative to it."""
    kind, arch, T, n, H, W, mseed, dseed = spec
    imgs, labs = O.synthetic_clip(T, n, H, W, NUM_CLASS, seed=dseed, block=16)
    rec = {"meta": np.array([T, n, H, W, mseed, dseed])}
    for mode in ("train", "fixbn"):
        grads = {}
        for threads in (8, 1):
            torch.set_num_threads(threads)
            m = build(ref, kind, arch, mseed)
            m.train(mode == "train")
            no_dropout(m)
            if name.endswith("_norelu"):
                # ReLU-free variant of the same reference modules (every ReLU of the path is an nn.ReLU module): no mask flips,
                # so whole-model gradients are reproducible to ~1e-6 and the CUDA path can be gated at kernel-level tolerances
                for mod in m.modules():
                    if isinstance(mod, torch.nn.ReLU):
                        mod.forward = lambda x: x
            captured = {}
            head = m.ppm_conv if kind.startswith("Clip_PSP") else m.head
            h = head.register_forward_hook(lambda mod, i, o: captured.__setitem__("logits", o.detach()))
            loss, acc = m(feed(imgs, labs, True))
            loss.backward()
            h.remove()
            grads[threads] = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
            if threads == 8:
                rec[mode + "/loss"] = np.float64(loss.item())
                rec[mode + "/acc"] = np.float64(acc.item())
                rec[mode + "/logits"] = captured["logits"].numpy().copy()
        torch.set_num_threads(8)
        floor = {}
        for k, g in grads[8].items():
            g = g.float().reshape(-1)
            idx = O.grad_sample_indices(g.numel())
            rec[mode + "/gnorm/" + k] = np.float64(g.double().norm().item())
            rec[mode + "/gsample/" + k] = g[idx].numpy().copy()
            gn = g.double().norm().item()
            floor[k] = float((grads[1][k].reshape(-1).double() - g.double()).norm().item() / gn) if gn > 1e-7 else 0.0
            rec[mode + "/gfloor/" + k] = np.float64(floor[k])
        worst = max(floor.items(), key=lambda kv: kv[1])
        print(f"{name}/{mode}: loss={rec[mode + '/loss']:.6f}; reference fp32 floor (1 vs 8 threads) worst rel-L2 "
              f"{worst[1]:.2e} ({worst[0]}), median {float(np.median(list(floor.values()))):.2e}")
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name} -> {os.path.getsize(path) / 1024:.0f} KiB")


def bn_formula_pin():
    """The one numeric pin the reference's own tests hold for this path
    (lib/nn/modules/tests/test_numeric_batchnorm.py:29-52): train-mode BN = (x-mean)/sqrt(var_biased+eps),
    running_var uses the unbiased variance.  Stored as a tiny fixture for the BN kernel tests."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 8, 5, 7, generator=g) * 3 + 1
    w, b = torch.rand(8, generator=g) + 0.5, torch.randn(8, generator=g)
    rm, rv = torch.zeros(8), torch.ones(8)
    y = torch.nn.functional.batch_norm(x, rm, rv, w, b, True, 0.1, 1e-5)
    mean = x.mean(dim=(0, 2, 3))
    var_b = x.var(dim=(0, 2, 3), unbiased=False)
    y_formula = (x - mean[None, :, None, None]) / torch.sqrt(var_b + 1e-5)[None, :, None, None] * w[None, :, None, None] + b[None, :, None, None]
    assert torch.allclose(y, y_formula, atol=1e-5)
    np.savez_compressed(os.path.join(OUT, "bn_formula.npz"), x=x.numpy(), w=w.numpy(), b=b.numpy(), y=y.numpy(),
                        running_mean=rm.numpy(), running_var=rv.numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    import importlib
    ref = importlib.import_module("models")
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(ref, name, spec)
    for name, spec in MID_CASES.items():
        if only and name not in only:
            continue
        run_mid_case(ref, name, spec)
    bn_formula_pin()


if __name__ == "__main__":
    main()
