"""GPU: the train_clip2.py / test_clip2.py entry points end to end on synthetic clips (reference call stacks
SURVEY.md 3.1 / 3.4): training steps lower the loss, checkpoints round-trip through the 'module.'-prefix convention,
and the inference entry reports the reference's metrics."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _entry(name):
    """Import this repo's entry point by path (another test puts the reference tree, which has files of the same
    name, in front of sys.path)."""
    import importlib.util
    if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(ROOT):
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


@pytest.mark.parametrize("method", ["clip_psp", "clip_ocr"])
def test_train_then_test_entry_points(need_gpu, method, tmp_path, monkeypatch):
    train_clip2 = _entry("train_clip2")
    test_clip2 = _entry("test_clip2")
    from cvpr2021_vspw_implement_b200.config import cfg, get_defaults
    monkeypatch.chdir(tmp_path)
    cfg.clear()
    cfg.update(get_defaults())
    yaml = os.path.join(ROOT, "config", "vsp-resnet101dilated-ppm_deepsup_clip.yaml")
    save = str(tmp_path / "ckpt")
    argv = ["--cfg", yaml, "--method", method, "--clip_num", "3", "--dilation2", "3,6", "--batchsize", "2", "--gpu_num", "1",
            "--lr", "0.01", "--totalepoch", "4", "--synthetic", "True", "--synthetic_size", "64x96", "--synthetic_clips", "2",
            "--saveroot", save, "--checkpoint_every", "1", "--precision", "bf16x3",
            "MODEL.arch_encoder", "resnet50dilated", "TRAIN.seed", "5"]
    args = train_clip2.make_parser().parse_args(argv)
    train_clip2.configure(args)
    hist = train_clip2.main(cfg, args)
    losses = hist["train"]["loss"]
    assert len(losses) == 4 and all(l == l and l < 20 for l in losses)
    assert losses[-1] < losses[0]  # four SGD steps on the same two clips: the loss must move down
    ck = os.path.join(save, "model_epoch_4.pth")
    sd = torch.load(ck, map_location="cpu")
    assert all(k.startswith("module.") for k in sd) and os.path.exists(os.path.join(save, "opt_epoch_4.pth"))

    targv = ["--cfg", yaml, "--method", method, "--clip_num", "3", "--dilation2", "3,6", "--batchsize", "2", "--load", ck,
             "--synthetic", "True", "--synthetic_size", "64x96", "--synthetic_videos", "1", "--synthetic_frames", "6",
             "--vc_clip_num", "2", "--use_memory", "True" if method == "clip_ocr" else "False", "--saveroot", str(tmp_path / "pred"),
             "--is_save", "True", "MODEL.arch_encoder", "resnet50dilated"]
    targs = test_clip2.make_parser().parse_args(targv)
    targs.max_distances = [10]
    cfg.merge_from_list(targs.opts)
    res = test_clip2.main(cfg, 0, targs)
    assert 0.0 <= res["mIoU"] <= 1.0 and 0.0 <= res["Acc"] <= 1.0 and res["VC"] == res["VC"]
    assert len(os.listdir(tmp_path / "pred" / "synthetic_000")) == 6


def test_train_entry_point_nonlocal3d(need_gpu, tmp_path, monkeypatch):
    """--method nonlocal3d through the training entry point: the whole clip is fed and supervised."""
    train_clip2 = _entry("train_clip2")
    from cvpr2021_vspw_implement_b200.config import cfg, get_defaults
    monkeypatch.chdir(tmp_path)
    cfg.clear()
    cfg.update(get_defaults())
    yaml = os.path.join(ROOT, "config", "vsp-resnet101dilated-ppm_deepsup_clip.yaml")
    argv = ["--cfg", yaml, "--method", "nonlocal3d", "--clip_num", "3", "--dilation2", "3,6", "--batchsize", "2", "--gpu_num", "1",
            "--lr", "0.01", "--totalepoch", "4", "--synthetic", "True", "--synthetic_size", "64x96", "--synthetic_clips", "2",
            "MODEL.arch_encoder", "resnet50dilated", "TRAIN.seed", "5"]
    args = train_clip2.make_parser().parse_args(argv)
    train_clip2.configure(args)
    losses = train_clip2.main(cfg, args)["train"]["loss"]
    assert len(losses) == 4 and all(l == l for l in losses) and losses[-1] < losses[0]


def test_entry_points_on_a_vspw_directory(need_gpu, tmp_path, monkeypatch):
    """train_clip2.py --dataroot / test_clip2.py --dataroot end to end on a generated VSPW-layout directory (JPEG frames,
    PNG masks, split lists): the f3 data path feeding the hot path through the reference's CLI."""
    import numpy as np
    from PIL import Image
    train_clip2 = _entry("train_clip2")
    test_clip2 = _entry("test_clip2")
    from cvpr2021_vspw_implement_b200.config import cfg, get_defaults
    root = tmp_path / "vspw"
    rng = np.random.RandomState(1)
    for v in ("v0", "v1"):
        os.makedirs(root / "data" / v / "origin")
        os.makedirs(root / "data" / v / "mask")
        for i in range(8):
            Image.fromarray(rng.randint(0, 256, (72, 104, 3), dtype=np.uint8)).save(root / "data" / v / "origin" / f"{i:08d}.jpg")
            Image.fromarray(np.repeat(np.repeat(rng.randint(0, 125, (9, 13), dtype=np.uint8), 8, 0), 8, 1)).save(
                root / "data" / v / "mask" / f"{i:08d}.png")
    for split in ("train", "val"):
        (root / f"{split}.txt").write_text("v0\nv1\n")
    monkeypatch.chdir(tmp_path)
    cfg.clear()
    cfg.update(get_defaults())
    yaml = os.path.join(ROOT, "config", "vsp-resnet101dilated-ppm_deepsup_clip.yaml")
    save = str(tmp_path / "ckpt")
    argv = ["--cfg", yaml, "--method", "clip_psp", "--clip_num", "3", "--dilation2", "1,3", "--batchsize", "2", "--gpu_num", "1",
            "--lr", "0.01", "--totalepoch", "2", "--dataroot", str(root), "--cropsize", "64", "--multi_scale", "True",
            "--saveroot", save, "--checkpoint_every", "2", "MODEL.arch_encoder", "resnet50dilated"]
    args = train_clip2.make_parser().parse_args(argv)
    train_clip2.configure(args)
    losses = train_clip2.main(cfg, args)["train"]["loss"]
    assert len(losses) == 2 and all(l == l and l < 20 for l in losses)
    targv = ["--cfg", yaml, "--method", "clip_psp", "--clip_num", "3", "--dilation2", "1,3", "--batchsize", "2",
             "--load", os.path.join(save, "model_epoch_2.pth"), "--dataroot", str(root), "--split", "val", "--vc_clip_num", "2",
             "--saveroot", str(tmp_path / "pred"), "MODEL.arch_encoder", "resnet50dilated"]
    targs = test_clip2.make_parser().parse_args(targv)
    targs.max_distances = [10]
    cfg.merge_from_list(targs.opts)
    res = test_clip2.main(cfg, 0, targs)
    assert 0.0 <= res["mIoU"] <= 1.0 and 0.0 < res["Acc"] <= 1.0
