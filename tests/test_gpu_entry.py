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


def test_test_all_sliding_window_average(need_gpu):
    """test_clip2.test_all (reference test_clip2.py:90-195) on a scripted module: every frame is decided from the mean of
    the first clip_num probability maps it receives (all of them for frames that never collect clip_num), and enters the
    evaluator exactly once; checked against a direct restatement over the same windows."""
    import argparse
    import numpy as np
    test_clip2 = _entry("test_clip2")
    from cvpr2021_vspw_implement_b200.data import SyntheticWindowTest
    from cvpr2021_vspw_implement_b200.utils import Evaluator
    K, H, W, T, F = 7, 12, 10, 3, 8
    args = argparse.Namespace(clip_num=T, num_class=K, dilation2="1,2", dilation_num=0, method="nonlocal3d", start_gpu=0, is_save=False,
                              saveroot="")
    ds = SyntheticWindowTest(args, "synthetic_000", frames=F, height=H, width=W, seed=3)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False)

    class Scripted(torch.nn.Module):
        """per-frame 'probabilities' that depend on the frame's content AND on its position in the window"""
        def forward(self, feed, segSize=None):
            outs = []
            for pos, f in enumerate(feed["clipimgs_data"]):
                g = torch.Generator().manual_seed(pos + 1)
                base = torch.rand(K, H, W, generator=g).to(f.device)
                key = f.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1)  # identifies the frame
                outs.append(torch.softmax(base.unsqueeze(0) * (1.0 + 10.0 * torch.sin(37.0 * key)), dim=1))
            return outs

    ev, evv = Evaluator(K), Evaluator(K)
    gts, preds, h, w = test_clip2.test_all(Scripted(), loader, 0, args, ev, evv, "synthetic_000")
    assert (h, w) == (H, W) and len(preds) == F == len(gts)
    # direct restatement
    maps, order = {}, []
    mod = Scripted()
    for b0 in range(0, F, 2):  # the loader's batches; arrivals are window-position-major, then batch item, as in test_all
        items = [ds[i] for i in range(b0, min(b0 + 2, F))]
        outs = [mod({"clipimgs_data": [c.unsqueeze(0).cuda() for c in it[2]]}) for it in items]
        for pos in range(T):
            for it, out in zip(items, outs):
                n = it[4][pos]
                maps.setdefault(n, [])
                if len(maps[n]) < T:
                    maps[n].append(out[pos])
                    if len(maps[n]) == T:
                        order.append(n)
    order += [n for n in maps if n not in order]
    assert len(order) == F
    ev2 = Evaluator(K)
    for k, n in enumerate(order):
        want = torch.argmax(torch.cat(maps[n], 0).mean(0, keepdim=True), 1)[0].cpu().numpy()
        assert np.array_equal(preds[k][0].cpu().numpy(), want), n
        frame_gt = ds.frames[int(n.split(".")[0])][1].squeeze(0).numpy()
        assert np.array_equal(gts[k][0].cpu().numpy(), frame_gt)
        ev2.add_batch(frame_gt[None], want[None])
    ev.sync_device()
    assert np.array_equal(ev.confusion_matrix, ev2.confusion_matrix)


def test_test_entry_point_nonlocal3d(need_gpu, tmp_path, monkeypatch):
    """test_clip2.py --method nonlocal3d end to end: Non_local3d's per-frame probability maps through
    the sliding-window loop, metrics and PNG dump, on a synthetic video with random weights."""
    test_clip2 = _entry("test_clip2")
    from cvpr2021_vspw_implement_b200.config import cfg, get_defaults
    monkeypatch.chdir(tmp_path)
    cfg.clear()
    cfg.update(get_defaults())
    yaml = os.path.join(ROOT, "config", "vsp-resnet101dilated-ppm_deepsup_clip.yaml")
    targv = ["--cfg", yaml, "--method", "nonlocal3d", "--clip_num", "3", "--dilation2", "3,6", "--batchsize", "2",
             "--synthetic", "True", "--synthetic_size", "64x96", "--synthetic_videos", "1", "--synthetic_frames", "7",
             "--vc_clip_num", "2", "--saveroot", str(tmp_path / "pred"), "--is_save", "True", "MODEL.arch_encoder", "resnet50dilated"]
    targs = test_clip2.make_parser().parse_args(targv)
    targs.max_distances = [10]
    cfg.merge_from_list(targs.opts)
    res = test_clip2.main(cfg, 0, targs)
    assert 0.0 <= res["mIoU"] <= 1.0 and 0.0 <= res["Acc"] <= 1.0
    assert len(os.listdir(tmp_path / "pred" / "synthetic_000")) == 7


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
