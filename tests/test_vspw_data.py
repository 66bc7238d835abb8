"""CPU: the VSPW clip loaders against the reference's dataset2.py classes on a generated directory —
same global RNG seeds in, bit-identical tensors out (sampling order, flips, multi-scale, padding, crop, label remap)."""
import argparse
import importlib.util
import os
import random
import sys

import numpy as np
import pytest
import torch
from PIL import Image

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")


def _ref_dataset2():
    spec = importlib.util.spec_from_file_location("ref_dataset2", os.path.join(REF, "dataset2.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def fake_vspw(tmp_path_factory):
    root = tmp_path_factory.mktemp("vspw")
    rng = np.random.RandomState(0)
    videos = {"vid_a": (14, 60, 80), "vid_b": (9, 48, 96), "vid_short": (3, 40, 40)}
    for v, (frames, h, w) in videos.items():
        os.makedirs(root / "data" / v / "origin")
        os.makedirs(root / "data" / v / "mask")
        os.makedirs(root / "data" / v / "mask_42label")
        for i in range(frames):
            Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)).save(root / "data" / v / "origin" / f"{i:08d}.jpg", quality=95)
            Image.fromarray(rng.randint(0, 125, (h, w), dtype=np.uint8)).save(root / "data" / v / "mask" / f"{i:08d}.png")
            Image.fromarray(rng.randint(0, 43, (h, w), dtype=np.uint8)).save(root / "data" / v / "mask_42label" / f"{i:08d}.png")
    for split in ("train", "val"):
        (root / f"{split}.txt").write_text("".join(v + "\n" for v in videos))
    return str(root)


def _args(root, **kw):
    base = dict(cropsize=56, dataroot=root, trainfps=1, clip_num=4, dilation2="1,2,4", multi_scale=True, lesslabel=False)
    base.update(kw)
    return argparse.Namespace(**base)


@pytest.mark.parametrize("multi_scale", [True, False])
def test_train_clips_are_bit_identical_to_the_reference(fake_vspw, multi_scale):
    from cvpr2021_vspw_implement_b200.vspw_data import VSPWClipTrain
    R = _ref_dataset2()
    a = R.BaseDataset_longclip(_args(fake_vspw, multi_scale=multi_scale), "train")
    b = VSPWClipTrain(_args(fake_vspw, multi_scale=multi_scale), "train")
    assert len(a) == len(b) == 3
    for seed in range(6):
        for idx in range(3):
            outs = []
            for ds in (a, b):
                np.random.seed(seed * 10 + idx)
                random.seed(seed * 10 + idx)
                outs.append(ds[idx])
            (ia, la), (ib, lb) = outs
            assert len(ia) == len(ib) == 4
            for x, y in zip(ia + la, ib + lb):
                assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y)
            assert tuple(ib[0].shape) == (3, 56, 56) and tuple(lb[0].shape) == (1, 56, 56)
            vals = torch.unique(lb[0])
            assert ((vals <= 123) | (vals == 255)).all()


@pytest.mark.parametrize("lesslabel", [False, True])
def test_test_items_are_bit_identical_to_the_reference(fake_vspw, lesslabel):
    from cvpr2021_vspw_implement_b200.vspw_data import VSPWClipTest
    R = _ref_dataset2()
    args = _args(fake_vspw, lesslabel=lesslabel)
    for video in ("vid_a", "vid_b"):
        a = R.TestDataset_longclip(fake_vspw, video, args, is_train=False)
        b = VSPWClipTest(fake_vspw, video, args, is_train=False)
        assert len(a) == len(b)
        for i in range(len(a)):
            xa, xb = a[i], b[i]
            assert xa[4] == xb[4]
            assert torch.equal(xa[0], xb[0]) and torch.equal(xa[1], xb[1])
            for p, q in zip(xa[2] + xa[3], xb[2] + xb[3]):
                assert torch.equal(p, q)


@pytest.mark.parametrize("method", ["nonlocal3d", "netwarp"])
@pytest.mark.parametrize("clip_num,dilation_num", [(5, 0), (4, 0), (4, 1), (5, 2), (2, 0)])
def test_window_items_are_bit_identical_to_the_reference(fake_vspw, method, clip_num, dilation_num):
    """VSPWWindowTest == dataset2.TestDataset_clip (:154-337): window placement and clamping at both ends, dilation
    sub-lists, the frame itself inside (nonlocal3d) or left out of (other methods) its window, the names `test_all` keys
    on, and the sub-list-shorter-than-2 corner (vid_short with dilation_num 2: one frame per sub-list)."""
    from cvpr2021_vspw_implement_b200.vspw_data import VSPWWindowTest
    R = _ref_dataset2()
    args = _args(fake_vspw, clip_num=clip_num, dilation_num=dilation_num, method=method)
    for video, is_train in (("vid_a", False), ("vid_b", False), ("vid_short", False), ("vid_a", True)):
        a = R.TestDataset_clip(fake_vspw, video, args, is_train=is_train)
        b = VSPWWindowTest(fake_vspw, video, args, is_train=is_train)
        assert len(a) == len(b)
        for i in range(len(a)):
            xa, xb = a[i], b[i]
            assert xa[4] == xb[4] and type(xa[4]) is type(xb[4])
            assert torch.equal(xa[0], xb[0]) and torch.equal(xa[1], xb[1])
            assert len(xa[2]) == len(xb[2]) and len(xa[3]) == len(xb[3])
            for p, q in zip(xa[2] + xa[3], xb[2] + xb[3]):
                assert torch.equal(p, q)


def test_loader_contract_feeds_the_entry_point_batches(fake_vspw):
    from cvpr2021_vspw_implement_b200.vspw_data import VSPWClipTrain
    ds = VSPWClipTrain(_args(fake_vspw), "train")
    imgs, gts = next(iter(torch.utils.data.DataLoader(ds, batch_size=2, drop_last=True)))
    assert len(imgs) == 4 and tuple(imgs[0].shape) == (2, 3, 56, 56) and tuple(gts[0].shape) == (2, 1, 56, 56)
    with pytest.raises(AssertionError):
        VSPWClipTrain(_args(fake_vspw, dilation2="1,2"), "train")
