"""GPU: the device-side end of the data path.  `VSPWClipTrain(device_finish=True)` ships uint8 crops and
raw masks; `DevicePrefetcher(finish_u8=True)` turns them into the model's tensors on the GPU (vspw_clip_finish_u8).  Under the
same RNG seeds the result must be BIT-IDENTICAL to the host transform (= the reference's dataset2.py:962-977, which
tests/test_vspw_data.py pins against the reference classes)."""
import argparse
import os
import random

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fake_vspw(tmp_path_factory):
    root = tmp_path_factory.mktemp("vspw_gpu")
    rng = np.random.RandomState(1)
    for v, (frames, h, w) in {"va": (9, 60, 80), "vb": (7, 48, 96), "vc": (6, 90, 70)}.items():
        os.makedirs(root / "data" / v / "origin")
        os.makedirs(root / "data" / v / "mask")
        for i in range(frames):
            Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)).save(root / "data" / v / "origin" / f"{i:08d}.jpg", quality=95)
            m = rng.randint(0, 125, (h, w), dtype=np.uint8)
            m[:3, :5] = 255  # a raw 255 must stay "ignore"
            Image.fromarray(m).save(root / "data" / v / "mask" / f"{i:08d}.png")
    (root / "train.txt").write_text("va\nvb\nvc\n")
    return str(root)


@pytest.mark.parametrize("multi_scale,cropsize", [(True, 56), (False, 72)])
def test_device_finish_is_bit_identical_to_the_host_transform(fake_vspw, multi_scale, cropsize):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cvpr2021_vspw_implement_b200.data import DevicePrefetcher
    from cvpr2021_vspw_implement_b200.vspw_data import VSPWClipTrain
    args = argparse.Namespace(cropsize=cropsize, dataroot=fake_vspw, trainfps=1, clip_num=3, dilation2="1,2", multi_scale=multi_scale,
                              lesslabel=False)
    host, dev = VSPWClipTrain(args, "train"), VSPWClipTrain(args, "train", device_finish=True)

    def batches(ds):
        random.seed(7); np.random.seed(7); torch.manual_seed(7)
        return list(torch.utils.data.DataLoader(ds, batch_size=3, shuffle=False, num_workers=0))

    ref = batches(host)
    u8 = batches(dev)
    assert u8[0][0][0].dtype == torch.uint8 and u8[0][1][0].dtype == torch.uint8
    got = list(DevicePrefetcher(iter(u8), torch.device("cuda", 0), finish_u8=True))
    torch.cuda.synchronize()
    assert len(got) == len(ref)
    for (ri, rl), (gi, gl) in zip(ref, got):
        for a, b in zip(ri, gi):
            assert b.dtype == torch.float32 and tuple(b.shape) == tuple(a.shape) and torch.equal(a, b.cpu())
        for a, b in zip(rl, gl):
            assert tuple(b.shape) == tuple(a.shape) and torch.equal(a, b.cpu())
            assert float(b.max()) <= 255.0
