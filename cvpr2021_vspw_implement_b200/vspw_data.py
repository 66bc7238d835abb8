"""VSPW clip datasets for the entry points: the loaders on either side of the hot path.

Reference: dataset2.py — `BaseDataset_longclip` (:852-1048, training clips), `TestDataset_longclip` (:344-490,
per-video inference of the TCB models) and `TestDataset_clip` (:154-337, sliding-window clips for `test_all`).  Directory layout: `<dataroot>/<split>.txt` lists the videos; `<dataroot>/data/<video>/origin/*.jpg`
are the frames and `<dataroot>/data/<video>/mask/*.png` the label maps (`mask_42label/` with `--lesslabel`).

The sampling consumes NumPy's and Python's global RNGs in the reference's order (direction flip, start frame, mirror
flag, scale, crop x, crop y), so a run seeded like the reference draws the same clips; `tests/test_vspw_data.py` checks
tensors bit for bit against the reference classes on a generated directory.  Label convention (`segm_transform`,
:970-977): raw 0 -> 255 (ignore), otherwise raw - 1; float32 (1, H, W).
"""
import os
import random

import numpy as np
import torch
from PIL import Image

_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
_STD = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)


def image_to_tensor(img_hwc01):
    """float32 HWC in [0,1] -> ImageNet-normalised CHW (dataset2.py:962-968)."""
    t = torch.from_numpy(np.ascontiguousarray(img_hwc01.transpose((2, 0, 1))))
    return t.sub(_MEAN).div(_STD)


def labels_to_tensor(segm):
    """Raw VSPW mask -> (1, H, W) float labels in {0..K-1, 255} (dataset2.py:970-977)."""
    segm = np.array(segm)  # private copy: the remap is done in place like the reference
    segm[segm == 0] = 255
    segm = segm - 1
    segm[segm == 254] = 255
    return torch.from_numpy(segm).float().unsqueeze(0)


def _parse_dilation(args):
    dil = [int(d) for d in str(args.dilation2).split(",")]
    if len(dil) + 1 != int(args.clip_num):
        raise AssertionError(f"--dilation2 {args.dilation2!r} must list clip_num-1 = {int(args.clip_num) - 1} frame offsets")
    return dil


def _read_video_list(dataroot, split):
    with open(os.path.join(dataroot, split + ".txt")) as f:
        return [line[:-1] for line in f.readlines()]


class VSPWClipTrain(torch.utils.data.Dataset):
    """One item = one clip of `clip_num` frames from a random position of video `idx`: frame offsets `--dilation2` after a
    random start, random playback direction, shared mirror flip, optional shared multi-scale resize, zero/255 padding up
    to the crop size and ONE shared random crop.  Returns (list of (3,h,w) images, list of (1,h,w) labels); frame 0 is the
    clip's "current" frame (train_clip2.py:75-83)."""

    SCALES = [0.8, 1., 1.5, 2.0]

    def __init__(self, args, split="train", device_finish=False):
        """device_finish=True: items are uint8 — (h, w, 3) image crops and (h, w) RAW masks — and the float conversion,
        normalisation, HWC->CHW and label remap run on the GPU (`data.DevicePrefetcher(..., finish_u8=True)` ->
        vspw_clip_finish_u8), bit-identical to the host transform: the workers skip three fp32 passes per frame and the H2D copy
        moves a quarter of the bytes."""
        self.args = args
        self.device_finish = bool(device_finish)
        self.split = split
        self.crop = (int(args.cropsize), int(args.cropsize))
        self.dataroot = args.dataroot
        self.dilation = _parse_dilation(args)
        self.videos = _read_video_list(self.dataroot, split)
        self.frames = {v: sorted(os.listdir(os.path.join(self.dataroot, "data", v, "origin"))) for v in self.videos}

    def __len__(self):
        return len(self.videos)

    def _pad_and_crop(self, images, labels):
        h, w = images[0].shape[:2]
        ph = self.crop[0] - h if h < self.crop[0] else 0   # the reference pads BOTH sides by the full deficit
        pw = self.crop[1] - w if w < self.crop[1] else 0
        H, W = h + 2 * ph, w + 2 * pw
        x = random.randint(0, W - self.crop[1])
        y = random.randint(0, H - self.crop[0])
        out_i, out_l = [], []
        for im, lb in zip(images, labels):
            # raw masks (device_finish) are padded with the RAW ignore value 0, which the device remaps to 255
            lb = np.pad(lb, ((ph, ph), (pw, pw)), "constant", constant_values=(0, 0) if self.device_finish else (255, 255))
            im = np.pad(im, ((ph, ph), (pw, pw), (0, 0)), "constant")
            out_i.append(im[y:y + self.crop[0], x:x + self.crop[1]])
            out_l.append(lb[y:y + self.crop[0], x:x + self.crop[1]])
        return out_i, out_l

    def __getitem__(self, idx):
        video = self.videos[idx]
        names = self.frames[video]
        if np.random.random() < 0.5:
            names = names[::-1]
        starts = names[:-self.dilation[-1]]
        while len(starts) < 1:            # video shorter than the clip span: repeat the last frame (grows the stored list
            names.append(names[-1])       # when the order was not reversed, exactly as the reference does)
            starts = names[:-self.dilation[-1]]
        first = np.random.choice(list(range(len(starts))))
        steps = [first] + [first + d for d in self.dilation]
        flip = np.random.choice([0, 1])
        scale = np.random.choice(self.SCALES)
        images, labels = [], []
        for i in steps:
            name = names[i]
            img = Image.open(os.path.join(self.dataroot, "data", video, "origin", name)).convert("RGB")
            seg = Image.open(os.path.join(self.dataroot, "data", video, "mask", name.split(".")[0] + ".png"))
            if self.split == "train":
                if flip:
                    img = img.transpose(Image.FLIP_LEFT_RIGHT)
                    seg = seg.transpose(Image.FLIP_LEFT_RIGHT)
                if self.args.multi_scale and scale != 1.:
                    w, h = img.size
                    size = (int(w * scale), int(h * scale))
                    img = img.resize(size, Image.BILINEAR)
                    seg = seg.resize(size, Image.NEAREST)
            images.append(np.array(img) if self.device_finish else np.float32(np.array(img)) / 255.)
            labels.append(np.array(seg))
        if self.split == "train":
            images, labels = self._pad_and_crop(images, labels)
        if self.device_finish:
            # (padding: image 0 = 0/255 exactly as the float path pads 0.0; mask padding 255 is the remapped ignore value, so
            # the raw value that maps to it is 0)
            return ([torch.from_numpy(np.ascontiguousarray(i)) for i in images],
                    [torch.from_numpy(np.ascontiguousarray(l)) for l in labels])
        return [image_to_tensor(i) for i in images], [labels_to_tensor(l) for l in labels]


class VSPWClipTest(torch.utils.data.Dataset):
    """All frames of one video, in order; item i = (frame i, its labels, the clip_num-1 neighbour frames at +offsets — or at
    -offsets once i + max offset would run past the end —, their labels, frame file name)."""

    def __init__(self, dataroot, video, args, is_train=False):
        self.dataroot, self.video, self.args = dataroot, video, args
        self.dilation = _parse_dilation(args)
        self.names = sorted(os.listdir(os.path.join(dataroot, "data", video, "origin")))
        self.is_train = is_train
        self.subset = [n for k, n in enumerate(self.names) if k % 15 == 0] if is_train else []
        self.mask_dir = "mask_42label" if getattr(args, "lesslabel", False) else "mask"

    def __len__(self):
        return len(self.subset) if self.is_train else len(self.names)

    def _load(self, name):
        img = Image.open(os.path.join(self.dataroot, "data", self.video, "origin", name))
        seg = Image.open(os.path.join(self.dataroot, "data", self.video, self.mask_dir, name.split(".")[0] + ".png"))
        return image_to_tensor(np.float32(np.array(img)) / 255.), labels_to_tensor(seg)

    def __getitem__(self, index):
        name = self.names[index]
        img, lab = self._load(name)
        back = index + self.dilation[-1] >= len(self.names)
        clip_i, clip_l = [], []
        for d in self.dilation:
            ci, cl = self._load(self.names[index - d if back else index + d])
            clip_i.append(ci)
            clip_l.append(cl)
        return img, lab, clip_i, clip_l, name


def dilation_sublists(names, num):
    """dataset2.py:143-151: the frame list split into num + 1 interleaved sub-lists (every (num+1)-th frame)."""
    return [[n for k, n in enumerate(names) if k % (num + 1) == a] for a in range(num + 1)]


def clip_window(length, index, clip_num):
    """[start, end) of the clip_num-frame window around position `index` of a `length`-frame list, clamped at both ends
    (dataset2.py:276-300): clip_num // 2 frames to the left, one fewer to the right when clip_num is even."""
    left = clip_num // 2
    right = left - 1 if clip_num % 2 == 0 else left
    if index - left < 0:
        start, end = 0, min(clip_num, length)
    elif index + right >= length:
        end = length
        start = max(end - clip_num, 0)
    else:
        start = index - left
        end = start + clip_num
    return start, end


class VSPWWindowTest(torch.utils.data.Dataset):
    """`TestDataset_clip` (dataset2.py:154-337): item i = (frame i, its labels, the frames of the clip_num window around it in
    its dilation sub-list, their labels, names).  With ``args.method == 'nonlocal3d'`` the window includes frame i itself and
    `names` is the list of the window's file names (what `test_all` averages over); otherwise frame i is left out of the
    window and `names` is its own file name.  A sub-list shorter than 2 frames yields the frame alone -- and, as in the
    reference, an EMPTY name list in nonlocal3d mode, so `test_all` never scores such a frame."""

    def __init__(self, dataroot, video, args, is_train=False):
        self.dataroot, self.video, self.args = dataroot, video, args
        self.clip_num = int(args.clip_num)
        self.names = sorted(os.listdir(os.path.join(dataroot, "data", video, "origin")))
        self.sublists = dilation_sublists(self.names, int(args.dilation_num))
        self.is_train = is_train
        self.subset = [n for k, n in enumerate(self.names) if k % 15 == 0] if is_train else []
        self.mask_dir = "mask_42label" if getattr(args, "lesslabel", False) else "mask"
        self.whole_clip = args.method == "nonlocal3d"

    def __len__(self):
        return len(self.subset) if self.is_train else len(self.names)

    _load = VSPWClipTest._load

    def __getitem__(self, index):
        name = (self.subset if self.is_train else self.names)[index]
        img, lab = self._load(name)
        sub = next(l for l in reversed(self.sublists) if name in l)  # the reference keeps the LAST sub-list that matches
        pos = sub.index(name)
        start, end = clip_window(len(sub), pos, self.clip_num)
        names = [] if self.whole_clip else name
        clip_i, clip_l = [], []
        if end - start < 2:
            clip_i.append(img)
            clip_l.append(lab)
        else:
            for k in range(start, end):
                if k == pos and not self.whole_clip:
                    continue
                if self.whole_clip:
                    names.append(sub[k])
                ci, cl = self._load(sub[k])
                clip_i.append(ci)
                clip_l.append(cl)
        return img, lab, clip_i, clip_l, names
