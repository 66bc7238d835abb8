"""Synthetic stand-ins for the VSPW clip datasets, with the reference datasets' OUTPUT CONTRACT
(dataset2.py: BaseDataset_longclip :852-1048 for training, TestDataset_longclip :344-490 for inference).

JPEG/PNG decoding and augmentation are outside the hot path; what the hot path needs
from a loader is the batch layout, which these reproduce exactly so the entry points run end to end without the
VSPW files:

  train item : (clip_imgs, clip_gts) = T tensors (3, H, W) float32 ImageNet-normalised, T tensors (1, H, W) float
               with values in {0..num_class-1, 255}; the default collate turns them into lists of (n, 3, H, W) /
               (n, 1, H, W) batches.  Frame 0 is the "current" frame (train_clip2.py:75-83).
  test item  : (img, gt, clip_imgs, clip_gts, gt_name) per frame of a video (test_clip2.py:33).
"""
import torch
from torch.utils.data import Dataset


def synthetic_frame(gen, h, w, num_class, block=32, ignore_frac=0.05, ignore_index=255):
    img = torch.randn(3, h, w, generator=gen)
    bh, bw = (h + block - 1) // block, (w + block - 1) // block
    tiles = torch.randint(0, num_class, (1, bh, bw), generator=gen).float()
    tiles[torch.rand(1, bh, bw, generator=gen) < ignore_frac] = float(ignore_index)
    lab = tiles.repeat_interleave(block, dim=1).repeat_interleave(block, dim=2)[:, :h, :w].contiguous()
    return img, lab


def synthetic_clip(T, n, H, W, num_class=124, seed=304, block=32, ignore_frac=0.05, ignore_index=255):
    """The benchmark / parity workload of SURVEY.md section 8d: T image tensors (n,3,H,W) ~ N(0,1) and T label tensors
    (n,1,H,W) float, piece-wise constant `block` x `block` tiles of uniform classes with `ignore_frac` of the tiles set to
    ignore_index.  (Same recipe, same bits as the test infrastructure's oracle/tcb_oracle.py::synthetic_clip.)"""
    g = torch.Generator().manual_seed(seed)
    imgs = [torch.randn(n, 3, H, W, generator=g) for _ in range(T)]
    labs = []
    bh, bw = (H + block - 1) // block, (W + block - 1) // block
    for _ in range(T):
        tiles = torch.randint(0, num_class, (n, 1, bh, bw), generator=g).float()
        drop = torch.rand(n, 1, bh, bw, generator=g) < ignore_frac
        tiles[drop] = float(ignore_index)
        labs.append(tiles.repeat_interleave(block, dim=2).repeat_interleave(block, dim=3)[:, :, :H, :W].contiguous())
    return imgs, labs


class SyntheticClipTrain(Dataset):
    """`length` random clips of `clip_num` frames at (height, width); deterministic per index."""

    def __init__(self, args, length=64, height=None, width=None, seed=304):
        self.t = int(args.clip_num)
        self.h = int(height or getattr(args, "cropsize", 480))
        self.w = int(width or getattr(args, "cropsize", 480))
        self.k = int(args.num_class)
        self.length, self.seed = int(length), int(seed)

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        gen = torch.Generator().manual_seed(self.seed * 100003 + i)
        frames = [synthetic_frame(gen, self.h, self.w, self.k) for _ in range(self.t)]
        return [f[0] for f in frames], [f[1] for f in frames]


class SyntheticClipTest(Dataset):
    """One synthetic 'video' of `frames` frames; item i = frame i plus its clip_num-1 neighbour frames."""

    def __init__(self, args, video="synthetic_000", frames=12, height=480, width=854, seed=304):
        self.t, self.k = int(args.clip_num), int(args.num_class)
        gen = torch.Generator().manual_seed(seed + sum(map(ord, video)))
        base, lab = synthetic_frame(gen, height, width, self.k)
        # a static scene seen through per-frame noise: labels are constant over the video so the VC metric is defined
        self.frames = [(base + 0.1 * torch.randn(3, height, width, generator=gen), lab) for _ in range(frames)]
        self.video = video
        self.offsets = [int(x) for x in str(args.dilation2).split(",")] if self.t > 1 else []
        assert len(self.offsets) + 1 == self.t  # dataset2.py:357

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, i):
        img, gt = self.frames[i]
        # neighbour frames as TestDataset_longclip picks them (dataset2.py:467-472): +offsets, or -offsets near the end
        back = i + self.offsets[-1] >= len(self.frames) if self.offsets else False
        nb = [self.frames[max(0, i - o) if back else i + o] for o in self.offsets]
        return img, gt, [f[0] for f in nb], [f[1] for f in nb], f"{i:08d}.png"


class SyntheticWindowTest(SyntheticClipTest):
    """The synthetic video served as `TestDataset_clip` serves a real one (vspw_data.VSPWWindowTest): the clip_num window
    around frame i inside its dilation sub-list, including frame i and with the window's frame names for nonlocal3d."""

    def __init__(self, args, video="synthetic_000", frames=12, height=480, width=854, seed=304):
        t, args_t = int(args.clip_num), args
        super().__init__(_WithOffsets(args_t, t), video, frames, height, width, seed)
        self.whole_clip = args.method == "nonlocal3d"
        self.step = int(args.dilation_num) + 1

    def __getitem__(self, i):
        from .vspw_data import clip_window
        img, gt = self.frames[i]
        sub = list(range(i % self.step, len(self.frames), self.step))
        pos = sub.index(i)
        start, end = clip_window(len(sub), pos, self.t)
        name = f"{i:08d}.png"
        if end - start < 2:
            return img, gt, [img], [gt], ([] if self.whole_clip else name)
        ks = [sub[k] for k in range(start, end) if self.whole_clip or k != pos]
        names = [f"{k:08d}.png" for k in ks] if self.whole_clip else name
        return img, gt, [self.frames[k][0] for k in ks], [self.frames[k][1] for k in ks], names


class _WithOffsets:
    """args view whose dilation2 always has clip_num - 1 entries (the window datasets do not use the offsets)."""

    def __init__(self, args, t):
        self._args, self.dilation2 = args, ",".join(str(k + 1) for k in range(t - 1))

    def __getattr__(self, k):
        return getattr(self._args, k)


_copy_streams = {}


class DevicePrefetcher:
    """Double-buffered host -> device feed for the train loop: the H2D copy of clip i+1 runs on a side stream while clip i
    computes (the reference copies synchronously with `.cuda()` before every step, train_clip2.py:45-47).  `source` yields
    (clip_imgs, clip_gts) lists of pinned host tensors; iteration yields the same lists on `device`, valid until the next
    item is requested."""

    def __init__(self, source, device, finish_u8=False):
        """finish_u8=True: `source` yields uint8 batches — images (n, h, w, 3), raw masks (n, h, w) — as
        `vspw_data.VSPWClipTrain(device_finish=True)` produces them; they are copied as bytes and turned into the model's
        float tensors on the copy stream by vspw_clip_finish_u8."""
        self.it = iter(source)
        self.device = device
        self.finish_u8 = bool(finish_u8)
        # one copy stream per device for the life of the process: the caching allocator keeps a pool per stream, so a fresh
        # stream per prefetcher would cudaMalloc its input buffers again every epoch
        key = torch.device(device)
        if key not in _copy_streams:
            _copy_streams[key] = torch.cuda.Stream(device=key)
        self.stream = _copy_streams[key]
        self.next = None
        self._preload()

    def _preload(self):
        try:
            imgs, gts = next(self.it)
        except StopIteration:
            self.next = None
            return
        with torch.cuda.stream(self.stream):
            d_imgs = [t.to(self.device, non_blocking=True) for t in imgs]
            d_gts = [t.to(self.device, non_blocking=True) for t in gts]
            if self.finish_u8:
                d_imgs, d_gts = self._finish(d_imgs, d_gts)
        self.next = (d_imgs, d_gts)

    def _finish(self, imgs_u8, gts_u8):
        import ctypes
        from ._lib import lib
        st = ctypes.c_void_p(self.stream.cuda_stream)
        out_i, out_l = [], []
        for im, lb in zip(imgs_u8, gts_u8):
            if im.dtype != torch.uint8 or lb.dtype != torch.uint8 or im.dim() != 4 or im.shape[3] != 3:
                raise ValueError("finish_u8 expects uint8 (n, h, w, 3) images and uint8 (n, h, w) masks")
            n, h, w, _ = im.shape
            fi = torch.empty((n, 3, h, w), device=self.device, dtype=torch.float32)
            fl = torch.empty((n, 1, h, w), device=self.device, dtype=torch.float32)
            lib.call("vspw_clip_finish_u8", ctypes.c_void_p(im.data_ptr()), ctypes.c_void_p(lb.data_ptr()), ctypes.c_void_p(fi.data_ptr()),
                     ctypes.c_void_p(fl.data_ptr()), n, h, w, st)
            out_i.append(fi)
            out_l.append(fl)
        self._keep = (imgs_u8, gts_u8)  # the byte buffers stay alive until the next preload (same stream: ordering is implicit)
        return out_i, out_l

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)          # this clip's copies have landed
        imgs, gts = self.next
        for t in imgs + gts:
            t.record_stream(cur)              # the caching allocator must not recycle them while the step still reads them
        self._preload()                       # start copying the next clip behind the step about to be launched
        return imgs, gts
