"""Host-side helpers of the entry points: metrics, meters, logging (reference: utils.py:37-122, 262-302).

``Evaluator`` keeps the reference's method names and formulas (confusion matrix rows = ground truth); the matrix
can be fed either from NumPy arrays like the reference (``add_batch``) or from the on-device histogram produced
by ``vspw_confusion_add`` (``add_confusion``), which avoids moving (n, H, W) predictions to the host per batch.
"""
import logging
import re
import sys

import numpy as np


def get_common(list_, predlist, clip_num, h, w):
    """Video-consistency accuracy VC_n (reference utils.py:37-53): for every window of `clip_num` consecutive frames, the
    share of pixels whose ground truth is constant over the window whose prediction is constant over it too."""
    accs = []
    for i in range(len(list_) - clip_num):
        global_common = np.ones((h, w))
        predglobal_common = np.ones((h, w))
        for j in range(1, clip_num):
            global_common = np.logical_and(global_common, list_[i] == list_[i + j])
            predglobal_common = np.logical_and(predglobal_common, predlist[i] == predlist[i + j])
        pred = predglobal_common * global_common
        accs.append(pred.sum() / global_common.sum())
    return accs


def get_common_device(labels, pred, clip_num):
    """get_common for one video resident on the device: `labels` float and `pred` integer CUDA tensors of
    shape (frames, H, W).  One launch of vspw_vc_counts; only the 2 integers per window cross to the host, and the ratios
    formed from them are the reference's bit for bit (nan where no pixel has a constant label, like NumPy's 0/0)."""
    import ctypes
    import torch
    from ._lib import lib
    assert labels.is_cuda and pred.is_cuda and labels.shape == pred.shape and labels.dim() == 3
    frames = int(labels.shape[0])
    windows = frames - int(clip_num)
    if windows <= 0:
        return []
    labels = labels.contiguous().float()
    pred = pred.contiguous().to(torch.int32)
    counts = torch.empty((windows, 2), device=labels.device, dtype=torch.int64)
    lib.call("vspw_vc_counts", ctypes.c_void_p(labels.data_ptr()), ctypes.c_void_p(pred.data_ptr()), frames,
             labels[0].numel(), int(clip_num), ctypes.c_void_p(counts.data_ptr()),
             ctypes.c_void_p(torch.cuda.current_stream(labels.device).cuda_stream))
    c = counts.cpu().numpy().astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return list(c[:, 0] / c[:, 1])


class Evaluator(object):
    def __init__(self, num_class):
        self.num_class = num_class
        self.confusion_matrix = np.zeros((self.num_class,) * 2)

    def beforeval(self):
        isval = np.sum(self.confusion_matrix, axis=1) > 0
        self.confusion_matrix = self.confusion_matrix * isval

    def Pixel_Accuracy(self):
        return np.diag(self.confusion_matrix).sum() / self.confusion_matrix.sum()

    def Pixel_Accuracy_Class(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(self.confusion_matrix) / self.confusion_matrix.sum(axis=1)
        return np.nanmean(acc)

    def _iu(self):
        cm = self.confusion_matrix
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))

    def Mean_Intersection_over_Union(self):
        isval = np.sum(self.confusion_matrix, axis=1) > 0
        return np.nansum(self._iu() * isval) / isval.sum()

    def Frequency_Weighted_Intersection_over_Union(self):
        freq = np.sum(self.confusion_matrix, axis=1) / np.sum(self.confusion_matrix)
        iu = self._iu()
        return (freq[freq > 0] * iu[freq > 0]).sum()

    def _generate_matrix(self, gt_image, pre_image):
        mask = (gt_image >= 0) & (gt_image < self.num_class)
        label = self.num_class * gt_image[mask].astype("int") + pre_image[mask]
        count = np.bincount(label, minlength=self.num_class ** 2)
        return count.reshape(self.num_class, self.num_class)

    def add_batch(self, gt_image, pre_image):
        assert gt_image.shape == pre_image.shape
        self.confusion_matrix += self._generate_matrix(gt_image, pre_image)

    def add_batch_device(self, gt, pred):
        """Same as add_batch for CUDA tensors: `gt` (n,1,H,W) or (n,H,W) float labels, `pred` (n,H,W) integer argmax.
        The 124x124 histogram is accumulated on the device by vspw_confusion_add; only the matrix
        crosses to the host, when a metric is read (`sync_device`)."""
        import ctypes
        import torch
        from ._lib import lib
        gt = gt.reshape(pred.shape).contiguous().float()
        pred = pred.contiguous().to(torch.int32)
        assert gt.is_cuda and pred.is_cuda and gt.shape == pred.shape
        if getattr(self, "_dev_conf", None) is None or self._dev_conf.device != gt.device:
            self._dev_conf = torch.zeros((self.num_class, self.num_class), device=gt.device, dtype=torch.int64)
        lib.call("vspw_confusion_add", ctypes.c_void_p(pred.data_ptr()), ctypes.c_void_p(gt.data_ptr()),
                 ctypes.c_void_p(self._dev_conf.data_ptr()), gt.numel(), self.num_class,
                 ctypes.c_void_p(torch.cuda.current_stream(gt.device).cuda_stream))

    def sync_device(self):
        """Fold the on-device histogram into confusion_matrix (call before reading metrics)."""
        if getattr(self, "_dev_conf", None) is not None:
            self.confusion_matrix += self._dev_conf.cpu().numpy()
            self._dev_conf.zero_()

    def add_confusion(self, conf):
        """Accumulate a (num_class, num_class) histogram computed on the device (vspw_confusion_add)."""
        conf = np.asarray(conf)
        assert conf.shape == self.confusion_matrix.shape
        self.confusion_matrix += conf

    def reset(self):
        self.confusion_matrix = np.zeros((self.num_class,) * 2)
        if getattr(self, "_dev_conf", None) is not None:
            self._dev_conf.zero_()


class AverageMeter(object):
    """Running average (reference utils.py:262-287)."""

    def __init__(self):
        self.initialized = False
        self.val = self.avg = self.sum = self.count = None

    def initialize(self, val, weight):
        self.val, self.avg, self.sum, self.count, self.initialized = val, val, val * weight, weight, True

    def update(self, val, weight=1):
        if not self.initialized:
            self.initialize(val, weight)
        else:
            self.val = val
            self.sum += val * weight
            self.count += weight
            self.avg = self.sum / self.count

    def value(self):
        return self.val

    def average(self):
        return self.avg


def setup_logger(distributed_rank=0, filename="log.txt"):
    logger = logging.getLogger("Logger")
    logger.setLevel(logging.DEBUG)
    if distributed_rank > 0 or logger.handlers:
        return logger
    ch = logging.StreamHandler(stream=sys.stdout)
    ch.setLevel(logging.DEBUG)
    ch.setFormatter(logging.Formatter("[%(asctime)s %(levelname)s %(filename)s line %(lineno)d %(process)d] %(message)s"))
    logger.addHandler(ch)
    return logger


def parse_devices(input_devices):
    """'0-3' / '0,1,2,3' / 'gpu0,gpu1' -> ['gpu0', ...] (reference utils.py:290-302 semantics)."""
    ret = []
    for d in input_devices.split(","):
        d = d.strip().lower()
        if d == "cpu":
            ret.append("cpu")
            continue
        m = re.fullmatch(r"(?:gpu)?(\d+)(?:-(\d+))?", d)
        if not m:
            raise NotImplementedError(f"Can not parse device: {d}")
        lo = int(m.group(1))
        hi = int(m.group(2)) if m.group(2) is not None else lo
        for i in range(lo, hi + 1):
            if f"gpu{i}" not in ret:
                ret.append(f"gpu{i}")
    return ret
