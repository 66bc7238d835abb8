"""vspw_b200 — B200-native (sm_100a) engine for the per-clip hot path of CVPR2021_VSPW_Implement.

Public surface = the reference's plugin surface for that path:
``models.ModelBuilder``, ``models.SegmentationModule``, ``models.Clip_PSP``, ``models.ClipOCRNet``.
Arithmetic is done by hand-written CUDA kernels behind the C ABI in ``include/vspw_b200.h``.
"""
from . import engine  # noqa: F401
from .engine import get_precision, precision, set_precision  # noqa: F401

__version__ = "0.1.0"
