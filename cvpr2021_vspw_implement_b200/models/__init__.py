"""Host-side mirror of the reference's ``models`` package for the TCB hot path
(reference models/__init__.py exports ModelBuilder, SegmentationModule and the video wrappers)."""
from .models import ModelBuilder, SegmentationModule, Resnet, ResnetDilated, PPMDeepsup
from .clip_psp import Clip_PSP, PPM_conv
from .clip_ocr import ClipOCRNet
from .non_local import NLBlockND, Non_local3d
