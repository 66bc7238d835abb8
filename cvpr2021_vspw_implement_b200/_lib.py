"""ctypes binding of ``csrc/libvspw_b200.so`` (the C ABI declared in ``include/vspw_b200.h``).

The product path has no CPU or library fallback: if the shared library is missing and cannot be
built, importing the engine raises.  Every call passes raw device pointers (``Tensor.data_ptr()``)
and the caller's current CUDA stream.
"""
import ctypes
import os
import threading

from . import build as _build

_c_int = ctypes.c_int32
_c_vp = ctypes.c_void_p
_c_sz = ctypes.c_size_t
_c_f = ctypes.c_float
_c_d = ctypes.c_double
_c_i64 = ctypes.c_int64

PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2


class ConvDesc(ctypes.Structure):
    """Mirror of ``vspw_conv_desc`` (include/vspw_b200.h)."""

    _fields_ = [(k, _c_int) for k in ("n", "h", "w", "cin", "cout", "kh", "kw", "stride", "pad", "dil", "ho", "wo", "precision", "cin_pitch")]


class PeerCtx(ctypes.Structure):
    """Mirror of ``vspw_peer_ctx`` (include/vspw_b200.h)."""

    _fields_ = [("inbox", ctypes.c_uint64 * 16), ("world", _c_int), ("rank", _c_int), ("ring", _c_int), ("max_elems", _c_int),
                ("seq", ctypes.c_uint64)]


# name -> argtypes; every function returns int (0 = ok)
_SIGNATURES = {
    "vspw_conv2d_tc_supported": [ctypes.POINTER(ConvDesc)],
    "vspw_conv2d_wgrad_tc_supported": [ctypes.POINTER(ConvDesc)],
    "vspw_permute4d": [_c_vp, _c_vp, ctypes.POINTER(_c_int * 4), ctypes.POINTER(_c_int * 4), _c_vp],
    "vspw_fill": [_c_vp, _c_f, _c_sz, _c_vp],
    "vspw_axpby": [_c_vp, _c_vp, _c_f, _c_f, _c_sz, _c_vp],
    "vspw_split_bf16": [_c_vp, _c_vp, _c_vp, _c_sz, _c_vp],
    "vspw_zero_insert2_bf16": [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_conv_weight_prep": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_conv_weight_prep_multi": [_c_vp, _c_int, _c_int, _c_int, _c_vp],
    "vspw_clip_finish_u8": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp],
    "vspw_cast_f64_f32": [_c_vp, _c_vp, _c_sz, _c_vp],
    "vspw_copy_channels": [_c_vp, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_int, _c_sz, _c_int, _c_vp],
    "vspw_conv2d_fwd": [ctypes.POINTER(ConvDesc), _c_vp, _c_vp, _c_vp, _c_vp, _c_vp],
    "vspw_conv2d_dgrad": [ctypes.POINTER(ConvDesc), _c_vp, _c_vp, _c_vp, _c_vp],
    "vspw_conv2d_wgrad": [ctypes.POINTER(ConvDesc), _c_vp, _c_vp, _c_vp, _c_vp],
    "vspw_conv2d_fwd_tc": [ctypes.POINTER(ConvDesc), _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp],
    "vspw_conv2d_dgrad_tc": [ctypes.POINTER(ConvDesc), _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp],
    "vspw_conv2d_wgrad_tc": [ctypes.POINTER(ConvDesc), _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp],
    "vspw_bn_stats": [_c_vp, _c_sz, _c_int, _c_vp, _c_vp, _c_vp],
    "vspw_bn_finalize_train": [_c_vp, _c_vp, _c_d, _c_vp, _c_vp, _c_f, _c_f, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp],
    "vspw_bn_fold_eval": [_c_vp, _c_vp, _c_vp, _c_vp, _c_f, _c_vp, _c_vp, _c_vp, _c_int, _c_vp],
    "vspw_bn_act_fwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_sz, _c_vp],
    "vspw_bn_train_fwd": [_c_vp, _c_vp, _c_vp, _c_d, _c_vp, _c_vp, _c_f, _c_f, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int,
                          _c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_sz, _c_vp],
    "vspw_bn_train_fwd_sync": [_c_vp, _c_vp, _c_d, _c_vp, _c_vp, _c_f, _c_f, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int,
                               _c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_sz, ctypes.POINTER(PeerCtx), _c_vp],
    "vspw_bn_bwd_apply_sync": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                               _c_vp, _c_sz, _c_int, _c_sz, _c_d, _c_d, ctypes.POINTER(PeerCtx), _c_vp],
    "vspw_bn_bwd_reduce": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_sz, _c_int, _c_sz, _c_vp, _c_vp, _c_vp],
    "vspw_bn_bwd_apply": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_sz, _c_int, _c_d, _c_d, _c_vp],
    "vspw_maxpool3x3s2_fwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_maxpool3x3s2_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_tcb_pool_fwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_vp],
    "vspw_tcb_pool_bwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_vp],
    "vspw_upsample_bilinear_fwd": [_c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_upsample_bilinear_bwd": [_c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_logsoftmax_up_nll_fwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_logsoftmax_up_nll_bwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_f, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_loss_finalize": [_c_vp, _c_vp, _c_f, _c_vp, _c_vp, _c_vp],
    "vspw_up_softmax_fwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_softmax_strided_fwd": [_c_vp, _c_vp, _c_sz, _c_int, _c_sz, _c_sz, _c_int, _c_sz, _c_f, _c_vp],
    "vspw_softmax_strided_bwd": [_c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_sz, _c_sz, _c_int, _c_sz, _c_f, _c_vp],
    "vspw_bgemm": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int] + [_c_i64] * 9 + [_c_f, _c_f, _c_vp],
    "vspw_bgemm_det": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int] + [_c_i64] * 9 + [_c_f, _c_f, _c_vp],
    "vspw_ocr_attention_fwd_tc": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_f, _c_vp],
    "vspw_ocr_gather_tc": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_ocr_region_softmax_fwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_f, _c_vp],
    "vspw_ocr_region_softmax_bwd": [_c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp],
    "vspw_ocr_attn_softmax_bwd_planes": [_c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_f, _c_vp],
    "vspw_ocr_operand_planes": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_vp],
    "vspw_ocr_region_planes": [_c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_f, _c_vp],
    "vspw_vc_counts": [_c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_vp, _c_vp],
    "vspw_sgd_momentum_step": [_c_vp, _c_vp, _c_vp, _c_int, _c_f, _c_vp],
    "vspw_confusion_add": [_c_vp, _c_vp, _c_vp, _c_sz, _c_int, _c_vp],
    "vspw_ppm_weight_slices": [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_ppm_pyramid_fwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp],
    "vspw_ppm_pyramid_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp],
    "vspw_peer_alloc": [_c_sz, ctypes.POINTER(_c_vp), _c_vp],
    "vspw_peer_open": [_c_vp, ctypes.POINTER(_c_vp)],
    "vspw_peer_close": [_c_vp],
    "vspw_peer_free": [_c_vp],
    "vspw_peer_allreduce_f64": [_c_vp, _c_int, _c_vp, _c_int, _c_int, ctypes.c_uint64, _c_int, _c_int, _c_vp],
}

EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["vspw_last_error", "vspw_version", "vspw_tcb_pool_workspace_floats", "vspw_sgd_chunk_elems",
                                                   "vspw_conv_weight_prep_tile", "vspw_peer_inbox_bytes",
                                                   "vspw_ocr_attention_workspace_bytes", "vspw_ocr_region_softmax_workspace_bytes"])


class VspwError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self._lock = threading.Lock()
        self.launches = 0  # number of C-ABI compute calls issued (bench.py reports it)
        self._prof = None  # live per-entry-point timing: list of (name, start_event, stop_event)

    def dll(self):
        if self._dll is None:
            with self._lock:
                if self._dll is None:
                    # (re)build when the sources are newer than the .so; raises if nvcc fails: no silent fallback
                    # VSPW_LIB_PATH: load this prebuilt variant instead (A/B timing of compile-time options only)
                    path = os.environ.get("VSPW_LIB_PATH") or _build.build_library()
                    dll = ctypes.CDLL(path)
                    for name, argtypes in _SIGNATURES.items():
                        fn = getattr(dll, name)
                        fn.argtypes = argtypes
                        fn.restype = ctypes.c_int
                    dll.vspw_last_error.restype = ctypes.c_char_p
                    dll.vspw_last_error.argtypes = []
                    dll.vspw_version.restype = ctypes.c_int
                    dll.vspw_version.argtypes = []
                    dll.vspw_sgd_chunk_elems.restype = ctypes.c_int32
                    dll.vspw_sgd_chunk_elems.argtypes = []
                    dll.vspw_conv_weight_prep_tile.restype = ctypes.c_int32
                    dll.vspw_conv_weight_prep_tile.argtypes = [_c_int] * 4
                    dll.vspw_ocr_region_softmax_workspace_bytes.restype = ctypes.c_size_t
                    dll.vspw_ocr_region_softmax_workspace_bytes.argtypes = [_c_int, _c_int]
                    dll.vspw_ocr_attention_workspace_bytes.restype = ctypes.c_size_t
                    dll.vspw_ocr_attention_workspace_bytes.argtypes = [_c_int]
                    dll.vspw_peer_inbox_bytes.restype = ctypes.c_size_t
                    dll.vspw_peer_inbox_bytes.argtypes = [_c_int, _c_int, _c_int]
                    dll.vspw_tcb_pool_workspace_floats.restype = ctypes.c_size_t
                    dll.vspw_tcb_pool_workspace_floats.argtypes = [_c_int, _c_int, _c_int, _c_int, _c_vp, _c_int]
                    self._dll = dll
        return self._dll

    def profile_begin(self):
        """Time every C-ABI call with CUDA events on the current stream (inside a real step, at real clocks)."""
        self._prof = []

    def profile_end(self):
        """-> {entry point: (calls, total ms)}; synchronises the device."""
        import torch
        rec, self._prof = self._prof or [], None
        torch.cuda.synchronize()
        out = {}
        self.last_profile_calls = []  # (index in call order, entry point, ms) of every call, for outlier hunting
        for i, (name, e0, e1) in enumerate(rec):
            ms = e0.elapsed_time(e1)
            n, tot = out.get(name, (0, 0.0))
            out[name] = (n + 1, tot + ms)
            self.last_profile_calls.append((i, name, ms))
        return out

    def call(self, name, *args):
        dll = self.dll()
        if self._prof is not None:
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = getattr(dll, name)(*args)
            e1.record()
            self._prof.append((name, e0, e1))
        else:
            rc = getattr(dll, name)(*args)
        if rc != 0:
            msg = dll.vspw_last_error().decode("utf-8", "replace")
            raise VspwError(f"{name} failed ({rc}): {msg}")
        self.launches += 1
        return rc

    def tc_supported(self, desc):
        return bool(self.dll().vspw_conv2d_tc_supported(ctypes.byref(desc)))

    def wgrad_tc_supported(self, desc):
        return bool(self.dll().vspw_conv2d_wgrad_tc_supported(ctypes.byref(desc)))

    def version(self):
        return self.dll().vspw_version()


lib = _Lib()


def i4(*vals):
    return ctypes.byref((_c_int * 4)(*vals))
