"""Tape engine: explicit forward/backward of the per-clip hot path over the C-ABI kernels.

Why a tape and not one ``autograd.Function`` per op: every activation buffer, every gradient
accumulation and every kernel launch is ours — no ATen compute kernel runs inside the hot path, the
fan-out sums (residual branches, layer4 -> pooling + decoder) use ``vspw_axpby`` and the whole step
is one node in torch's autograd graph (``run_graph``), so ``loss.backward()`` at the caller
(reference: train_clip2.py:97-104) works unchanged.

Layout: activations are fp32 NHWC ``(N, H, W, C)`` torch tensors (torch = allocator + stream only).
"""
import contextlib
import ctypes
import math
import os
import threading
import weakref
from contextlib import contextmanager

import numpy as np
import torch

from ._lib import ConvDesc, PREC_BF16, PREC_BF16X3, PREC_FP32, VspwError, i4, lib

_PRECISION = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}
_state = {"precision": os.environ.get("VSPW_PRECISION", "bf16x3"), "syncbn_clamp": False, "syncbn": False,
          # VSPW_WGRAD_STREAM=1: weight gradients of the tensor-core convs run on a second stream (nothing downstream in
          # the backward chain reads them) so that the tensor-bound wgrad kernels overlap the HBM-bound BN backward kernels.
          # Measured gain on B200: 1.2 % of the step (the persistent conv CTAs already fill every SM), and per-kernel event
          # times stop being meaningful under overlap, so it is off by default.
          "wgrad_stream": os.environ.get("VSPW_WGRAD_STREAM", "0") == "1",
          # VSPW_WPREP_MULTI=0: rebuild the conv weights' bf16 operand planes with one launch per weight instead of the
          # per-step batched launch (_WeightPrepPlan); for A/B timing only.
          "wprep_multi": os.environ.get("VSPW_WPREP_MULTI", "1") != "0",
          # VSPW_WGRAD_SINGLE=1 (reported option, NOT the parity mode): weight gradients with single-pass bf16 operands (1 MMA
          # per product instead of 3) while forward and dgrad stay bf16x3.  A weight gradient is a sum over ~64 200 pixels of
          # random-signed products, so operand rounding of 2^-9 leaves it ~2e-3 rel-L2 off — below the 3e-3..1e-2 floor that two
          # fp32 implementations of the reference show against each other on these gradients (oracle/NOISE_FLOOR.md), but far
          # above the 1e-4 the kernels are pinned at in parity mode, so it is opt-in.
          "wgrad_single": os.environ.get("VSPW_WGRAD_SINGLE", "0") == "1"}
_side_streams = {}


def _side_stream(device):
    s = _side_streams.get(device)
    if s is None:
        s = _side_streams[device] = torch.cuda.Stream(device=device)
    return s



def set_precision(mode):
    """'fp32' (CUDA-core FFMA), 'bf16x3' (tcgen05, 3-MMA split, parity mode) or 'bf16' (tcgen05 fast mode)."""
    if mode not in _PRECISION:
        raise ValueError(f"unknown precision {mode!r}; expected one of {sorted(_PRECISION)}")
    _state["precision"] = mode


def get_precision():
    return _state["precision"]


@contextmanager
def precision(mode):
    old = _state["precision"]
    set_precision(mode)
    try:
        yield
    finally:
        _state["precision"] = old


_capture = None
_prof = None  # live conv-kernel profile: list of (start_event, stop_event, flops, used_tc)
_prof_on = True


def conv_profile_begin():
    """Start timing every implicit-GEMM conv launch (fwd/dgrad/wgrad) with CUDA events on the launching stream."""
    global _prof, _prof_on
    _prof = []
    _prof_on = True


def conv_profile_sample(on):
    """Between conv_profile_begin() and conv_profile_end(): record (True) or skip (False) the launches that follow — bench.py
    samples every 4th timed step so that the event pairs do not weigh on the step they measure."""
    global _prof_on
    _prof_on = bool(on)


def conv_profile_end():
    """Stop; returns {'launches', 'ms', 'tflop', 'kernel'} summed over the profiled launches (synchronises)."""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    ms = sum(r[0].elapsed_time(r[1]) for r in rec)
    tc = sum(1 for r in rec if r[3])
    kern = ("conv_tc2_kernel / conv_tc_kernel / wgrad_tc2_kernel (tcgen05 implicit-GEMM family)" if tc * 2 > len(rec)
            else "igemm_kernel/wgrad_kernel (fp32 FFMA implicit GEMM)")
    return {"launches": len(rec), "ms": ms, "tflop": sum(r[2] for r in rec) / 1e12, "kernel": kern, "tc_launches": tc,
            # the launches that stayed on the CUDA-core arm, with their geometry: what is left to move
            "fp32_arm": [(r[4], r[0].elapsed_time(r[1])) for r in rec if not r[3]],
            # (tag, ms, flops) of the tensor-core launches that carry a geometry tag: the per-layer table of bench.py --kernel-profile
            "tc_arm": [(r[4], r[0].elapsed_time(r[1]), r[2]) for r in rec if r[3] and r[4]]}


class _ConvTimer:
    __slots__ = ("e0", "flops", "tc", "tag")

    def __init__(self, flops, tc, tag=None):
        self.flops, self.tc, self.tag = flops, tc, tag
        self.e0 = None

    def __enter__(self):
        if _prof is not None and _prof_on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.e0, e1, self.flops, self.tc, self.tag))
        return False


class TorchDistGroup:
    """SyncBN statistics exchange through torch.distributed's default group (NCCL on GPUs)."""

    def size(self):
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def all_reduce_sums(self, t):
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def set_syncbn(on, clamp=None, group=None):
    """Cross-rank BN statistics (reference multi-GPU semantics, sync_batchnorm/batchnorm.py:110-150): the per-channel
    sum / sum-of-squares (and the two backward sums) are summed over the ranks of `group` -- any object with
    ``size()`` and ``all_reduce_sums(fp64 tensor)`` (in place, on the current stream): `parallel.PeerSums` (one-shot
    exchange over NVLink peer memory, the default of the entry points), `TorchDistGroup` (torch.distributed's default
    group; used when none is given), or a test double."""
    _state["syncbn"] = bool(on)
    _state["syncbn_group"] = group if on else None
    if clamp is not None:
        _state["syncbn_clamp"] = bool(clamp)


def _syncbn_group():
    if not _state.get("syncbn"):
        return None
    g = _state.get("syncbn_group")
    if g is None:
        g = _state["syncbn_group"] = TorchDistGroup()
    return g if g.size() > 1 else None


@contextmanager
def capturing():
    """Test hook: collect named intermediate tensors (NHWC) that the graphs publish via ``publish``."""
    global _capture
    old, _capture = _capture, {}
    try:
        yield _capture
    finally:
        _capture = old


def publish(name, var):
    if _capture is not None:
        _capture[name] = var.data


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _require_cuda(t, what):
    if not t.is_cuda:
        raise VspwError(f"{what}: tensor must live on a CUDA device (the engine has no CPU path)")


class Var:
    """An activation (NHWC fp32) with an optional gradient slot and cached bf16 planes."""

    __slots__ = ("data", "grad", "needs_grad", "planes", "stats", "grad_planes", "wants_grad_planes", "wants_grad_fp32")

    def __init__(self, data, needs_grad=False):
        self.data = data          # fp32 NHWC tensor; None when only the bf16 planes were materialised
        self.grad = None
        self.needs_grad = needs_grad
        self.planes = None
        self.grad_planes = None         # bf16 (hi, lo) planes of the COMPLETE gradient, written by the consumer's backward
        self.wants_grad_planes = False  # set by a tcgen05 conv on its output: its backward reads planes
        self.wants_grad_fp32 = True     # ... and whether it needs the fp32 gradient as well (bias gradient)
        self.stats = None  # (2, C) fp64 per-channel [sum, sum of squares] when the producing conv's epilogue computed them

    @property
    def shape(self):
        return self.data.shape if self.data is not None else self.planes[0].shape

    def add_grad(self, g):
        """Accumulate ``g`` (takes ownership when the slot is empty)."""
        if self.grad is None:
            self.grad = g
        else:
            lib.call("vspw_axpby", _p(g), _p(self.grad), 1.0, 1.0, g.numel(), _stream())

    def add_grad_rows(self, g, n0, n1):
        """Accumulate into images [n0, n1) of the gradient (backward of a batch-axis slice)."""
        if self.grad is None:
            self.grad = torch.empty_like(self.data)
            lib.call("vspw_fill", _p(self.grad), 0.0, self.grad.numel(), _stream())
        dst = self.grad[n0:n1]
        lib.call("vspw_axpby", _p(g), _p(dst), 1.0, 1.0, g.numel(), _stream())


_grad_sink = None  # object with destination(param) -> tensor | None, mark(param), late(param), node_done(): parallel.GradBucket


def set_grad_sink(sink):
    """Register where parameter gradients go: `sink.destination(param)` returns the fp32 tensor (the parameter's slice of
    a flat gradient bucket, pre-zeroed, aliased by ``param.grad``) that the backward kernels write the FIRST contribution
    into directly; later contributions of the same tape are accumulated into it.  None = hand the gradients to autograd."""
    global _grad_sink
    _grad_sink = sink


class PVar:
    """A parameter seen by the tape: reference-layout data plus per-step derived layouts."""

    __slots__ = ("param", "grad", "cache", "sunk")

    def __init__(self, param):
        self.param = param
        self.grad = None
        self.cache = {}
        self.sunk = False  # the gradient lives in the sink's tensor (== param.grad's storage): autograd gets None for it

    def first_dst(self):
        """Destination for a kernel that OVERWRITES its output with this parameter's first gradient contribution: the sink's
        slice (contiguous, parameter-shaped) or None (then the caller allocates and calls add_grad)."""
        if self.grad is not None or _grad_sink is None:
            return None
        t = _grad_sink.destination(self.param)
        if t is None or not t.is_contiguous() or t.dtype != torch.float32:
            return None
        self.grad = t
        self.sunk = True
        return t

    @property
    def data(self):
        d = self.param.data
        # nn.Conv3d(kernel_size=1) weights (NLBlockND, non_local.py:51-72) are 1x1 convs over the (t, h, w) positions
        return d.view(d.shape[0], d.shape[1], 1, 1) if d.dim() == 5 and tuple(d.shape[2:]) == (1, 1, 1) else d

    @property
    def needs_grad(self):
        return self.param.requires_grad

    def add_grad(self, g):
        if self.grad is None:
            self.grad = g
        else:
            if self.sunk and _grad_sink is not None:
                _grad_sink.late(self.param)  # a weight used twice: its chunk of the bucket is not complete yet
            lib.call("vspw_axpby", _p(g), _p(self.grad), 1.0, 1.0, g.numel(), _stream())


_ARENA_BYTES = 8 << 20
_arenas = {}  # device -> [fp64 buffer, owning tape or None]: per-step pool of zeroed accumulators


class Tape:
    def __init__(self, grad_enabled):
        self.grad_enabled = grad_enabled
        self._nodes = []
        self._params = {}
        self._arena = None
        self._arena_off = 0
        self._done = False  # set once backward has run (or the tape was released): its arena slices are dead
        self._side = None     # side stream used by this tape's backward (joined at the end of backward)
        self._keepalive = []  # tensors read by side-stream kernels: not returned to the allocator before the join
        self._wprep_done = set()  # weight-prep plans this tape has already refreshed

    def zeros_f64(self, shape, device):
        """Zero-initialised fp64 accumulator (BN sums, bias gradients).  One memset per step instead of one fill launch
        per BN layer: slices of a per-device arena that this tape zeroes once, the first time it asks."""
        n = 1
        for d in shape:
            n *= int(d)
        n_al = (n + 15) // 16 * 16
        if self._arena is None:
            with _host_lock:
                slot = _arenas.get(device)
                if slot is None:
                    slot = _arenas[device] = [torch.empty(_ARENA_BYTES // 8, device=device, dtype=torch.float64), None]
                owner = slot[1]() if slot[1] is not None else None
                if owner is None or owner._done:
                    slot[1] = weakref.ref(self)
                    self._arena = slot[0]
                    # fp64 zeros are zero words: one fill of the whole arena per step
                    lib.call("vspw_fill", _p(self._arena), 0.0, self._arena.numel() * 2, _stream())
                else:
                    self._arena = False  # another tape is still between its forward and backward on this device
        if self._arena is False or self._arena_off + n_al > self._arena.numel():
            return torch.zeros(shape, device=device, dtype=torch.float64)
        out = self._arena[self._arena_off:self._arena_off + n].view(shape)
        self._arena_off += n_al
        return out

    def param(self, p):
        if p is None:
            return None
        v = self._params.get(id(p))
        if v is None:
            v = self._params[id(p)] = PVar(p)
        return v

    def record(self, fn):
        if self.grad_enabled:
            self._nodes.append(fn)

    def backward(self):
        for fn in reversed(self._nodes):
            fn()
            # every kernel of this node is enqueued: the sink may start reducing the gradient chunks it completed (not while
            # weight gradients run on the side stream — then the sink reduces at the end of the step)
            if _grad_sink is not None and self._side is None:
                _grad_sink.node_done()
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side = None
        self._keepalive = []
        self._nodes = []
        self._done = True

    def release(self):
        self._nodes = []
        self._done = True


# ------------------------------------------------------------------------------------------------
# layout helpers
def permute4d(src, dst, dims, perm):
    lib.call("vspw_permute4d", _p(src), _p(dst), i4(*dims), i4(*perm), _stream())


def nchw_to_nhwc_into(src_nchw, dst_nhwc):
    n, c, h, w = src_nchw.shape
    permute4d(src_nchw, dst_nhwc, (n, c, h, w), (0, 2, 3, 1))


def nhwc_to_nchw(x):
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    permute4d(x, out, (n, h, w, c), (0, 3, 1, 2))
    return out


def input_from_frames(frames):
    """cat(frames, dim=0) + NCHW->NHWC in one pass per frame (clip_psp.py:142-143)."""
    n, c, h, w = frames[0].shape
    out = torch.empty((n * len(frames), h, w, c), device=frames[0].device, dtype=torch.float32)
    for t, f in enumerate(frames):
        _require_cuda(f, "input frame")
        if f.shape != frames[0].shape:
            raise VspwError("all frames of a clip must have the same shape")
        f = f.contiguous() if not f.is_contiguous() else f
        if f.dtype != torch.float32:
            f = f.float()
        nchw_to_nhwc_into(f, out[t * n:(t + 1) * n])
    return out


def _weight_ohwi(tape, wv):
    """OIHW -> OHWI ([Cout][kh][kw][Cin], K contiguous); identity for 1x1."""
    w = wv.data
    co, ci, kh, kw = w.shape
    if kh == 1 and kw == 1:
        return w
    t = wv.cache.get("ohwi")
    if t is None:
        t = torch.empty((co, kh, kw, ci), device=w.device, dtype=torch.float32)
        permute4d(w, t, (co, ci, kh, kw), (0, 2, 3, 1))
        wv.cache["ohwi"] = t
    return t


def _weight_ihwo(tape, wv):
    """OIHW -> [Cin][kh][kw][Cout] (rows = Cin, K = taps*Cout contiguous) for dgrad."""
    w = wv.data
    co, ci, kh, kw = w.shape
    t = wv.cache.get("ihwo")
    if t is None:
        t = torch.empty((ci, kh, kw, co), device=w.device, dtype=torch.float32)
        # (co, ci*kh*kw) -> (ci*kh*kw, co): tiled transpose fast path
        permute4d(w, t, (1, co, ci * kh * kw, 1), (0, 2, 3, 1))
        wv.cache["ihwo"] = t
    return t


def _planes_of(t):
    """(hi, lo) bf16 planes of an fp32 tensor (operands of the tcgen05 convs)."""
    hi = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    lo = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16) if _state["precision"] == "bf16x3" else None
    lib.call("vspw_split_bf16", _p(t), _p(hi), _p(lo), t.numel(), _stream())
    return hi, lo


def _var_planes(v):
    if v.planes is None:
        v.planes = _planes_of(v.data)
    return v.planes


def _tc_weight_planes_once(w, x3):
    """The four operand planes of one OIHW weight by the single-tensor kernel."""
    co, ci, kh, kw = w.shape
    mk = lambda shape: torch.empty(shape, device=w.device, dtype=torch.bfloat16)
    oh, ih = mk((co, kh, kw, ci)), mk((ci, kh, kw, co))
    ol, il = (mk((co, kh, kw, ci)), mk((ci, kh, kw, co))) if x3 else (None, None)
    lib.call("vspw_conv_weight_prep", _p(w), _p(oh), _p(ol), _p(ih), _p(il), co, ci, kh, kw, _stream())
    return {"ohwi": (oh, ol), "ihwo": (ih, il)}


_WPREP_ENTRY = np.dtype([("w", "<u8"), ("oh", "<u8"), ("ol", "<u8"), ("ih", "<u8"), ("il", "<u8"), ("cout", "<i4"), ("cin", "<i4"),
                         ("taps", "<i4"), ("block0", "<i4"), ("blocks_x", "<i4"), ("reserved", "<i4")])  # == struct vspw_wprep_tensor
assert _WPREP_ENTRY.itemsize == 64


class _WeightPrepPlan:
    """The conv weights the tcgen05 path has used on one device in one precision, their (persistent) bf16 operand planes and
    the device table that rebuilds ALL of them with one launch per tile kind when a new tape first asks for a weight: the
    weights change every optimizer step, and ~110 single-tensor launches of 9-15 us are latency, not bytes.

    The first tape that meets a weight registers it (single-tensor launch, as before); a parameter whose storage moved or
    that died is dropped and re-registered on its next use.  Planes are shared between tapes: a tape that is still waiting
    for its backward sees the planes of the CURRENT weights, which is only wrong for a program that updates weights between
    a forward and its backward -- torch.autograd rejects that program too (in-place version check)."""

    def __init__(self, device, x3):
        self.device, self.x3 = device, x3
        self.entries = {}   # id(param) -> [weakref(param), data_ptr, shape, planes, tile]
        self.tables = None  # [(tile, device table, n_tensors, n_blocks)]; None = rebuild before the next batched launch

    def refresh(self):
        """Rebuild the planes of every registered weight from its current values."""
        for k in [k for k, e in self.entries.items() if not self._current(e)]:
            del self.entries[k]
            self.tables = None
        if not self.entries:
            return
        if self.tables is None:
            self._build_tables()
        for tile, tab, n_t, n_b in self.tables:
            lib.call("vspw_conv_weight_prep_multi", _p(tab), n_t, n_b, tile, _stream())

    @staticmethod
    def _view(param):
        d = param.data
        return d.view(d.shape[0], d.shape[1], 1, 1) if d.dim() == 5 and tuple(d.shape[2:]) == (1, 1, 1) else d

    def _current(self, e):
        p = e[0]()
        if p is None:
            return False
        d = self._view(p)
        return d.data_ptr() == e[1] and tuple(d.shape) == e[2] and d.is_contiguous() and d.device == self.device

    def _build_tables(self):
        self.tables = []
        for tile in (64, 32):
            ents = [e for e in self.entries.values() if e[4] == tile]
            if not ents:
                continue
            table = np.zeros(len(ents), dtype=_WPREP_ENTRY)
            block0 = 0
            for i, e in enumerate(ents):
                co, ci, kh, kw = e[2]
                (oh, ol), (ih, il) = e[3]["ohwi"], e[3]["ihwo"]
                bx, by = (ci + tile - 1) // tile, (co + tile - 1) // tile
                table[i] = (e[1], oh.data_ptr(), ol.data_ptr() if ol is not None else 0, ih.data_ptr(),
                            il.data_ptr() if il is not None else 0, co, ci, kh * kw, block0, bx, 0)
                block0 += bx * by
            dev = torch.from_numpy(table.view(np.uint8).copy()).to(self.device)  # pageable copy: synchronous, only on a rebuild
            self.tables.append((tile, dev, len(ents), block0))

    def planes(self, tape, wv):
        if id(self) not in tape._wprep_done:  # this tape's first weight: bring every known weight up to date at once
            tape._wprep_done.add(id(self))
            self.refresh()
        e = self.entries.get(id(wv.param))
        if e is None or e[0]() is not wv.param or not self._current(e):
            w = wv.data
            tile = int(lib.dll().vspw_conv_weight_prep_tile(*[int(d) for d in w.shape]))
            e = self.entries[id(wv.param)] = [weakref.ref(wv.param), w.data_ptr(), tuple(w.shape), _tc_weight_planes_once(w, self.x3), tile]
            self.tables = None
        return e[3]


_wprep_plans = {}  # (device, precision) -> _WeightPrepPlan
_host_lock = threading.RLock()  # host-side registries (weight-prep plans, accumulator arenas) are shared by the threads of a process


def _tc_weight_planes(tape, wv):
    """{'ohwi': (hi, lo), 'ihwo': (hi, lo)} bf16 operand planes of a conv weight, rebuilt from the OIHW parameter once per
    step: by the plan's batched launch (contiguous parameters), else by one single-tensor launch cached on the tape's PVar."""
    prec = _state["precision"]
    pk = "tc_planes_" + prec
    pl = wv.cache.get(pk)
    if pl is None:
        w = wv.data
        if w.is_contiguous() and _state["wprep_multi"]:
            with _host_lock:
                plan = _wprep_plans.get((w.device, prec))
                if plan is None:
                    plan = _wprep_plans[(w.device, prec)] = _WeightPrepPlan(w.device, prec == "bf16x3")
                pl = plan.planes(tape, wv)
        else:
            pl = _tc_weight_planes_once(w if w.is_contiguous() else w.contiguous(), prec == "bf16x3")
        wv.cache[pk] = pl
    return pl


# ------------------------------------------------------------------------------------------------
def conv2d(tape, x, weight, bias=None, stride=1, pad=0, dil=1, want_stats=False):
    """nn.Conv2d forward + recorded dgrad/wgrad (reference: models/resnet.py:61-66 etc.).
    want_stats: the output feeds a train-mode BN; the tcgen05 epilogue then also emits the per-channel sums."""
    wv = tape.param(weight)
    bv = tape.param(bias)
    n, h, w, cin = x.shape
    dev = x.data.device if x.data is not None else x.planes[0].device
    co, ci, kh, kw = wv.data.shape
    if ci != cin:
        raise VspwError(f"conv2d: input has {cin} channels, weight expects {ci}")
    ho = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    prec = _PRECISION[_state["precision"]]
    desc = ConvDesc(n, h, w, cin, co, kh, kw, stride, pad, dil, ho, wo, prec)
    use_tc = prec != PREC_FP32 and lib.tc_supported(desc)
    if not use_tc:
        desc.precision = PREC_FP32
    if not use_tc and x.data is None:
        raise VspwError("conv2d: the input was materialised as bf16 planes only but this geometry runs on the fp32 arm")
    y = torch.empty((n, ho, wo, co), device=dev, dtype=torch.float32)
    flops = 2.0 * n * ho * wo * co * kh * kw * ci
    geom = f"{n}x{h}x{w} {kh}x{kw} {ci}->{co} s{stride} d{dil}"
    stats = None
    if use_tc:
        xh, xl = _var_planes(x)
        wh, wl = _tc_weight_planes(tape, wv)["ohwi"]
        if want_stats and co % 64 == 0:
            stats = tape.zeros_f64((2, co), y.device)
        with _ConvTimer(flops, True, "fwd " + geom):
            lib.call("vspw_conv2d_fwd_tc", ctypes.byref(desc), _p(xh), _p(xl), _p(wh), _p(wl), _p(bv.data if bv else None), _p(y),
                     _p(stats[0]) if stats is not None else None, _p(stats[1]) if stats is not None else None, _stream())
    else:
        w_ohwi = _weight_ohwi(tape, wv)
        with _ConvTimer(flops, False, "fwd " + geom):
            lib.call("vspw_conv2d_fwd", ctypes.byref(desc), _p(x.data), _p(w_ohwi), _p(bv.data if bv else None), _p(y), _stream())
    out = Var(y, needs_grad=tape.grad_enabled and (x.needs_grad or wv.needs_grad))
    out.stats = stats
    wgrad_tc_ok = use_tc and lib.wgrad_tc_supported(desc)
    # a classifier head (1x1, 124 classes): forward on the conv kernel with TMA zero-fill / clipping of the missing 4 classes;
    # backward on 128-pitch operand planes of the logit gradient — wgrad on the weight-gradient kernel, dgrad as a 1x1 GEMM
    head = use_tc and co % 64 != 0
    if use_tc and not head:
        out.wants_grad_planes = True
        out.wants_grad_fp32 = (bv is not None and bv.needs_grad) or (wv.needs_grad and not wgrad_tc_ok)

    def head_backward(dy):
        st = _stream()
        x3 = prec == PREC_BF16X3
        pixels = n * h * w
        ph = torch.empty((pixels, 128), device=dev, dtype=torch.bfloat16)
        pl = torch.empty((pixels, 128), device=dev, dtype=torch.bfloat16) if x3 else None
        lib.call("vspw_ocr_region_planes", _p(dy), _p(ph), _p(pl), pixels, co, 1.0, st)
        if wv.needs_grad:
            xh, xl = _var_planes(x)
            dst = wv.first_dst()
            dw = dst.view(co, ci) if dst is not None else torch.empty((co, ci), device=dev, dtype=torch.float32)
            with _ConvTimer(flops, True):
                lib.call("vspw_ocr_gather_tc", _p(ph), _p(pl), _p(xh), _p(xl if x3 else None), _p(dw), n, 1, h * w, co, ci, st)
            if dst is None:
                wv.add_grad(dw.view(wv.data.shape))
        if bv is not None and bv.needs_grad:
            sums = tape.zeros_f64((co,), dev)
            lib.call("vspw_bn_stats", _p(dy), pixels, co, _p(sums), None, st)
            bdst = bv.first_dst()
            if bdst is not None:
                _double_to_float(sums, out=bdst)
            else:
                bv.add_grad(_double_to_float(sums))
        if x.needs_grad:
            wt_h, wt_l = _operand_planes(wv.data.view(1, co, ci), 1, co, ci, 128, True, 1.0, x3, st)  # [1][ci][128] = W^T, classes padded
            dx = torch.empty((n, h, w, cin), device=dev, dtype=torch.float32)
            with _ConvTimer(flops, True):
                _conv1x1_tc_raw(ph, pl, wt_h[0], wt_l[0] if x3 else None, dx, pixels, 128, ci, st)
            x.add_grad(dx)

    def backward():
        dy, dyp = out.grad, out.grad_planes
        out.grad = out.grad_planes = None
        if dy is None and dyp is None:
            return
        if dy is not None and dyp is not None and not getattr(out, "wants_grad_fp32", False):
            # planes = the COMPLETE gradient written by the one consumer (a BN backward); a second consumer's fp32 share
            # deposited afterwards would be dropped by the tensor-core kernels, which read the planes only
            raise VspwError("conv2d backward: the output received both gradient planes and an fp32 gradient — a conv output read by "
                            "a BN and by a second consumer must not take the planes-only path")
        if head:
            return head_backward(dy)
        st = _stream()
        wgrad_tc = wgrad_tc_ok and wv.needs_grad
        if dyp is None and use_tc and (x.needs_grad or wgrad_tc):
            dyp = _planes_of(dy)
        if wv.needs_grad:
            side = None
            if wgrad_tc and _state["wgrad_stream"]:
                xh, xl = _var_planes(x)  # (materialised on the main stream if they were not yet)
                side = tape._side = _side_stream(dev)
                side.wait_stream(torch.cuda.current_stream())
                tape._keepalive.append((dyp, xh, xl))
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                sw = _stream()
                dst = wv.first_dst()  # the parameter's slice of the gradient bucket (OIHW), or None
                one = kh == 1 and kw == 1
                dw = dst.view(co, kh, kw, ci) if (dst is not None and one) else torch.empty((co, kh, kw, ci), device=dev, dtype=torch.float32)
                if wgrad_tc:
                    xh, xl = _var_planes(x)
                    wdesc, w_xl, w_dl = desc, xl, dyp[1]
                    if _state["wgrad_single"] and prec == PREC_BF16X3:
                        wdesc = ConvDesc(n, h, w, cin, co, kh, kw, stride, pad, dil, ho, wo, PREC_BF16)
                        w_xl = w_dl = None
                    with _ConvTimer(flops, True, "wgrad " + geom):
                        lib.call("vspw_conv2d_wgrad_tc", ctypes.byref(wdesc), _p(xh), _p(w_xl), _p(dyp[0]), _p(w_dl), _p(dw), sw)
                else:
                    with _ConvTimer(flops, False, "wgrad " + geom):
                        lib.call("vspw_conv2d_wgrad", ctypes.byref(desc), _p(x.data), _p(dy), _p(dw), sw)
                if one:
                    if dst is None:
                        wv.add_grad(dw.view(wv.data.shape))
                else:
                    dw_oihw = dst.view(co, ci, kh, kw) if dst is not None else torch.empty((co, ci, kh, kw), device=dev, dtype=torch.float32)
                    permute4d(dw, dw_oihw, (co, kh, kw, ci), (0, 3, 1, 2))
                    if dst is None:
                        wv.add_grad(dw_oihw)
        if bv is not None and bv.needs_grad:
            sums = tape.zeros_f64((co,), dev)
            lib.call("vspw_bn_stats", _p(dy), n * ho * wo, co, _p(sums), None, st)
            bdst = bv.first_dst()
            if bdst is not None:
                _double_to_float(sums, out=bdst)
            else:
                bv.add_grad(_double_to_float(sums))
        if x.needs_grad:
            # gradient fan-in: when another consumer of x already deposited its share, the tcgen05 epilogue adds into it
            fan_in = use_tc and x.grad is not None and x.grad.is_contiguous() and tuple(x.grad.shape) == (n, h, w, cin)
            dx = x.grad if fan_in else torch.empty((n, h, w, cin), device=dev, dtype=torch.float32)
            if use_tc:
                th, tl = _tc_weight_planes(tape, wv)["ihwo"]
                d1, g_hi, g_lo = desc, dyp[0], dyp[1]
                if stride == 2:
                    # dgrad of a stride-2 conv = stride-1 dgrad of dy laid on the input grid with zeros in between
                    d1 = ConvDesc(n, h, w, cin, co, kh, kw, 1, pad, dil, h, w, prec)
                    g_hi = torch.empty((n, h, w, co), device=dev, dtype=torch.bfloat16)
                    lib.call("vspw_zero_insert2_bf16", _p(dyp[0]), _p(g_hi), n, ho, wo, co, h, w, st)
                    if dyp[1] is not None:
                        g_lo = torch.empty((n, h, w, co), device=dev, dtype=torch.bfloat16)
                        lib.call("vspw_zero_insert2_bf16", _p(dyp[1]), _p(g_lo), n, ho, wo, co, h, w, st)
                with _ConvTimer(flops, True, "dgrad " + geom):
                    lib.call("vspw_conv2d_dgrad_tc", ctypes.byref(d1), _p(g_hi), _p(g_lo), _p(th), _p(tl), _p(dx), 1 if fan_in else 0, st)
            else:
                w_t = _weight_ihwo(tape, wv)
                with _ConvTimer(flops, False, "dgrad " + geom):
                    lib.call("vspw_conv2d_dgrad", ctypes.byref(desc), _p(dy), _p(w_t), _p(dx), st)
            if not fan_in:
                x.add_grad(dx)

    tape.record(backward)
    return out


def conv_will_use_tc(x_shape, weight_shape, stride, pad, dil):
    """True when conv2d would take the tcgen05 path for an NHWC input of `x_shape` in the current precision mode
    (lets a producer skip materialising the fp32 copy of an activation that only tensor-core convs read)."""
    prec = _PRECISION[_state["precision"]]
    if prec == PREC_FP32:
        return False
    n, h, w, cin = x_shape
    co, ci, kh, kw = weight_shape
    ho = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    return bool(lib.tc_supported(ConvDesc(n, h, w, cin, co, kh, kw, stride, pad, dil, ho, wo, prec)))


def _double_to_float(d, out=None):
    """fp64 -> fp32 of a tiny per-channel vector (no ATen compute kernel in the path)."""
    f = out if out is not None else torch.empty(d.shape, device=d.device, dtype=torch.float32)
    lib.call("vspw_cast_f64_f32", _p(d), _p(f), d.numel(), _stream())
    return f


_RELU_BITS_MIN_ELEMS = int(os.environ.get("VSPW_RELU_BITS_MIN", str(24 << 20)))


def set_relu(on):
    """Test switch: False turns every ReLU of the graphs into the identity.  A ReLU-free network has no mask flips, so its
    whole-model gradients can be compared with the reference at kernel-level tolerances (tests/golden/*_norelu.npz,
    oracle/NOISE_FLOOR.md) instead of the ~sqrt(eps) floor a ReLU network has."""
    _state["relu"] = bool(on)


def batchnorm_act(tape, y, bn, relu=True, residual=None, chan_scale=None, training=None, fp32_out=True, planes_out=True):
    """BN (train: batch stats, eval: running stats) [+ residual] [+ ReLU] [* Dropout2d mask].

    Reference: SynchronizedBatchNorm2d.forward (sync_batchnorm/batchnorm.py:68-73) followed by
    nn.ReLU / `out += residual` (resnet.py:72-92) / nn.Dropout2d (clip_psp.py:39).
    fp32_out=False: the caller guarantees every consumer is a tcgen05 conv (or the residual add of the next block, which
    then reads the planes), so only the bf16 planes are written (and the backward ReLU mask is read from the hi plane);
    ignored when no planes are produced.  planes_out=False: no tensor-core conv reads this output (a downsample branch).
    `residual` may be a Var that exists as planes only.
    """
    n, h, w, c = y.shape
    pixels = n * h * w
    relu = relu and _state.get("relu", True)  # (set_relu(False): ReLU-free test graphs)
    training = bn.training if training is None else training
    gv, bv = tape.param(bn.weight), tape.param(bn.bias)
    dev = y.data.device
    st = _stream()
    scale = torch.empty(c, device=dev, dtype=torch.float32)
    shift = torch.empty(c, device=dev, dtype=torch.float32)
    mean = invstd = None
    if training:
        if pixels <= 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {[n, c, h, w]}")
        sums = y.stats
        if sums is None:
            sums = tape.zeros_f64((2, c), dev)
            lib.call("vspw_bn_stats", _p(y.data), pixels, c, _p(sums[0]), _p(sums[1]), st)
        group = _syncbn_group()
        world = group.size() if group is not None else 1
        # the statistics exchange: inside the BN kernel's prologue when the group offers it (parallel.PeerSums), else a call of
        # its own (torch.distributed all-reduce, test doubles)
        peer_ctx = group.next_ctx(2 * c) if (group is not None and hasattr(group, "next_ctx") and sums.is_contiguous()) else None
        if group is not None and peer_ctx is None:
            group.all_reduce_sums(sums)
        count = float(pixels * world)
        mean = torch.empty(c, device=dev, dtype=torch.float32)
        invstd = torch.empty(c, device=dev, dtype=torch.float32)
    else:
        invstd = torch.empty(c, device=dev, dtype=torch.float32)
        lib.call("vspw_bn_fold_eval", _p(gv.data), _p(bv.data), _p(bn.running_mean), _p(bn.running_var), float(bn.eps),
                 _p(scale), _p(shift), _p(invstd), c, st)
    want_planes = planes_out and _state["precision"] != "fp32" and c % 64 == 0
    res_f = res_h = res_l = None
    if residual is not None:
        if residual.data is not None:
            res_f = residual.data
        else:
            res_h, res_l = residual.planes
    hi = lo = None
    if want_planes:
        hi = torch.empty(y.shape, device=dev, dtype=torch.bfloat16)
        lo = torch.empty(y.shape, device=dev, dtype=torch.bfloat16) if _state["precision"] == "bf16x3" else None
    o = torch.empty_like(y.data) if (fp32_out or not want_planes) else None
    needs = tape.grad_enabled and (y.needs_grad or gv.needs_grad or (residual is not None and residual.needs_grad))
    # the backward needs [out != 0] only: one bit per element (written by the forward kernel with warp ballots) instead of
    # re-reading the bf16 hi plane (2 B) or the fp32 output (4 B) in both backward passes
    bits = None
    # (tools/bench_bn.py: on the 16 M-element interior tensors the forward's ballots cost what the backward saves; from 32 M
    # elements up — block outputs, stem — the bits win 15-25 us per layer)
    if relu and needs and (c == 64 or c % 128 == 0) and pixels * c >= _RELU_BITS_MIN_ELEMS and os.environ.get("VSPW_RELU_BITS", "1") != "0":
        bits = torch.empty(((pixels * c // 4 + 31) // 32) * 4, device=dev, dtype=torch.int32)
    if training:
        # finalize (mean / invstd / running statistics from the fp64 sums) + normalise + residual + ReLU + planes: one launch
        if peer_ctx is not None:
            lib.call("vspw_bn_train_fwd_sync", _p(y.data), _p(sums), count, _p(gv.data), _p(bv.data), float(bn.eps),
                     float(bn.momentum), _p(bn.running_mean), _p(bn.running_var), _p(mean), _p(invstd),
                     1 if _state["syncbn_clamp"] else 0, _p(res_f), _p(res_h), _p(res_l), _p(chan_scale),
                     1 if relu else 0, _p(o), _p(hi), _p(lo), _p(bits), pixels, c, h * w, ctypes.byref(peer_ctx), st)
        else:
            lib.call("vspw_bn_train_fwd", _p(y.data), _p(sums[0]), _p(sums[1]), count, _p(gv.data), _p(bv.data), float(bn.eps),
                     float(bn.momentum), _p(bn.running_mean), _p(bn.running_var), _p(mean), _p(invstd),
                     1 if _state["syncbn_clamp"] else 0, _p(res_f), _p(res_h), _p(res_l), _p(chan_scale),
                     1 if relu else 0, _p(o), _p(hi), _p(lo), _p(bits), pixels, c, h * w, st)
    else:
        lib.call("vspw_bn_act_fwd", _p(y.data), _p(scale), _p(shift), _p(bn.running_mean), _p(bv.data),
                 _p(res_f), _p(res_h), _p(res_l), _p(chan_scale), 1 if relu else 0, _p(o), _p(hi), _p(lo), _p(bits), pixels, c, h * w, st)
    out = Var(o, needs_grad=needs)
    if want_planes:
        out.planes = (hi, lo)
    # ReLU mask source in the backward: the bf16 hi plane when there is one (2 bytes/element instead of 4), else fp32
    mask_hi = hi
    mask_o = o if hi is None else None

    def backward():
        dout = out.grad
        out.grad = None
        if dout is None:
            return
        st = _stream()
        # the consumer of dy is the producing conv's dgrad/wgrad: tcgen05 kernels read bf16 planes, so write those
        # here instead of an fp32 dy plus a separate split pass (fp32 dy only if someone else needs it)
        dy = dy_hi = dy_lo = None
        if y.wants_grad_planes and y.grad is None:
            dy_hi = torch.empty(dout.shape, device=dev, dtype=torch.bfloat16)
            dy_lo = torch.empty(dout.shape, device=dev, dtype=torch.bfloat16) if _state["precision"] == "bf16x3" else None
        if dy_hi is None or y.wants_grad_fp32:
            dy = torch.empty_like(dout)
        dres = torch.empty_like(dout) if (residual is not None and residual.needs_grad) else None
        if training:
            dsum = tape.zeros_f64((2, c), dev)
            lib.call("vspw_bn_bwd_reduce", _p(dout), _p(mask_o), _p(mask_hi), _p(y.data), _p(mean), _p(invstd), _p(chan_scale),
                     1 if relu else 0, _p(bits), pixels, c, h * w, _p(dsum[0]), _p(dsum[1]), st)
            bctx = group.next_ctx(2 * c) if (group is not None and hasattr(group, "next_ctx")) else None
            if group is not None and bctx is None:
                # dx needs the all-rank sums; gamma/beta gradients leave as sums/world because the gradient all-reduce that
                # follows AVERAGES the ranks' parameter gradients and every rank holds the same all-rank total here
                group.all_reduce_sums(dsum)
            gdst = gv.first_dst() if gv.needs_grad else None
            bdst = bv.first_dst() if bv.needs_grad else None
            dgam = gdst if gdst is not None else torch.empty(c, device=dev, dtype=torch.float32)
            dbet = bdst if bdst is not None else torch.empty(c, device=dev, dtype=torch.float32)
            if bctx is not None:
                lib.call("vspw_bn_bwd_apply_sync", _p(dout), _p(mask_o), _p(mask_hi), _p(y.data), _p(mean), _p(invstd), _p(gv.data),
                         _p(chan_scale), 1 if relu else 0, _p(dsum), _p(dy), _p(dy_hi), _p(dy_lo), _p(dres), _p(dgam), _p(dbet),
                         _p(bits), pixels, c, h * w, float(pixels * world), 1.0 / world, ctypes.byref(bctx), st)
            else:
                lib.call("vspw_bn_bwd_apply", _p(dout), _p(mask_o), _p(mask_hi), _p(y.data), _p(mean), _p(invstd), _p(gv.data),
                         _p(chan_scale), 1 if relu else 0, _p(dsum[0]), _p(dsum[1]), _p(dy), _p(dy_hi), _p(dy_lo), _p(dres), _p(dgam),
                         _p(dbet), _p(bits), pixels, c, h * w, 0, float(pixels * world), 1.0 / world, st)
            if gv.needs_grad and gdst is None:
                gv.add_grad(dgam)
            if bv.needs_grad and bdst is None:
                bv.add_grad(dbet)
        else:
            # frozen statistics (cfg.TRAIN.fix_bn): dy = g * gamma/sqrt(var+eps) = g * scale; gamma/beta still learn:
            # dbeta = sum g, dgamma = sum g * (y - running_mean) * invstd
            dsum = dgam = dbet = None
            if gv.needs_grad or bv.needs_grad:
                dsum = tape.zeros_f64((2, c), dev)
                lib.call("vspw_bn_bwd_reduce", _p(dout), _p(mask_o), _p(mask_hi), _p(y.data), _p(bn.running_mean), _p(invstd),
                         _p(chan_scale), 1 if relu else 0, _p(bits), pixels, c, h * w, _p(dsum[0]), _p(dsum[1]), st)
                dgam = torch.empty(c, device=dev, dtype=torch.float32)
                dbet = torch.empty(c, device=dev, dtype=torch.float32)
            lib.call("vspw_bn_bwd_apply", _p(dout), _p(mask_o), _p(mask_hi), None, None, _p(scale), None, _p(chan_scale),
                     1 if relu else 0, _p(dsum[0]) if dsum is not None else None, _p(dsum[1]) if dsum is not None else None,
                     _p(dy), _p(dy_hi), _p(dy_lo), _p(dres), _p(dgam), _p(dbet), _p(bits), pixels, c, h * w, 1, float(pixels), 1.0, st)
            if dsum is not None:
                if gv.needs_grad:
                    gv.add_grad(dgam)
                if bv.needs_grad:
                    bv.add_grad(dbet)
        if y.needs_grad:
            if dy_hi is not None:
                y.grad_planes = (dy_hi, dy_lo)
            if dy is not None:
                y.add_grad(dy)
        if dres is not None:
            residual.add_grad(dres)

    tape.record(backward)
    return out


def maxpool3x3s2(tape, x):
    n, h, w, c = x.shape
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = torch.empty((n, ho, wo, c), device=x.data.device, dtype=torch.float32)
    idx = torch.empty((n, ho, wo, c), device=x.data.device, dtype=torch.uint8) if tape.grad_enabled and x.needs_grad else None
    lib.call("vspw_maxpool3x3s2_fwd", _p(x.data), _p(y), _p(idx), n, h, w, c, ho, wo, _stream())
    out = Var(y, needs_grad=tape.grad_enabled and x.needs_grad)

    def backward():
        dy = out.grad
        out.grad = None
        if dy is None or not x.needs_grad:
            return
        dx = torch.empty_like(x.data)
        lib.call("vspw_maxpool3x3s2_bwd", _p(dy), _p(idx), _p(dx), n, h, w, c, ho, wo, _stream())
        x.add_grad(dx)

    tape.record(backward)
    return out


def slice_images(tape, x, n0, n1):
    """x[n0:n1] along the image axis (torch.split(dim=0), clip_psp.py:154-156)."""
    out = Var(x.data[n0:n1], needs_grad=x.needs_grad)
    if x.planes is not None:
        out.planes = tuple(p[n0:n1] if p is not None else None for p in x.planes)

    def backward():
        g = out.grad
        out.grad = None
        if g is not None and x.needs_grad:
            x.add_grad_rows(g, n0, n1)

    tape.record(backward)
    return out


def tcb_pool(tape, feat, t_frames, n_clips, scales, frame_w=None):
    """Temporal pyramid pooling: list of Vars (n_clips, s, s, C), one per scale (clip_psp.py:157-188)."""
    N, h, w, c = feat.shape
    assert N == t_frames * n_clips
    dev = feat.data.device
    total_bins = sum(s * s for s in scales)
    pooled = torch.empty(n_clips * total_bins * c, device=dev, dtype=torch.float32)
    sc = (ctypes.c_int32 * len(scales))(*scales)
    fw = frame_w.data if frame_w is not None else None
    ws = torch.empty(lib.dll().vspw_tcb_pool_workspace_floats(t_frames, n_clips, h, c, sc, len(scales)), device=dev, dtype=torch.float32)
    lib.call("vspw_tcb_pool_fwd", _p(feat.data), _p(fw), _p(pooled), _p(ws), t_frames, n_clips, h, w, c, sc, len(scales), _stream())
    outs, off = [], 0
    needs = tape.grad_enabled and (feat.needs_grad or (frame_w is not None and frame_w.needs_grad))
    for s in scales:
        cnt = n_clips * s * s * c
        outs.append(Var(pooled[off:off + cnt].view(n_clips, s, s, c), needs_grad=needs))
        off += cnt

    def backward():
        dp = torch.empty_like(pooled)
        off = 0
        any_grad = False
        for s, o in zip(scales, outs):
            cnt = n_clips * s * s * c
            if o.grad is not None:
                any_grad = True
                lib.call("vspw_axpby", _p(o.grad), _p(dp[off:off + cnt]), 1.0, 0.0, cnt, _stream())
            else:
                lib.call("vspw_fill", _p(dp[off:off + cnt]), 0.0, cnt, _stream())
            o.grad = None
            off += cnt
        if not any_grad:
            return
        dfeat = torch.empty_like(feat.data) if feat.needs_grad else None
        dfw = None
        if frame_w is not None and frame_w.needs_grad:
            dfw = torch.zeros_like(frame_w.data)
        if dfeat is None and dfw is None:
            return
        if dfeat is None:
            dfeat = torch.empty_like(feat.data)
        lib.call("vspw_tcb_pool_bwd", _p(dp), _p(fw), _p(feat.data), _p(dfeat), _p(dfw), t_frames, n_clips, h, w, c, sc,
                 len(scales), _stream())
        if feat.needs_grad:
            feat.add_grad(dfeat)
        if dfw is not None:
            frame_w.add_grad(dfw)

    tape.record(backward)
    return outs


def frame_weights(tape, score, t_frames, n_clips):
    """The psp_weight branch of Clip_PSP after its 1x1 conv (clip_psp.py:147-152, 184-187): global average of the
    (N, h, w, 1) score map per image (AdaptiveAvgPool2d((1,1))), softmax over the T frames of each clip, and the
    reference's list-order quirk Q3: the weight taken from batch chunk j multiplies list position j, where position 0
    is the CURRENT frame (batch chunk T-1) and position k >= 1 is batch chunk k-1.  Returns a (T, n) Var laid out for
    vspw_tcb_pool: row t = the weight applied to batch chunk t = softmax row (t+1) % T."""
    N, h, w, one = score.shape
    assert one == 1 and N == t_frames * n_clips
    dev = score.data.device
    hw = h * w
    st = _stream()
    ones = torch.empty(hw, device=dev, dtype=torch.float32)
    lib.call("vspw_fill", _p(ones), 1.0, hw, st)
    raw = torch.empty(N, device=dev, dtype=torch.float32)
    # raw[b] = (1/hw) sum_p score[b][p] * 1
    lib.call("vspw_bgemm", _p(score.data), _p(ones), _p(raw), N, 1, 1, hw, hw, hw, 1, 0, 1, 1, 1, 1, 1, 1.0 / hw, 0.0, st)
    sm = torch.empty_like(raw)  # [t][i]; softmax along t for every clip i
    lib.call("vspw_softmax_strided_fwd", _p(raw), _p(sm), n_clips, t_frames, 1, n_clips, n_clips, 0, 1.0, st)
    fw = torch.empty((t_frames, n_clips), device=dev, dtype=torch.float32)
    for t in range(t_frames):
        src = sm[((t + 1) % t_frames) * n_clips:((t + 1) % t_frames + 1) * n_clips]
        lib.call("vspw_axpby", _p(src), _p(fw[t]), 1.0, 0.0, n_clips, st)
    out = Var(fw, needs_grad=tape.grad_enabled and score.needs_grad)

    def backward():
        g = out.grad
        out.grad = None
        if g is None or not score.needs_grad:
            return
        st = _stream()
        dsm = torch.empty_like(sm)
        for t in range(t_frames):
            j = (t + 1) % t_frames
            lib.call("vspw_axpby", _p(g[t]), _p(dsm[j * n_clips:(j + 1) * n_clips]), 1.0, 0.0, n_clips, st)
        draw = torch.empty_like(raw)
        lib.call("vspw_softmax_strided_bwd", _p(sm), _p(dsm), _p(draw), n_clips, t_frames, 1, n_clips, n_clips, 0, 1.0, st)
        dscore = torch.empty_like(score.data)
        # dscore[b][p] = draw[b] / hw
        lib.call("vspw_bgemm", _p(draw), _p(ones), _p(dscore), N, 1, hw, 1, 1, 1, 1, 0, 1, 1, hw, hw, 1, 1.0 / hw, 0.0, st)
        score.add_grad(dscore)

    tape.record(backward)
    return out


def ppm_concat(tape, base, pyramids):
    """cat([base] + [bilinear_up(p) for p in pyramids], channel axis) (clip_psp.py:45-53)."""
    n, h, w, c0 = base.shape
    ctot = c0 + sum(p.shape[3] for p in pyramids)
    dev = base.data.device
    cat = torch.empty((n, h, w, ctot), device=dev, dtype=torch.float32)
    st = _stream()
    lib.call("vspw_copy_channels", _p(base.data), c0, 0, _p(cat), ctot, 0, c0, n * h * w, 0, st)
    offs, off = [], c0
    for p in pyramids:
        pn, sh, sw, pc = p.shape
        lib.call("vspw_upsample_bilinear_fwd", _p(p.data), pn, sh, sw, pc, _p(cat), h, w, ctot, off, st)
        offs.append(off)
        off += pc
    out = Var(cat, needs_grad=tape.grad_enabled and (base.needs_grad or any(p.needs_grad for p in pyramids)))

    def backward():
        g = out.grad
        out.grad = None
        if g is None:
            return
        st = _stream()
        if base.needs_grad:
            db = torch.empty_like(base.data)
            lib.call("vspw_copy_channels", _p(g), ctot, 0, _p(db), c0, 0, c0, n * h * w, 0, st)
            base.add_grad(db)
        for p, o in zip(pyramids, offs):
            if not p.needs_grad:
                continue
            pn, sh, sw, pc = p.shape
            dp = torch.empty_like(p.data)
            lib.call("vspw_upsample_bilinear_bwd", _p(g), h, w, ctot, o, _p(dp), pn, sh, sw, pc, st)
            p.add_grad(dp)

    tape.record(backward)
    return out


def ppm_fused_supported(base_shape, pyramid_shapes, weight_shape, pad, dil):
    """True when ppm_conv_fused can take this head in the current precision mode (else: ppm_concat + conv2d)."""
    prec = _PRECISION[_state["precision"]]
    if prec == PREC_FP32 or os.environ.get("VSPW_PPM_FUSED", "1") == "0":
        return False
    n, h, w, c0 = base_shape
    co, ctot, kh, kw = weight_shape
    if kh != kw or kh not in (1, 3) or pad != dil * (kh - 1) // 2 or co % 4:
        return False
    cps = {s[3] for s in pyramid_shapes}
    if len(cps) != 1 or c0 + len(pyramid_shapes) * next(iter(cps)) != ctot or len(pyramid_shapes) > 8:
        return False
    if any(s[1] != s[2] or s[0] != n for s in pyramid_shapes):
        return False
    return bool(lib.tc_supported(ConvDesc(n, h, w, c0, co, kh, kw, 1, pad, dil, h, w, prec, ctot)))


def ppm_conv_fused(tape, base, pyramids, weight, pad=1, dil=1, want_stats=False):
    """conv(cat([base] + [bilinear_up(p) for p in pyramids], channels), weight) WITHOUT the up-sampled maps and the concat
    (PPM_conv.forward, clip_psp.py:45-56; PPMDeepsup.forward, models/models.py:975-990; bias-free conv).

    Linearity: the `base` channels run on the tcgen05 conv kernel straight from base's operand planes against the first C0
    input channels of the weight (ConvDesc.cin_pitch); every pyramid branch is evaluated in bin space — Z_s = P_s . W_s^T
    (a (n s^2) x Cp x (9 Cout) GEMM), then y[p] += sum_tap sum_bins B_s[p + off(tap), bin] Z_s[bin, tap] (csrc/ppm.cu).
    At 480p: 242 instead of 485 GFLOP forward, no 210 MB concat (+ its planes and gradient)."""
    wv = tape.param(weight)
    n, h, w, c0 = base.shape
    co, ctot, kh, kw = wv.data.shape
    S = len(pyramids)
    cp = pyramids[0].shape[3]
    scales = [p.shape[1] for p in pyramids]
    taps = kh * kw
    dev = base.planes[0].device if base.data is None else base.data.device
    prec = _PRECISION[_state["precision"]]
    desc = ConvDesc(n, h, w, c0, co, kh, kw, 1, pad, dil, h, w, prec, ctot)
    st = _stream()
    xh, xl = _var_planes(base)
    wplanes = _tc_weight_planes(tape, wv)
    wh, wl = wplanes["ohwi"]
    y = torch.empty((n, h, w, co), device=dev, dtype=torch.float32)
    flops_base = 2.0 * n * h * w * co * taps * c0
    flops_pyr = 2.0 * n * h * w * co * taps * cp * S   # algorithmic (reference) count of the part evaluated in bin space
    with _ConvTimer(flops_base, True):
        lib.call("vspw_conv2d_fwd_tc", ctypes.byref(desc), _p(xh), _p(xl), _p(wh), _p(wl), None, _p(y), None, None, st)
    sc = (ctypes.c_int32 * S)(*scales)
    with _ConvTimer(flops_pyr, False, f"ppm pyramid (bin space) fwd {n}x{h}x{w} {S}x{cp}->{co}"):
        wp = torch.empty((S, taps, co, cp), device=dev, dtype=torch.float32)
        lib.call("vspw_ppm_weight_slices", _p(wv.data), _p(wp), co, ctot, kh, kw, c0, cp, S, st)
        zs = []
        for i, pv in enumerate(pyramids):
            m = n * scales[i] * scales[i]
            z = torch.empty((m, taps * co), device=dev, dtype=torch.float32)
            # Z[bin][j=(tap,co)] = sum_c P[bin][c] * Wp[i][j][c]
            # (no split-K: the inference path stays free of atomics, i.e. bit-reproducible)
            lib.call("vspw_bgemm_det", _p(pv.data), _p(wp[i]), _p(z), 1, m, taps * co, cp, 0, cp, 1, 0, 1, cp, 0, taps * co, 1, 1.0, 0.0, st)
            zs.append(z)
        stats = tape.zeros_f64((2, co), dev) if want_stats else None
        zp = (ctypes.c_void_p * S)(*[z.data_ptr() for z in zs])
        lib.call("vspw_ppm_pyramid_fwd", _p(y), zp, sc, S, n, h, w, co, kh, pad, dil, _p(stats[0]) if stats is not None else None,
                 _p(stats[1]) if stats is not None else None, st)
    del zs
    out = Var(y, needs_grad=tape.grad_enabled and (base.needs_grad or wv.needs_grad or any(p.needs_grad for p in pyramids)))
    out.stats = stats
    out.wants_grad_planes = True
    out.wants_grad_fp32 = True  # the bin-space gather reads the fp32 gradient

    def backward():
        dy, dyp = out.grad, out.grad_planes
        out.grad = out.grad_planes = None
        if dy is None:
            return
        st = _stream()
        if dyp is None:
            dyp = _planes_of(dy)
        # ---- pyramid part: dZ (transposed gather), then dP_s and dW_s as small GEMMs ---------------------------------
        dzs = [torch.empty((n * s * s, taps * co), device=dev, dtype=torch.float32) for s in scales]
        with _ConvTimer(2 * flops_pyr, False, f"ppm pyramid (bin space) bwd {n}x{h}x{w} {S}x{cp}->{co}"):
            dzp = (ctypes.c_void_p * S)(*[z.data_ptr() for z in dzs])
            lib.call("vspw_ppm_pyramid_bwd", _p(dy), dzp, sc, S, n, h, w, co, kh, pad, dil, st)
            for i, pv in enumerate(pyramids):
                if not pv.needs_grad:
                    continue
                m = n * scales[i] * scales[i]
                dp = torch.empty((m, cp), device=dev, dtype=torch.float32)
                # dP[bin][c] = sum_j dZ[bin][j] * Wp[i][j][c]
                lib.call("vspw_bgemm", _p(dzs[i]), _p(wp[i]), _p(dp), 1, m, cp, taps * co, 0, taps * co, 1, 0, cp, 1, 0, cp, 1, 1.0, 0.0, st)
                pv.add_grad(dp.view(pv.shape))
        if wv.needs_grad:
            dw = torch.empty((co, kh, kw, ctot), device=dev, dtype=torch.float32)
            with _ConvTimer(flops_base, True):
                lib.call("vspw_conv2d_wgrad_tc", ctypes.byref(desc), _p(xh), _p(xl), _p(dyp[0]), _p(dyp[1]), _p(dw), st)
            for i, pv in enumerate(pyramids):
                m = n * scales[i] * scales[i]
                # dW[co][tap][c0 + i*cp + c] = sum_bin dZ[bin][tap][co] * P[bin][c]  (batch = tap)
                lib.call("vspw_bgemm", _p(dzs[i]), _p(pv.data), _p(dw.view(-1)[c0 + i * cp:]), taps, co, cp, m, co, 1, taps * co, 0, cp, 1,
                         ctot, taps * ctot, 1, 1.0, 0.0, st)
            dst = wv.first_dst()
            dw_oihw = dst.view(co, ctot, kh, kw) if dst is not None else torch.empty((co, ctot, kh, kw), device=dev, dtype=torch.float32)
            permute4d(dw, dw_oihw, (co, kh, kw, ctot), (0, 3, 1, 2))
            if dst is None:
                wv.add_grad(dw_oihw)
        if base.needs_grad:
            th, tl = wplanes["ihwo"]  # [ctot][kh][kw][co]: rows [0, c0) are the base channels
            fan_in = base.grad is not None and base.grad.is_contiguous() and tuple(base.grad.shape) == (n, h, w, c0)
            dx = base.grad if fan_in else torch.empty((n, h, w, c0), device=dev, dtype=torch.float32)
            d1 = ConvDesc(n, h, w, c0, co, kh, kw, 1, pad, dil, h, w, prec)
            with _ConvTimer(flops_base, True):
                lib.call("vspw_conv2d_dgrad_tc", ctypes.byref(d1), _p(dyp[0]), _p(dyp[1]), _p(th), _p(tl), _p(dx), 1 if fan_in else 0, st)
            if not fan_in:
                base.add_grad(dx)

    tape.record(backward)
    return out


def upsample_bilinear(tape, x, H, W):
    """F.interpolate(x, size=(H, W), mode='bilinear', align_corners=False) on an NHWC Var (UPerNet's top-down and fusion
    branches, models/models.py:1138-1164)."""
    n, h, w, c = x.shape
    if (h, w) == (H, W):
        return x
    dev = x.data.device
    y = torch.empty((n, H, W, c), device=dev, dtype=torch.float32)
    lib.call("vspw_upsample_bilinear_fwd", _p(x.data), n, h, w, c, _p(y), H, W, c, 0, _stream())
    out = Var(y, needs_grad=tape.grad_enabled and x.needs_grad)

    def backward():
        g = out.grad
        out.grad = None
        if g is None or not x.needs_grad:
            return
        dx = torch.empty_like(x.data)
        lib.call("vspw_upsample_bilinear_bwd", _p(g), H, W, c, 0, _p(dx), n, h, w, c, _stream())
        x.add_grad(dx)

    tape.record(backward)
    return out


def concat_channels(tape, parts):
    """torch.cat(parts, dim=1) in NHWC (spatial_ocr_block.py:375)."""
    n, h, w, _ = parts[0].shape
    ctot = sum(p.shape[3] for p in parts)
    cat = torch.empty((n, h, w, ctot), device=parts[0].data.device, dtype=torch.float32)
    st = _stream()
    offs, off = [], 0
    for p in parts:
        pc = p.shape[3]
        lib.call("vspw_copy_channels", _p(p.data), pc, 0, _p(cat), ctot, off, pc, n * h * w, 0, st)
        offs.append(off)
        off += pc
    out = Var(cat, needs_grad=tape.grad_enabled and any(p.needs_grad for p in parts))

    def backward():
        g = out.grad
        out.grad = None
        if g is None:
            return
        for p, o in zip(parts, offs):
            if not p.needs_grad:
                continue
            pc = p.shape[3]
            dp = torch.empty_like(p.data)
            lib.call("vspw_copy_channels", _p(g), ctot, o, _p(dp), pc, 0, pc, n * h * w, 0, _stream())
            p.add_grad(dp)

    tape.record(backward)
    return out


class LossTerm:
    """One `crit(interpolate(log_softmax(logits)), label)` term; accumulators stay on the device."""

    def __init__(self, logits, labels, ignore_index, want_acc):
        self.logits, self.labels, self.ignore_index, self.want_acc = logits, labels, ignore_index, want_acc
        self.acc = None
        self.logp = None


def nll_term(tape, logits, labels, ignore_index, want_acc):
    """log_softmax -> bilinear up -> NLL (+pixel_acc) sums (clip_psp.py:196-217)."""
    n, h, w, k = logits.shape
    ln, lc, H, W = labels.shape
    if ln != n or lc != 1:
        raise VspwError(f"labels {tuple(labels.shape)} do not match logits batch {n}")
    term = LossTerm(logits, labels, ignore_index, want_acc)
    dev = logits.data.device
    term.acc = torch.empty(4, device=dev, dtype=torch.float64)
    term.logp = torch.empty_like(logits.data)
    lib.call("vspw_logsoftmax_up_nll_fwd", _p(logits.data), _p(labels), _p(term.logp), _p(term.acc), n, h, w, k, H, W,
             ignore_index, 1 if want_acc else 0, _stream())
    return term


def loss_combine(tape, main, aux, aux_scale):
    """loss = main + aux_scale*aux, acc from main; returns (loss, acc) 0-d tensors and records backward."""
    dev = main.logits.data.device
    loss = torch.empty((), device=dev, dtype=torch.float32)
    pixacc = torch.empty((), device=dev, dtype=torch.float32)
    lib.call("vspw_loss_finalize", _p(main.acc), _p(aux.acc if aux is not None else None), float(aux_scale), _p(loss),
             _p(pixacc), _stream())
    gslot = {"g": None}

    def backward():
        g = gslot["g"]  # device scalar: upstream d/d(loss)
        st = _stream()
        for term, scale in ((main, 1.0), (aux, aux_scale)):
            if term is None or not term.logits.needs_grad:
                continue
            n, h, w, k = term.logits.shape
            H, W = term.labels.shape[2], term.labels.shape[3]
            dl = torch.empty_like(term.logits.data)
            scratch = torch.empty_like(term.logits.data)
            lib.call("vspw_logsoftmax_up_nll_bwd", _p(term.logp), _p(term.labels), _p(term.acc), _p(g), float(scale), _p(dl),
                     _p(scratch), n, h, w, k, H, W, term.ignore_index, st)
            term.logits.add_grad(dl)

    tape.record(backward)
    return loss, pixacc, gslot


def loss_mean_of_terms(tape, terms):
    """loss = mean_t NLL_t, acc = mean_t acc_t over per-frame terms (Non_local3d.forward, non_local_models.py:49-61:
    every frame is supervised and the per-frame MEANS are averaged).  Returns (loss, acc, gslot)."""
    dev = terms[0].logits.data.device
    st = _stream()
    loss = torch.empty((), device=dev, dtype=torch.float32)
    pixacc = torch.empty((), device=dev, dtype=torch.float32)
    inv = 1.0 / len(terms)
    for i, term in enumerate(terms):
        li = torch.empty((), device=dev, dtype=torch.float32)
        ai = torch.empty((), device=dev, dtype=torch.float32)
        lib.call("vspw_loss_finalize", _p(term.acc), None, 0.0, _p(li), _p(ai), st)
        lib.call("vspw_axpby", _p(li), _p(loss), inv, 0.0 if i == 0 else 1.0, 1, st)
        lib.call("vspw_axpby", _p(ai), _p(pixacc), inv, 0.0 if i == 0 else 1.0, 1, st)
    gslot = {"g": None}

    def backward():
        g = gslot["g"]
        st = _stream()
        for term in terms:
            if not term.logits.needs_grad:
                continue
            n, h, w, k = term.logits.shape
            H, W = term.labels.shape[2], term.labels.shape[3]
            dl = torch.empty_like(term.logits.data)
            scratch = torch.empty_like(term.logits.data)
            lib.call("vspw_logsoftmax_up_nll_bwd", _p(term.logp), _p(term.labels), _p(term.acc), _p(g), float(inv), _p(dl),
                     _p(scratch), n, h, w, k, H, W, term.ignore_index, st)
            term.logits.add_grad(dl)

    tape.record(backward)
    return loss, pixacc, gslot


def nl_dot_affinity(tape, theta, phi, g, t_frames, n_clips):
    """NLBlockND(mode='dot', dimension=3) core (non_local.py:106-136): for every clip, over its P = T*h*w positions,
    y = (theta . phi^T / P) . g.  Without a softmax the P x P affinity never has to exist: by associativity
    y = theta . M with M = phi^T . g / P, a C x C matrix per clip (C = 128) — 2*P*C^2 instead of 2*P^2*C multiply-adds
    (2.1 GFLOP instead of 0.53 TFLOP per clip at 480p T=5), the same function up to fp32 rounding.
    theta, phi, g: (T*n, h, w, C) Vars with image index t*n + clip.  Returns y with the same layout."""
    N, h, w, c = theta.shape
    assert N == t_frames * n_clips
    hw = h * w
    P = t_frames * hw
    dev = theta.data.device
    st = _stream()
    M = torch.empty((n_clips, c, c), device=dev, dtype=torch.float32)  # [clip][c_phi][c_g]
    for t in range(t_frames):
        ph, gg = phi.data[t * n_clips:(t + 1) * n_clips], g.data[t * n_clips:(t + 1) * n_clips]
        # M[b][i][j] (+)= (1/P) sum_p phi[b][p][i] * g[b][p][j]
        lib.call("vspw_bgemm", _p(ph), _p(gg), _p(M), n_clips, c, c, hw, hw * c, 1, c, hw * c, c, 1, c * c, c, 1, 1.0 / P,
                 0.0 if t == 0 else 1.0, st)
    y = torch.empty_like(theta.data)
    for t in range(t_frames):
        th = theta.data[t * n_clips:(t + 1) * n_clips]
        # y[b][p][j] = sum_i theta[b][p][i] * M[b][i][j]
        lib.call("vspw_bgemm", _p(th), _p(M), _p(y[t * n_clips:(t + 1) * n_clips]), n_clips, hw, c, c, hw * c, c, 1, c * c, c, 1,
                 hw * c, c, 1, 1.0, 0.0, st)
    out = Var(y, needs_grad=tape.grad_enabled and (theta.needs_grad or phi.needs_grad or g.needs_grad))

    def backward():
        dy = out.grad
        out.grad = None
        if dy is None:
            return
        st = _stream()
        dM = torch.empty_like(M)
        for t in range(t_frames):
            sl = slice(t * n_clips, (t + 1) * n_clips)
            # dM[b][i][j] (+)= sum_p theta[b][p][i] * dy[b][p][j]
            lib.call("vspw_bgemm", _p(theta.data[sl]), _p(dy[sl]), _p(dM), n_clips, c, c, hw, hw * c, 1, c, hw * c, c, 1, c * c, c, 1,
                     1.0, 0.0 if t == 0 else 1.0, st)
        if theta.needs_grad:
            dth = torch.empty_like(theta.data)
            for t in range(t_frames):
                sl = slice(t * n_clips, (t + 1) * n_clips)
                # dtheta[b][p][i] = sum_j dy[b][p][j] * M[b][i][j]
                lib.call("vspw_bgemm", _p(dy[sl]), _p(M), _p(dth[sl]), n_clips, hw, c, c, hw * c, c, 1, c * c, 1, c, hw * c, c, 1,
                         1.0, 0.0, st)
            theta.add_grad(dth)
        if phi.needs_grad:
            dph = torch.empty_like(phi.data)
            for t in range(t_frames):
                sl = slice(t * n_clips, (t + 1) * n_clips)
                # dphi[b][p][i] = (1/P) sum_j g[b][p][j] * dM[b][i][j]
                lib.call("vspw_bgemm", _p(g.data[sl]), _p(dM), _p(dph[sl]), n_clips, hw, c, c, hw * c, c, 1, c * c, 1, c, hw * c, c, 1,
                         1.0 / P, 0.0, st)
            phi.add_grad(dph)
        if g.needs_grad:
            dg = torch.empty_like(g.data)
            for t in range(t_frames):
                sl = slice(t * n_clips, (t + 1) * n_clips)
                # dg[b][p][j] = (1/P) sum_i phi[b][p][i] * dM[b][i][j]
                lib.call("vspw_bgemm", _p(phi.data[sl]), _p(dM), _p(dg[sl]), n_clips, hw, c, c, hw * c, c, 1, c * c, c, 1, hw * c, c, 1,
                         1.0 / P, 0.0, st)
            g.add_grad(dg)

    tape.record(backward)
    return out


def add_vars(tape, a, b):
    """a + b (the residual `z = W_y + x` of NLBlockND, non_local.py:150)."""
    o = torch.empty_like(a.data)
    lib.call("vspw_axpby", _p(a.data), _p(o), 1.0, 0.0, o.numel(), _stream())
    lib.call("vspw_axpby", _p(b.data), _p(o), 1.0, 1.0, o.numel(), _stream())
    out = Var(o, needs_grad=tape.grad_enabled and (a.needs_grad or b.needs_grad))

    def backward():
        gr = out.grad
        out.grad = None
        if gr is None:
            return
        for v in (a, b):
            if v.needs_grad:
                d = torch.empty_like(gr)
                lib.call("vspw_axpby", _p(gr), _p(d), 1.0, 0.0, gr.numel(), _stream())
                v.add_grad(d)

    tape.record(backward)
    return out


def up_softmax(logits, H, W, want_pred=False):
    """Inference tail: bilinear to (H, W) then softmax(dim=1); NCHW probabilities (clip_psp.py:190-194)."""
    n, h, w, k = logits.shape
    dev = logits.data.device
    probs = torch.empty((n, k, H, W), device=dev, dtype=torch.float32)
    pred = torch.empty((n, H, W), device=dev, dtype=torch.int32) if want_pred else None
    lib.call("vspw_up_softmax_fwd", _p(logits.data), _p(probs), _p(pred), n, h, w, k, H, W, _stream())
    return (probs, pred) if want_pred else probs


# ------------------------------------------------------------------------------------------------
# OCR ops
def _ocr_tc_ok():
    """The tcgen05 OCR kernels (csrc/ocr_tc.cu) run in the tensor-core precision modes; VSPW_OCR_TC=0 keeps the CUDA-core
    fp32 GEMMs (A/B timing)."""
    return _state["precision"] != "fp32" and os.environ.get("VSPW_OCR_TC", "1") != "0"


def _conv1x1_tc_raw(xh, xl, wh, wl, y, pixels, cin, cout, st):
    """y[pixels][cout] = x[pixels][cin] . w[cout][cin]^T on the tcgen05 conv kernel from raw operand planes (one image's slice of
    a planes tensor, a per-image weight): the small GEMMs of the OCR backward."""
    prec = PREC_BF16X3 if xl is not None else PREC_BF16
    desc = ConvDesc(1, 1, pixels, cin, cout, 1, 1, 1, 0, 1, 1, pixels, prec)
    lib.call("vspw_conv2d_fwd_tc", ctypes.byref(desc), _p(xh), _p(xl), _p(wh), _p(wl), None, _p(y), None, None, st)


def _operand_planes(src, n, rows, cols, rows_pad, transpose, scale, x3, st):
    shape = (n, cols, rows_pad) if transpose else (n, rows_pad, cols)
    hi = torch.empty(shape, device=src.device, dtype=torch.bfloat16)
    lo = torch.empty(shape, device=src.device, dtype=torch.bfloat16) if x3 else None
    lib.call("vspw_ocr_operand_planes", _p(src), _p(hi), _p(lo), n, rows, cols, rows_pad, 1 if transpose else 0, float(scale), st)
    return hi, lo


def region_gather(tape, feats, dsn, t_frames, n_clips):
    """SpatialTemporalGather_Module (spatial_ocr_block.py:97-109): per frame softmax over hw of the dsn
    logits, probs[K x hw] . feats[hw x C], mean over the T frames -> context Var (n_clips, K, 1, C)."""
    N, h, w, c = feats.shape
    k = dsn.shape[3]
    hw = h * w
    dev = feats.data.device
    st = _stream()
    probs = torch.empty_like(dsn.data)  # [N][hw][K]
    ctx = torch.empty((n_clips, k, 1, c), device=dev, dtype=torch.float32)
    inv_t = 1.0 / t_frames
    use_tc = _ocr_tc_ok() and k <= 128 and c % 128 == 0
    sm_ws = torch.empty(int(lib.dll().vspw_ocr_region_softmax_workspace_bytes(N, k)), device=dev, dtype=torch.uint8) if k <= 128 else None
    if use_tc:
        # tcgen05: the gather is the weight-gradient kernel's GEMM (K axis = pixels, both operands channel-contiguous) over
        # the probability planes (classes padded to 128, 1/T folded in; written by the softmax sweep itself) and the operand
        # planes of feats
        x3 = _state["precision"] == "bf16x3"
        ph = torch.empty((N, hw, 128), device=dev, dtype=torch.bfloat16)
        pl = torch.empty((N, hw, 128), device=dev, dtype=torch.bfloat16) if x3 else None
        lib.call("vspw_ocr_region_softmax_fwd", _p(dsn.data), _p(probs), _p(ph), _p(pl), _p(sm_ws), N, hw, k, inv_t, st)
        fh, fl = _var_planes(feats)
        with _ConvTimer(2.0 * N * hw * k * c, True):
            lib.call("vspw_ocr_gather_tc", _p(ph), _p(pl), _p(fh), _p(fl), _p(ctx), t_frames, n_clips, hw, k, c, st)
        if not (tape.grad_enabled and (feats.needs_grad or dsn.needs_grad)):
            del ph, pl
    else:
        if sm_ws is not None:
            lib.call("vspw_ocr_region_softmax_fwd", _p(dsn.data), _p(probs), None, None, _p(sm_ws), N, hw, k, 1.0, st)
        else:
            lib.call("vspw_softmax_strided_fwd", _p(dsn.data), _p(probs), N * k, hw, 1, k, k, hw * k, 1.0, st)
        for t in range(t_frames):
            pr = probs[t * n_clips:(t + 1) * n_clips]
            ft = feats.data[t * n_clips:(t + 1) * n_clips]
            # C[b][i=class][j=ch] = sum_p probs[b][p][i] * feats[b][p][j]
            lib.call("vspw_bgemm", _p(pr), _p(ft), _p(ctx), n_clips, k, c, hw, hw * k, 1, k, hw * c, c, 1, k * c, c, 1, inv_t,
                     0.0 if t == 0 else 1.0, st)
    out = Var(ctx, needs_grad=tape.grad_enabled and (feats.needs_grad or dsn.needs_grad))

    def backward():
        g = out.grad
        out.grad = None
        if g is None:
            return
        st = _stream()
        if use_tc and hw >= 256:  # (a per-image 1x1 GEMM needs >= 256 pixels on the conv kernel; tiny maps keep the fp32 GEMMs)
            # tcgen05 backward: dF = P . g and dP = F . g^T / T as per-image 1x1 GEMMs on the conv kernel (the context gradient of
            # the image's clip is the weight), then the column softmax backward on the 128-pitch dP
            gc = g.view(n_clips, k, c)
            if feats.needs_grad:
                gt_h, gt_l = _operand_planes(gc, n_clips, k, c, 128, True, 1.0, x3, st)      # [n][c][128]
                df = torch.empty_like(feats.data)
                with _ConvTimer(2.0 * N * hw * k * c, True):
                    for img in range(N):
                        b = img % n_clips
                        _conv1x1_tc_raw(ph[img], pl[img] if x3 else None, gt_h[b], gt_l[b] if x3 else None, df[img], hw, 128, c, st)
                feats.add_grad(df)
            if dsn.needs_grad:
                gp_h, gp_l = _operand_planes(gc, n_clips, k, c, 128, False, inv_t, x3, st)   # [n][128][c]
                dprobs = torch.empty((N, hw, 128), device=dev, dtype=torch.float32)
                with _ConvTimer(2.0 * N * hw * k * c, True):
                    for img in range(N):
                        b = img % n_clips
                        _conv1x1_tc_raw(fh[img], fl[img] if x3 else None, gp_h[b], gp_l[b] if x3 else None, dprobs[img], hw, c, 128, st)
                dd = torch.empty_like(dsn.data)
                lib.call("vspw_ocr_region_softmax_bwd", _p(probs), _p(dprobs), 128, _p(dd), _p(sm_ws), N, hw, k, st)
                dsn.add_grad(dd)
            return
        if feats.needs_grad:
            df = torch.empty_like(feats.data)
            for t in range(t_frames):
                pr = probs[t * n_clips:(t + 1) * n_clips]
                # dF[b][p][j] = (1/T) sum_i probs[b][p][i] * g[b][i][j]
                lib.call("vspw_bgemm", _p(pr), _p(g), _p(df[t * n_clips:(t + 1) * n_clips]), n_clips, hw, c, k, hw * k, k, 1,
                         k * c, c, 1, hw * c, c, 1, inv_t, 0.0, st)
            feats.add_grad(df)
        if dsn.needs_grad:
            dprobs = torch.empty_like(probs)
            for t in range(t_frames):
                ft = feats.data[t * n_clips:(t + 1) * n_clips]
                # dP[b][p][i] = (1/T) sum_j feats[b][p][j] * g[b][i][j]
                lib.call("vspw_bgemm", _p(ft), _p(g), _p(dprobs[t * n_clips:(t + 1) * n_clips]), n_clips, hw, k, c, hw * c, c, 1,
                         k * c, 1, c, hw * k, k, 1, inv_t, 0.0, st)
            dd = torch.empty_like(dsn.data)
            if sm_ws is not None:
                lib.call("vspw_ocr_region_softmax_bwd", _p(probs), _p(dprobs), k, _p(dd), _p(sm_ws), N, hw, k, st)
            else:
                lib.call("vspw_softmax_strided_bwd", _p(probs), _p(dprobs), _p(dd), N * k, hw, 1, k, k, hw * k, 1.0, st)
            dsn.add_grad(dd)

    tape.record(backward)
    return out


def object_attention(tape, query, key, value, key_channels):
    """sim = softmax(kc^-0.5 * Q.K^T) over regions; ctx = sim.V (spatial_ocr_block.py:258-275).
    query (n,h,w,kc), key (n,K,1,kc), value (n,K,1,kc) -> (n,h,w,kc)."""
    n, h, w, kc = query.shape
    K = key.shape[1]
    hw = h * w
    dev = query.data.device
    st = _stream()
    scale = float(key_channels) ** -0.5
    needs = tape.grad_enabled and (query.needs_grad or key.needs_grad or value.needs_grad)
    ctx = torch.empty((n, h, w, kc), device=dev, dtype=torch.float32)
    planes = None
    if _ocr_tc_ok() and kc == 256 and K <= 128:
        # ONE tcgen05 kernel: Q.K^T into TMEM, softmax in registers, P as an smem operand, P.V into TMEM; the scores never
        # touch HBM and `sim` is written only when a backward will read it
        x3 = _state["precision"] == "bf16x3"
        qh, ql = _var_planes(query)
        ws = torch.empty(int(lib.dll().vspw_ocr_attention_workspace_bytes(n)), device=dev, dtype=torch.uint8)
        sim = torch.empty((n, hw, K), device=dev, dtype=torch.float32) if needs else None
        chi = torch.empty((n, h, w, kc), device=dev, dtype=torch.bfloat16)
        clo = torch.empty((n, h, w, kc), device=dev, dtype=torch.bfloat16) if x3 else None
        with _ConvTimer(4.0 * n * hw * K * kc, True):
            lib.call("vspw_ocr_attention_fwd_tc", _p(qh), _p(ql), _p(key.data), _p(value.data), _p(ctx), _p(chi), _p(clo), _p(sim), _p(ws),
                     n, hw, K, kc, scale, st)
        planes = (chi, clo)
        attn_tc = (qh, ql, ws, x3)
    else:
        attn_tc = None
        raw = torch.empty((n, hw, K), device=dev, dtype=torch.float32)
        # raw[b][p][i] = sum_c Q[b][p][c] * Key[b][i][c]
        lib.call("vspw_bgemm", _p(query.data), _p(key.data), _p(raw), n, hw, K, kc, hw * kc, kc, 1, K * kc, 1, kc, hw * K, K, 1, 1.0,
                 0.0, st)
        sim = torch.empty_like(raw)
        lib.call("vspw_softmax_strided_fwd", _p(raw), _p(sim), n * hw, K, K, 1, n * hw, 0, scale, st)
        del raw
        lib.call("vspw_bgemm", _p(sim), _p(value.data), _p(ctx), n, hw, kc, K, hw * K, K, 1, K * kc, kc, 1, hw * kc, kc, 1, 1.0, 0.0,
                 st)
    out = Var(ctx, needs_grad=needs)
    out.planes = planes

    def backward():
        g = out.grad
        out.grad = None
        if g is None:
            return
        st = _stream()
        if attn_tc is not None and hw >= 256:
            # tcgen05 backward: dV = sim^T . g and dK = draw^T . Q on the weight-gradient kernel (K axis = pixels), dsim = g . V^T
            # and dQ = draw . K as per-image 1x1 GEMMs on the conv kernel, the softmax backward writing operand planes directly
            qh, ql, ws, x3 = attn_tc
            wsv = ws.view(torch.bfloat16).view(4, n, 128, kc)  # k_hi, k_lo, v_hi, v_lo operand planes of the forward
            gh, gl = _planes_of(g)
            sh = torch.empty((n, hw, 128), device=dev, dtype=torch.bfloat16)
            sl = torch.empty((n, hw, 128), device=dev, dtype=torch.bfloat16) if x3 else None
            lib.call("vspw_ocr_region_planes", _p(sim), _p(sh), _p(sl), n * hw, K, 1.0, st)
            with _ConvTimer(8.0 * n * hw * K * kc, True):
                if value.needs_grad:
                    dv = torch.empty_like(value.data)
                    lib.call("vspw_ocr_gather_tc", _p(sh), _p(sl), _p(gh), _p(gl), _p(dv), 1, n, hw, K, kc, st)
                    value.add_grad(dv)
                if query.needs_grad or key.needs_grad:
                    dsim = torch.empty((n, hw, 128), device=dev, dtype=torch.float32)
                    for b in range(n):
                        _conv1x1_tc_raw(gh[b], gl[b] if x3 else None, wsv[2, b], wsv[3, b] if x3 else None, dsim[b], hw, kc, 128, st)
                    dh = torch.empty((n, hw, 128), device=dev, dtype=torch.bfloat16)
                    dl = torch.empty((n, hw, 128), device=dev, dtype=torch.bfloat16) if x3 else None
                    lib.call("vspw_ocr_attn_softmax_bwd_planes", _p(sim), _p(dsim), _p(dh), _p(dl), n * hw, K, scale, st)
                    if query.needs_grad:
                        kt_h, kt_l = _operand_planes(key.data.view(n, K, kc), n, K, kc, 128, True, 1.0, x3, st)  # [n][kc][128]
                        dq = torch.empty_like(query.data)
                        for b in range(n):
                            _conv1x1_tc_raw(dh[b], dl[b] if x3 else None, kt_h[b], kt_l[b] if x3 else None, dq[b], hw, 128, kc, st)
                        query.add_grad(dq)
                    if key.needs_grad:
                        dk = torch.empty_like(key.data)
                        lib.call("vspw_ocr_gather_tc", _p(dh), _p(dl), _p(qh), _p(ql), _p(dk), 1, n, hw, K, kc, st)
                        key.add_grad(dk)
            return
        if value.needs_grad:
            dv = torch.empty_like(value.data)
            # dV[b][i][c] = sum_p sim[b][p][i] * g[b][p][c]
            lib.call("vspw_bgemm", _p(sim), _p(g), _p(dv), n, K, kc, hw, hw * K, 1, K, hw * kc, kc, 1, K * kc, kc, 1, 1.0, 0.0, st)
            value.add_grad(dv)
        if query.needs_grad or key.needs_grad:
            dsim = torch.empty_like(sim)
            # dsim[b][p][i] = sum_c g[b][p][c] * V[b][i][c]
            lib.call("vspw_bgemm", _p(g), _p(value.data), _p(dsim), n, hw, K, kc, hw * kc, kc, 1, K * kc, 1, kc, hw * K, K, 1, 1.0,
                     0.0, st)
            draw = torch.empty_like(sim)
            lib.call("vspw_softmax_strided_bwd", _p(sim), _p(dsim), _p(draw), n * hw, K, K, 1, n * hw, 0, scale, st)
            if query.needs_grad:
                dq = torch.empty_like(query.data)
                # dQ[b][p][c] = sum_i draw[b][p][i] * Key[b][i][c]
                lib.call("vspw_bgemm", _p(draw), _p(key.data), _p(dq), n, hw, kc, K, hw * K, K, 1, K * kc, kc, 1, hw * kc, kc, 1,
                         1.0, 0.0, st)
                query.add_grad(dq)
            if key.needs_grad:
                dk = torch.empty_like(key.data)
                # dKey[b][i][c] = sum_p draw[b][p][i] * Q[b][p][c]
                lib.call("vspw_bgemm", _p(draw), _p(query.data), _p(dk), n, K, kc, hw, hw * K, 1, K, hw * kc, kc, 1, K * kc, kc, 1,
                         1.0, 0.0, st)
                key.add_grad(dk)

    tape.record(backward)
    return out


def mean_over_stack(tape, items):
    """torch.mean(torch.cat(items, dim=0), dim=0) for equally shaped Vars (memory bank, :122-125)."""
    out_t = torch.empty_like(items[0].data)
    inv = 1.0 / len(items)
    for i, it in enumerate(items):
        lib.call("vspw_axpby", _p(it.data), _p(out_t), inv, 0.0 if i == 0 else 1.0, out_t.numel(), _stream())
    out = Var(out_t, needs_grad=tape.grad_enabled and any(i.needs_grad for i in items))

    def backward():
        g = out.grad
        out.grad = None
        if g is None:
            return
        for it in items:
            if it.needs_grad:
                d = torch.empty_like(g)
                lib.call("vspw_axpby", _p(g), _p(d), inv, 0.0, g.numel(), _stream())
                it.add_grad(d)

    tape.record(backward)
    return out


def dropout2d_mask(p, n, c, device, training):
    """Per-(image, channel) keep mask scaled by 1/(1-p) (nn.Dropout2d); None when inactive.
    The Bernoulli draw is host-side plumbing (torch RNG); it is applied inside vspw_bn_act_fwd."""
    if not training or p <= 0.0:
        return None
    keep = torch.empty((n, c), device=device, dtype=torch.float32).bernoulli_(1.0 - p)
    lib.call("vspw_axpby", _p(keep), _p(keep), 1.0 / (1.0 - p), 0.0, keep.numel(), _stream())
    return keep


# ------------------------------------------------------------------------------------------------
class _GraphFunction(torch.autograd.Function):
    """One autograd node for a whole tape: forward runs `runner(tape)`, backward replays the tape."""

    @staticmethod
    def forward(ctx, runner, params, grad_on, *param_tensors):
        tape = Tape(grad_enabled=grad_on)
        outs, seed = runner(tape)
        if not grad_on:
            tape.release()  # no backward will come: the per-step accumulator arena is free again
        ctx.tape = tape
        ctx.seed = seed
        ctx.params = params
        ctx.mark_non_differentiable(*[o for o in outs[1:]])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        tape, params = ctx.tape, ctx.params
        if tape is None or not tape.grad_enabled:
            raise VspwError("backward called on a graph that was run without gradients (or twice)")
        ctx.seed(gouts[0])
        tape.backward()
        grads = []
        for p in params:
            pv = tape._params.get(id(p))
            g = pv.grad if pv is not None else None
            if g is not None and _grad_sink is not None:
                _grad_sink.mark(p)
            if g is not None and pv.sunk:
                g = None  # already in param.grad's storage (the gradient bucket): nothing for autograd to accumulate
            if g is not None and g.shape != p.shape:
                g = g.view(p.shape)
            grads.append(g)
        ctx.tape = None
        return (None, None, None, *grads)


def run_graph(module, runner):
    """Execute `runner(tape) -> (outputs, seed_fn)` as a single autograd node over module's parameters."""
    params = [p for p in module.parameters()]
    grad_on = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _GraphFunction.apply(runner, params, grad_on, *params)
