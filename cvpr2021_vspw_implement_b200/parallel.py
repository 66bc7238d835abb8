"""One process per GPU data parallelism for the clip batch (replaces the reference's single-process
``nn.DataParallel`` + SyncBN thread rendez-vous, train_clip2.py:359-364, models/sync_batchnorm/comm.py:46-137).

Clips shard over ranks (all T frames of a clip stay together); parameters live replicated per rank, so the
reference's per-step 282 MB parameter broadcast and scatter/gather disappear.  What remains:

  * gradient averaging (`GradBucket`): every parameter's ``.grad`` IS a view of one flat buffer that the backward kernels
    write into directly (weight-gradient kernels, BN backward, bias sums: `engine.set_grad_sink`), so there is no pack /
    unpack pass; the buffer is all-reduced (NCCL AVG over NVLink on GPUs, gloo on CPU tests) — DataParallel's "mean over
    replicas of per-replica mean losses" (train_clip2.py:98) — in a few contiguous chunks, each launched on a side stream
    as soon as the backward pass has enqueued the last kernel that writes into it (the pattern is learned in the first
    step), so only the small chunk of the stem/layer1/layer2 gradients is reduced after the backward pass has ended;
  * SyncBN statistics (`PeerSums`): a one-shot exchange over NVLink peer memory per BN layer (csrc/peer.cu) instead of a
    library all-reduce per layer — the reference's multi-GPU BN semantics (sync_batchnorm/batchnorm.py:110-131);
  * scalar loss/acc averaging for logging.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None, device_offset=0):
    """Initialise torch.distributed from torchrun's environment; returns (world, rank, local_rank).
    `device_offset` (the entry points' --start_gpu): this rank computes on cuda:(device_offset + LOCAL_RANK); the same
    device is made current AND bound to the NCCL communicator."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            dev = torch.device("cuda", device_offset + local)
            torch.cuda.set_device(dev)
            kw["device_id"] = dev
        dist.init_process_group(backend, **kw)
    return world, rank, local


def is_parallel():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_clips(n_clips_global, world, rank):
    """[lo, hi) clip range of this rank; the global batch must divide evenly (DataLoader drop_last=True upstream)."""
    if n_clips_global % world:
        raise ValueError(f"global batch of {n_clips_global} clips does not divide over {world} ranks")
    per = n_clips_global // world
    return rank * per, (rank + 1) * per


class GradBucket:
    """All gradients of a module in ONE flat fp32 buffer.

    `zero_grad()` zeroes the buffer with one memset, points every ``p.grad`` at its slice and registers the bucket as the
    engine's gradient sink: the tape's backward kernels then write each parameter gradient straight into its slice (no
    per-parameter tensor, no pack/unpack), and `FusedSGD` reads the slices in place.  `all_reduce_mean()` averages the whole
    buffer over the ranks with one collective.  Parameters that received no gradient in a step end the step with
    ``p.grad = None`` (as on a single device and under the reference's DataParallel: the optimizer skips them).

    The legacy protocol still works: gradients produced without `zero_grad()` (plain tensors, possibly None on some ranks)
    are packed into the buffer by `all_reduce_mean()`, and a parameter nobody touched stays None."""

    def __init__(self, params, overlap=None, chunk_elems=None):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        self.views = None
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._marked = set()   # parameter indices that received a gradient from an engine graph since zero_grad()
        self._armed = False    # zero_grad() was called: p.grad are bucket views
        # -- overlap of the all-reduce with the backward pass (module docstring) --------------------------------------
        if overlap is None:
            overlap = os.environ.get("VSPW_GRAD_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        self._offsets, off = [], 0
        for p in self.params:
            self._offsets.append(off)
            off += p.numel()
        self._chunks = self._make_chunks(chunk_elems)   # [(first param, one past the last param, lo, hi)] in PARAMETER order
        self._chunk_of = [0] * len(self.params)
        for c, (a, b, _, _) in enumerate(self._chunks):
            for i in range(a, b):
                self._chunk_of[i] = c
        self._pattern = None     # parameter indices whose gradient arrived through destination() in the previous step
        self._sunk = set()       # ... in this step
        self._left = None        # per chunk: expected parameters still missing in this step
        self._ready = []         # chunks complete except for the launch that was about to be issued when they completed
        self._launched = set()
        self._side = None
        self._pg = None
        self.last_overlapped = 0

    def _make_chunks(self, chunk_elems):
        """Contiguous parameter ranges.  The backward pass fills the buffer from its END towards its start, so the chunk at
        the start (stem, layer1, layer2: done last, its reduction is exposed) is small and the later ones grow."""
        if chunk_elems is None:
            chunk_elems = [int(x) for x in os.environ.get("VSPW_GRAD_CHUNKS", "2097152,8388608,16777216").split(",")]
        elif isinstance(chunk_elems, int):
            chunk_elems = [chunk_elems]
        chunks, a, lo, k = [], 0, 0, 0
        for i, p in enumerate(self.params):
            hi = self._offsets[i] + p.numel()
            cap = chunk_elems[min(k, len(chunk_elems) - 1)]
            if hi - lo >= cap or i == len(self.params) - 1:
                chunks.append((a, i + 1, lo, hi))
                a, lo, k = i + 1, hi, k + 1
        return chunks

    def _ensure(self, device):
        if self.flat is None or self.flat.device != device:
            # tail: one float per parameter = "some rank has a gradient for it" (reduced together with the gradients)
            self.flat = torch.zeros(self.numel + len(self.params), device=device, dtype=torch.float32)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
            self.tail = self.flat[self.numel:]

    def zero_grad(self):
        """Replaces ``module.zero_grad()`` in the step loop."""
        if not self.params:
            return
        self._ensure(self.params[0].device)
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v
        self._marked.clear()
        self._armed = True
        self._begin_overlap()
        from . import engine as E
        E.set_grad_sink(self)

    # -- overlap ----------------------------------------------------------------------------------------------------------
    def reset_overlap(self):
        """Forget the learned gradient pattern (call when the graph that produces the gradients changes)."""
        self._pattern = None

    def _begin_overlap(self):
        self._sunk = set()
        self._ready = []
        self._launched = set()
        self._left = None
        if self.overlap and is_parallel() and self._pattern is not None:
            self._left = [0] * len(self._chunks)
            for i in self._pattern:
                self._left[self._chunk_of[i]] += 1

    def _reduce(self, t):
        if dist.get_backend(self._pg) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self._pg)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._pg)
            t.mul_(1.0 / dist.get_world_size())

    def _launch(self, lo, hi):
        """All-reduce flat[lo:hi] behind everything enqueued so far on the current stream, without blocking that stream."""
        t = self.flat[lo:hi]
        if not t.is_cuda:
            self._reduce(t)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=t.device)
            ctas = int(os.environ.get("VSPW_GRAD_NCCL_CTAS", "0"))
            if ctas > 0 and dist.get_backend() == "nccl":
                # its own communicator with few CTAs: the reduction shares the SMs with the persistent backward kernels
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = ctas
                self._pg = dist.new_group(backend="nccl", pg_options=opts)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(t.device))
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            self._reduce(t)

    def _flush_ready(self):
        for c in self._ready:
            _, _, lo, hi = self._chunks[c]
            self._launch(lo, hi)
            self._launched.add(c)
        self._ready = []

    # -- engine-facing sink protocol ------------------------------------------------------------------------------
    def destination(self, param):
        """The tensor the FIRST gradient contribution of `param` may be written into (overwriting zeros), or None."""
        i = self._index.get(id(param))
        if i is None or not self._armed or i in self._marked or param.grad is not self.views[i]:
            if i is not None and self._chunk_of[i] in self._launched:
                raise RuntimeError("GradBucket: a second gradient contribution arrived for a parameter whose chunk is already being "
                                   "all-reduced; build the bucket with overlap=False for steps with several graphs")
            return None  # (a second graph of the same step accumulates through autograd instead of overwriting)
        self._marked.add(i)
        self._sunk.add(i)
        if self._left is not None:
            c = self._chunk_of[i]
            if c in self._launched or c in self._ready:
                raise RuntimeError("GradBucket: the set of parameters that receive gradients changed since the previous step; "
                                   "call reset_overlap() when switching graphs")
            if i in self._pattern:
                self._left[c] -= 1
                if self._left[c] == 0:
                    self._ready.append(c)  # launched by node_done(): this parameter's kernel is not enqueued yet
        return self.views[i]

    def node_done(self):
        """The engine finished enqueueing the backward kernels of one tape node on the current stream."""
        if self._ready:
            self._flush_ready()

    def late(self, param):
        """The engine is about to ACCUMULATE a further contribution into a parameter it already wrote (weight used twice)."""
        i = self._index.get(id(param))
        if i is None:
            return
        c = self._chunk_of[i]
        if c in self._ready:
            self._ready.remove(c)
            self._left[c] = -1  # reduced at the end of the step
        elif c in self._launched:
            raise RuntimeError("GradBucket: gradient accumulated into a chunk that is already being all-reduced")

    def mark(self, param):
        i = self._index.get(id(param))
        if i is not None:
            self._marked.add(i)

    def finish_step(self):
        """After backward on one device: parameters no graph produced a gradient for get ``grad = None``."""
        if self._armed and self._marked:
            for i, p in enumerate(self.params):
                if i not in self._marked and p.grad is self.views[i]:
                    p.grad = None
        self._armed = False

    def all_reduce_mean(self):
        if not is_parallel():
            self.finish_step()
            return
        armed = self._armed
        self.finish_step()
        have = [p for p in self.params if p.grad is not None]
        if not have:
            dev = self.flat.device if self.flat is not None else None
            if dev is None:
                raise RuntimeError("GradBucket.all_reduce_mean: no parameter has a gradient and the bucket was never used")
        else:
            dev = have[0].grad.device
        self._ensure(dev)
        self._flush_ready()
        stray_v, stray_g = [], []
        missing = []
        for i, (v, p) in enumerate(zip(self.views, self.params)):
            if p.grad is None:
                if self._chunk_of[i] not in self._launched:
                    v.zero_()  # this rank contributes zero (a launched chunk already holds the average of zeros)
                missing.append(i)
                continue
            if p.grad is not v and p.grad.data_ptr() != v.data_ptr():
                if self._chunk_of[i] in self._launched:
                    raise RuntimeError("GradBucket: a gradient outside the bucket belongs to a chunk that is already all-reduced")
                stray_v.append(v)
                stray_g.append(p.grad)
        if stray_v:
            torch._foreach_copy_(stray_v, stray_g)  # legacy protocol: pack as one multi-tensor launch
        self.tail.fill_(1.0)  # device-side: the host must stay ahead of the stream here
        if missing:
            self.tail[torch.tensor(missing, device=dev)] = 0.0
        # what is left: maximal runs of chunks not launched during the backward pass; the flags travel with the last run
        runs, c = [], len(self._chunks) - 1
        while c >= 0:
            if c in self._launched:
                c -= 1
                continue
            hi = self._chunks[c][3]
            while c - 1 >= 0 and (c - 1) not in self._launched:
                c -= 1
            runs.append([self._chunks[c][2], hi])
            c -= 1
        if runs and runs[0][1] == self.numel:
            runs[0][1] = self.flat.numel()
        else:
            runs.append([self.numel, self.flat.numel()])
        for lo, hi in runs:
            self._launch(lo, hi)
        if self._side is not None and self.flat.is_cuda:
            torch.cuda.current_stream(dev).wait_stream(self._side)
        if self.overlap and armed:
            self._pattern = frozenset(self._sunk)  # the next step reduces each chunk as soon as these have all arrived
        self.last_overlapped = len(self._launched)  # chunks whose all-reduce was launched during the backward pass
        self._launched = set()
        self._left = None
        for v, p in zip(self.views, self.params):
            if p.grad is not None and p.grad is not v:
                p.grad = v  # adopt the reduced slice: no unpack copy
        if missing:
            got = self.tail.cpu()  # (a host sync, only on the rare step where this rank lacks a gradient another rank has)
            for i in missing:
                if float(got[i]) > 0.0:
                    self.params[i].grad = self.views[i]


class PeerSums:
    """SyncBN statistics exchange over NVLink peer memory (csrc/peer.cu): `all_reduce_sums(t)` sums a small fp64 vector over
    the ranks of one node with ONE single-block kernel — every rank pushes its vector into an inbox on each peer, raises a
    flag, waits for the peers' flags in its own memory and adds the vectors in rank order (bit-identical totals).
    Set-up (collective): each rank cudaMallocs an inbox, the CUDA IPC handles travel through torch.distributed's
    all_gather_object, peers map each other's inboxes.  Pass it to `engine.set_syncbn(True, group=PeerSums())`."""

    RING = 4

    def __init__(self, max_elems=8192):
        from ._lib import lib
        if not is_parallel():
            raise RuntimeError("PeerSums needs an initialised torch.distributed group with more than one rank")
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 16:
            raise ValueError("PeerSums: at most 16 ranks (one NVLink domain)")
        self.max_elems = int(max_elems)
        self._lib = lib
        dll = lib.dll()
        nbytes = int(dll.vspw_peer_inbox_bytes(self.world, self.RING, self.max_elems))
        mine = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        lib.call("vspw_peer_alloc", nbytes, ctypes.byref(mine), handle)
        self._mine = mine
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        self._opened = []
        bases = []
        for r, h in enumerate(handles):
            if r == self.rank:
                bases.append(mine.value)
                continue
            buf = (ctypes.c_uint8 * 64).from_buffer_copy(h)
            ptr = ctypes.c_void_p()
            lib.call("vspw_peer_open", buf, ctypes.byref(ptr))
            self._opened.append(ptr)
            bases.append(ptr.value)
        self._bases = (ctypes.c_uint64 * self.world)(*bases)
        self.seq = 0
        # VSPW_SYNCBN_FUSED=0: one stand-alone exchange launch per BN layer instead of the exchange in the BN kernels' prologue
        self.fused = os.environ.get("VSPW_SYNCBN_FUSED", "1") != "0"
        dist.barrier()  # every inbox is mapped everywhere before the first exchange

    def size(self):
        return self.world

    def next_ctx(self, n_elems):
        """The `vspw_peer_ctx` of the NEXT exchange, for the BN kernels that run it in their own prologue
        (vspw_bn_train_fwd_sync / vspw_bn_bwd_apply_sync); None when the vector does not fit one inbox slot."""
        from ._lib import PeerCtx
        if not self.fused or n_elems > self.max_elems:
            return None
        self.seq += 1
        ctx = PeerCtx()
        for i in range(self.world):
            ctx.inbox[i] = self._bases[i]
        ctx.world, ctx.rank, ctx.ring, ctx.max_elems, ctx.seq = self.world, self.rank, self.RING, self.max_elems, self.seq
        return ctx

    def all_reduce_sums(self, t):
        if t.dtype != torch.float64 or not t.is_contiguous():
            raise ValueError("PeerSums.all_reduce_sums: contiguous fp64 tensor expected")
        n, off = t.numel(), 0
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        base = t.data_ptr()
        while off < n:
            m = min(self.max_elems, n - off)
            self.seq += 1
            self._lib.call("vspw_peer_allreduce_f64", ctypes.c_void_p(base + 8 * off), m, self._bases, self.world, self.rank,
                           ctypes.c_uint64(self.seq), self.RING, self.max_elems, st)
            off += m

    def close(self):
        """Collective: unmap the peers' inboxes and free this rank's."""
        if self._mine is None:
            return
        torch.cuda.synchronize()
        if is_parallel():
            dist.barrier()
        for p in self._opened:
            self._lib.call("vspw_peer_close", p)
        self._opened = []
        if is_parallel():
            dist.barrier()
        self._lib.call("vspw_peer_free", self._mine)
        self._mine = None


def make_syncbn_group(kind="peer"):
    """The statistics-exchange object for `engine.set_syncbn`: 'peer' = PeerSums (NVLink peer memory, single node),
    'nccl' = one torch.distributed all-reduce per BN layer (`engine.TorchDistGroup`; the round-1 path, kept for A/B runs)."""
    from . import engine as E
    if kind == "peer" and torch.cuda.is_available() and dist.get_backend() == "nccl":
        return PeerSums()
    return E.TorchDistGroup()


def mean_scalar(t):
    """Average a 0-d tensor over ranks (the DataParallel gather + .mean() of train_clip2.py:98-99)."""
    if not is_parallel():
        return t
    t = t.detach().clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t / dist.get_world_size()


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s weights and buffers (DataParallel replicates device 0's module)."""
    if not is_parallel():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)
