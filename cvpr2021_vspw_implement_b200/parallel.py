"""One process per GPU data parallelism for the clip batch (replaces the reference's single-process
``nn.DataParallel`` + SyncBN thread rendez-vous, train_clip2.py:359-364, models/sync_batchnorm/comm.py:46-137).

Clips shard over ranks (all T frames of a clip stay together); parameters live replicated per rank, so the
reference's per-step 282 MB parameter broadcast and scatter/gather disappear.  What remains:

  * gradient averaging (`GradBucket`): every parameter's ``.grad`` IS a view of one flat buffer that the backward kernels
    write into directly (weight-gradient kernels, BN backward, bias sums: `engine.set_grad_sink`), so there is no pack /
    unpack pass; one all-reduce (NCCL AVG over NVLink on GPUs, gloo on CPU tests) per step — DataParallel's "mean over
    replicas of per-replica mean losses" (train_clip2.py:98);
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

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        self.views = None
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._marked = set()   # parameter indices that received a gradient from an engine graph since zero_grad()
        self._armed = False    # zero_grad() was called: p.grad are bucket views

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
        from . import engine as E
        E.set_grad_sink(self)

    # -- engine-facing sink protocol ------------------------------------------------------------------------------
    def destination(self, param):
        """The tensor the FIRST gradient contribution of `param` may be written into (overwriting zeros), or None."""
        i = self._index.get(id(param))
        if i is None or not self._armed or i in self._marked or param.grad is not self.views[i]:
            return None  # (a second graph of the same step accumulates through autograd instead of overwriting)
        self._marked.add(i)
        return self.views[i]

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
        world = dist.get_world_size()
        self.finish_step()
        have = [p for p in self.params if p.grad is not None]
        if not have:
            dev = self.flat.device if self.flat is not None else None
            if dev is None:
                raise RuntimeError("GradBucket.all_reduce_mean: no parameter has a gradient and the bucket was never used")
        else:
            dev = have[0].grad.device
        self._ensure(dev)
        stray_v, stray_g = [], []
        missing = []
        for i, (v, p) in enumerate(zip(self.views, self.params)):
            if p.grad is None:
                v.zero_()  # this rank contributes zero
                missing.append(i)
                continue
            if p.grad is not v and p.grad.data_ptr() != v.data_ptr():
                stray_v.append(v)
                stray_g.append(p.grad)
        if stray_v:
            torch._foreach_copy_(stray_v, stray_g)  # legacy protocol: pack as one multi-tensor launch
        self.tail.fill_(1.0)  # device-side: the host must stay ahead of the stream here
        if missing:
            self.tail[torch.tensor(missing, device=dev)] = 0.0
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / world)
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
