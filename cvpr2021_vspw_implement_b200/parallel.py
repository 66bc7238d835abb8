"""One process per GPU data parallelism for the clip batch (replaces the reference's single-process
``nn.DataParallel`` + SyncBN thread rendez-vous, train_clip2.py:359-364, models/sync_batchnorm/comm.py:46-137).

Clips shard over ranks (all T frames of a clip stay together); parameters live replicated per rank, so the
reference's per-step 282 MB parameter broadcast and scatter/gather disappear.  What remains:

  * gradient averaging: one flat bucket, one all-reduce (NCCL over NVLink on GPUs, gloo on CPU tests), matching
    DataParallel's "mean over replicas of per-replica mean losses" (train_clip2.py:98);
  * optional SyncBN statistics all-reduce (``engine.set_syncbn``), the reference's multi-GPU BN semantics;
  * scalar loss/acc averaging for logging.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (world, rank, local_rank)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
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
    """Flat gradient bucket: parameters' ``.grad`` are packed into one contiguous buffer, all-reduced once and
    averaged; the buffer is allocated once and reused every step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def all_reduce_mean(self):
        if not is_parallel():
            return
        world = dist.get_world_size()
        ref = next(p for p in self.params if p.grad is not None)
        if self.flat is None or self.flat.device != ref.grad.device:
            self.flat = torch.zeros(self.numel, device=ref.grad.device, dtype=torch.float32)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()  # a rank whose shard did not touch this parameter contributes zero
        # pack / unpack as two multi-tensor launches instead of one tiny copy kernel per parameter
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / world)
        torch._foreach_copy_([g for _, g in have], [v for v, _ in have])
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.clone()


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
