"""FusedSGD — the reference's optimizer (torch.optim.SGD with momentum, per-group lr and weight decay; train_clip2.py:215-252)
as ONE kernel launch per step (vspw_sgd_momentum_step) instead of ~50 multi-tensor ATen launches.

Drop-in for `torch.optim.SGD(params, lr, momentum, weight_decay)`: same `param_groups` (so `adjust_learning_rate` works
unchanged), same `state[p]['momentum_buffer']` and therefore the same `state_dict()` layout as `torch.optim.SGD` over the same
(de-duplicated) parameter groups.  A REFERENCE `opt_epoch_E.pth` repeats every parameter 2-5 times per group (quirk Q10):
`train_clip2.remap_reference_optimizer_state` maps it onto these groups on resume.  Dampening and Nesterov are not
supported (the reference uses neither).
"""
import ctypes

import numpy as np
import torch

from ._lib import lib

_ENTRY = np.dtype([("p", "<u8"), ("g", "<u8"), ("buf", "<u8"), ("n", "<u8"), ("lr", "<f4"), ("wd", "<f4")])
assert _ENTRY.itemsize == 40


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0, dampening=0, nesterov=False):
        if dampening != 0 or nesterov:
            raise NotImplementedError("FusedSGD implements the reference's settings: dampening=0, nesterov=False")
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("lr, momentum and weight_decay must be non-negative")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0, nesterov=False))
        self._chunk = None
        # ring of pinned staging buffers [table | block_tensor | block_chunk] with their device twins: the host runs steps
        # ahead of the device, so a slot is rewritten only after the copy that read it has completed (event)
        self._ring = []
        self._slot = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._chunk is None:
            self._chunk = int(lib.dll().vspw_sgd_chunk_elems())
        by_momentum = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedSGD: parameters must live on a CUDA device (the engine has no CPU path)")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedSGD: fp32 contiguous parameters only")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                by_momentum.setdefault((float(group["momentum"]), p.device), []).append(
                    (p, g, st["momentum_buffer"], float(group["lr"]), float(group["weight_decay"])))
        for (momentum, dev), items in by_momentum.items():
            self._launch(items, momentum, dev)
        return loss

    def _launch(self, items, momentum, dev):
        n_t = len(items)
        table = np.empty(n_t, dtype=_ENTRY)
        blocks_t, blocks_c = [], []
        for i, (p, g, buf, lr, wd) in enumerate(items):
            n = p.numel()
            table[i] = (p.data_ptr(), g.data_ptr(), buf.data_ptr(), n, lr, wd)
            nb = (n + self._chunk - 1) // self._chunk
            blocks_t.append(np.full(nb, i, dtype=np.uint32))
            blocks_c.append(np.arange(nb, dtype=np.uint32))
        bt, bc = np.concatenate(blocks_t), np.concatenate(blocks_c)
        nbytes = table.nbytes + bt.nbytes + bc.nbytes
        if len(self._ring) < 4:
            host = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
            self._ring.append([host, torch.empty(host.numel(), dtype=torch.uint8, device=dev), None])
            slot = self._ring[-1]
        else:
            slot = self._ring[self._slot % len(self._ring)]
            self._slot += 1
            if slot[2] is not None:
                slot[2].synchronize()
            if slot[0].numel() < nbytes or slot[1].device != dev:
                slot[0] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                slot[1] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        host, devbuf = slot[0], slot[1]
        h = host.numpy()
        o1, o2 = table.nbytes, table.nbytes + bt.nbytes
        h[:o1] = table.view(np.uint8)
        h[o1:o2] = bt.view(np.uint8)
        h[o2:nbytes] = bc.view(np.uint8)
        devbuf[:nbytes].copy_(host[:nbytes], non_blocking=True)
        base = devbuf.data_ptr()
        lib.call("vspw_sgd_momentum_step", ctypes.c_void_p(base), ctypes.c_void_p(base + o1), ctypes.c_void_p(base + o2), int(bt.size),
                 float(momentum), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        slot[2] = torch.cuda.Event()
        slot[2].record()
        # the gradient tensors named in the table outlive the launch: p.grad holds them until the caller's zero_grad
