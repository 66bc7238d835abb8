"""GPU: SyncBN (cross-rank BN statistics) and the data-parallel step against the single-device global batch.

Reference semantics: `_SynchronizedBatchNorm.forward` under DataParallel (models/sync_batchnorm/batchnorm.py:75-131) sums
sum(x), sum(x^2) over all replicas, normalises every replica with the global statistics and back-propagates through them;
the loss is the mean of the per-replica mean losses (train_clip2.py:98).  So a 2-rank step on clips [0,1] | [2,3] must equal
ONE device running all four clips: loss, logits, every parameter gradient (after the all-rank average), running statistics.

Two ways to get two ranks:
  * `test_two_emulated_ranks_*`: two host threads on ONE GPU, each with its own replica and tape, joined by a test double
    of the statistics exchange (`engine.set_syncbn(group=...)`).  Runs in the ordinary 1-GPU `-m gpu` tier and exercises
    all of the engine's SyncBN arithmetic (global count, all-rank sums in dx, 1/world on the gamma/beta gradients).
  * `test_two_nccl_ranks_*`: two real processes over NCCL + the peer-memory exchange (`parallel.PeerSums`), launched with
    torchrun; skipped when the box has one GPU.
"""
import os
import subprocess
import sys
import threading

import pytest
import torch

import cases as C
import tcb_oracle as O

pytestmark = pytest.mark.gpu
T, H, W = 3, 65, 97


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cvpr2021_vspw_implement_b200 import engine
    return engine


class ThreadGroup:
    """Test double of the statistics exchange: `n` host threads meet at a barrier and sum their tensors."""

    def __init__(self, n):
        self.n = n
        self.barrier = threading.Barrier(n)
        self.slots = [None] * n
        self.tls = threading.local()
        self.calls = 0

    def size(self):
        return self.n

    def all_reduce_sums(self, t):
        r = self.tls.rank
        torch.cuda.synchronize()
        self.slots[r] = t.clone()
        self.barrier.wait()
        total = self.slots[0].clone()
        for s in self.slots[1:]:
            total += s
        torch.cuda.synchronize()
        self.barrier.wait()
        t.copy_(total)
        if r == 0:
            self.calls += 1


def test_peer_exchange_kernel_three_ranks_on_one_device(E):
    """csrc/peer.cu on ONE GPU: three 'ranks' = three inboxes of this process, one stream each (the kernels of one exchange
    must be co-resident, as they are on three GPUs).  Integer-valued fp64 vectors: the totals are exact and must be identical
    on every rank; 12 exchanges wrap the 4-slot ring three times."""
    import ctypes
    from cvpr2021_vspw_implement_b200._lib import lib
    os.environ.setdefault("VSPW_PEER_TIMEOUT_S", "20")
    world, ring, max_elems, n = 3, 4, 4096, 3000
    dll = lib.dll()
    nbytes = int(dll.vspw_peer_inbox_bytes(world, ring, max_elems))
    ptrs = []
    for _ in range(world):
        p, h = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
        lib.call("vspw_peer_alloc", nbytes, ctypes.byref(p), h)
        ptrs.append(p)
    bases = (ctypes.c_uint64 * world)(*[p.value for p in ptrs])
    streams = [torch.cuda.Stream() for _ in range(world)]
    base = torch.arange(n, device="cuda", dtype=torch.float64) + 1
    vecs = [base * (r + 1) for r in range(world)]
    torch.cuda.synchronize()
    try:
        for seq in range(1, 13):
            for r in range(world):
                lib.call("vspw_peer_allreduce_f64", ctypes.c_void_p(vecs[r].data_ptr()), n, bases, world, r, ctypes.c_uint64(seq), ring,
                         max_elems, ctypes.c_void_p(streams[r].cuda_stream))
        torch.cuda.synchronize()
        # after the first exchange every rank holds 6*base; each further exchange multiplies by 3
        want = base * 6 * 3 ** 11
        for r in range(world):
            assert torch.equal(vecs[r], want), r
    finally:
        torch.cuda.synchronize()
        for p in ptrs:
            lib.call("vspw_peer_free", p)


def test_bn_kernels_with_the_exchange_in_their_prologue_two_ranks_on_one_device(E):
    """vspw_bn_train_fwd_sync / vspw_bn_bwd_apply_sync: two 'ranks' = two inboxes + two streams of this process, tensors small
    enough that both one-wave grids are co-resident (on two GPUs each has a device to itself).  Every rank must produce
    exactly what the plain kernels produce from the pre-summed statistics."""
    import ctypes
    from cvpr2021_vspw_implement_b200._lib import PeerCtx, lib
    os.environ.setdefault("VSPW_PEER_TIMEOUT_S", "20")
    world, ring, max_elems, c, pixels = 2, 4, 1024, 64, 1500
    dll = lib.dll()
    nbytes = int(dll.vspw_peer_inbox_bytes(world, ring, max_elems))
    ptrs = []
    for _ in range(world):
        p, h = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
        lib.call("vspw_peer_alloc", nbytes, ctypes.byref(p), h)
        ptrs.append(p)
    g = torch.Generator(device="cuda").manual_seed(3)
    ys = [torch.randn(pixels, c, generator=g, device="cuda") * (r + 1) + r for r in range(world)]
    douts = [torch.randn(pixels, c, generator=g, device="cuda") for _ in range(world)]
    gamma = torch.rand(c, generator=g, device="cuda") + 0.5
    beta = torch.randn(c, generator=g, device="cuda")
    vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    streams = [torch.cuda.Stream() for _ in range(world)]

    def ctx_for(rank, seq):
        ctx = PeerCtx()
        for i in range(world):
            ctx.inbox[i] = ptrs[i].value
        ctx.world, ctx.rank, ctx.ring, ctx.max_elems, ctx.seq = world, rank, ring, max_elems, seq
        return ctx

    def local_sums(y):
        d = y.double()
        return torch.stack([d.sum(0), (d * d).sum(0)]).contiguous()

    try:
        count = float(world * pixels)
        total = sum(local_sums(y) for y in ys)
        outs, means, invstds = [], [], []
        sums_after = [local_sums(ys[r]) for r in range(world)]
        for r in range(world):
            outs.append(torch.empty_like(ys[r])); means.append(torch.empty(c, device="cuda")); invstds.append(torch.empty(c, device="cuda"))
        torch.cuda.synchronize()  # (no host synchronisation between the two launches: rank 0's kernel waits for rank 1's)
        for r in range(world):
            sums, o, mean, invstd = sums_after[r], outs[r], means[r], invstds[r]
            ctx = ctx_for(r, 1)
            lib.call("vspw_bn_train_fwd_sync", vp(ys[r]), vp(sums), count, vp(gamma), vp(beta), 1e-5, 0.1, None, None, vp(mean), vp(invstd), 0,
                     None, None, None, None, 1, vp(o), None, None, None, pixels, c, pixels, ctypes.byref(ctx), ctypes.c_void_p(streams[r].cuda_stream))
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(sums_after[r], sums_after[0]) and C.rel_err(sums_after[r].cpu(), total.cpu()) <= 1e-14
            ref_o = torch.empty_like(ys[r]); rm = torch.empty(c, device="cuda"); ri = torch.empty(c, device="cuda")
            t = sums_after[r]
            lib.call("vspw_bn_train_fwd", vp(ys[r]), vp(t[0]), vp(t[1]), count, vp(gamma), vp(beta), 1e-5, 0.1, None, None, vp(rm), vp(ri), 0,
                     None, None, None, None, 1, vp(ref_o), None, None, None, pixels, c, pixels, None)
            torch.cuda.synchronize()
            assert torch.equal(outs[r], ref_o) and torch.equal(means[r], rm) and torch.equal(invstds[r], ri)
        # backward: (dbeta, dgamma) exchanged in the prologue of the apply pass
        dys, dsums = [], []
        for r in range(world):
            ds = torch.zeros(2, c, device="cuda", dtype=torch.float64)
            lib.call("vspw_bn_bwd_reduce", vp(douts[r]), vp(outs[r]), None, vp(ys[r]), vp(means[r]), vp(invstds[r]), None, 1, None, pixels, c, pixels,
                     vp(ds[0]), vp(ds[1]), None)
            dsums.append(ds)
        torch.cuda.synchronize()
        dtotal = dsums[0] + dsums[1]
        bufs = [(torch.empty_like(ys[r]), torch.empty(c, device="cuda"), torch.empty(c, device="cuda")) for r in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            dy, dg, db = bufs[r]
            ctx = ctx_for(r, 2)
            lib.call("vspw_bn_bwd_apply_sync", vp(douts[r]), vp(outs[r]), None, vp(ys[r]), vp(means[r]), vp(invstds[r]), vp(gamma), None, 1,
                     vp(dsums[r]), vp(dy), None, None, None, vp(dg), vp(db), None, pixels, c, pixels, count, 1.0 / world, ctypes.byref(ctx),
                     ctypes.c_void_p(streams[r].cuda_stream))
            dys.append((dy, dg, db))
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(dsums[r], dtotal)
            ry = torch.empty_like(ys[r]); rg = torch.empty(c, device="cuda"); rb = torch.empty(c, device="cuda")
            lib.call("vspw_bn_bwd_apply", vp(douts[r]), vp(outs[r]), None, vp(ys[r]), vp(means[r]), vp(invstds[r]), vp(gamma), None, 1,
                     vp(dtotal[0]), vp(dtotal[1]), vp(ry), None, None, None, vp(rg), vp(rb), None, pixels, c, pixels, 0, count, 1.0 / world, None)
            torch.cuda.synchronize()
            assert torch.equal(dys[r][0], ry) and torch.equal(dys[r][1], rg) and torch.equal(dys[r][2], rb)
    finally:
        torch.cuda.synchronize()
        for p in ptrs:
            lib.call("vspw_peer_free", p)


def _global_clip(n_clips, seed=41):
    # no ignore labels: every rank then has the same number of valid pixels and the mean of the rank means IS the global mean
    return O.synthetic_clip(T, n_clips, H, W, C.NUM_CLASS, seed=seed, block=16, ignore_frac=0.0)


def _run_single(E, kind, prec, clamp, imgs, labs):
    m = C.no_dropout(C.build(kind, "resnet50dilated", 31).cuda().train())
    E.set_syncbn(False, clamp=clamp)
    with E.precision(prec), E.capturing() as cap:
        loss, acc = m(C.feed(imgs, labs, True, "cuda"))
        loss.backward()
    torch.cuda.synchronize()
    return m, loss.item(), cap["logits"].clone()


@pytest.mark.parametrize("kind,prec,clamp", [("Clip_PSP", "fp32", False), ("Clip_PSP", "bf16x3", True), ("ClipOCRNet", "bf16x3", False)])
def test_two_emulated_ranks_equal_the_single_device_global_batch(E, kind, prec, clamp):
    world, n_loc = 2, 2
    imgs, labs = _global_clip(world * n_loc)
    ref_m, ref_loss, ref_logits = _run_single(E, kind, prec, clamp, imgs, labs)
    group = ThreadGroup(world)
    reps = [C.no_dropout(C.build(kind, "resnet50dilated", 31).cuda().train()) for _ in range(world)]
    out = [None] * world
    errs = []

    def rank_main(r):
        try:
            group.tls.rank = r
            torch.cuda.set_device(0)
            sl = slice(r * n_loc, (r + 1) * n_loc)
            feed = C.feed([i[sl] for i in imgs], [l[sl] for l in labs], True, "cuda")
            loss, acc = reps[r](feed)
            # the tape's backward in THIS thread (torch's autograd engine would run both ranks on one worker thread)
            grads = loss.grad_fn.apply(torch.ones_like(loss), torch.zeros_like(acc))[3:]
            torch.cuda.synchronize()
            out[r] = (loss.item(), grads)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            group.barrier.abort()

    E.set_syncbn(True, clamp=clamp, group=group)
    try:
        with E.precision(prec):
            ths = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
    finally:
        E.set_syncbn(False, clamp=False)
    if errs:
        raise errs[0]
    assert group.calls >= 2 * 50, "every train-mode BN layer exchanges its sums forward and backward"
    loss2 = sum(o[0] for o in out) / world
    assert abs(loss2 - ref_loss) <= 2e-5 * abs(ref_loss), (loss2, ref_loss)
    params = [p for p in reps[0].parameters()]
    worst = (0.0, "")
    checked = 0
    for i, (name, p_ref) in enumerate(ref_m.named_parameters()):
        g_ref = p_ref.grad
        gs = [o[1][i] for o in out]
        if g_ref is None:
            assert all(g is None for g in gs), name
            continue
        g = sum(x.double() for x in gs) / world  # GradBucket.all_reduce_mean
        rn = float(g_ref.double().norm())
        if rn < 1e-7 or (name.endswith(".0.bias") or name.endswith(".3.bias")) and "f_" in name or name in ("conv_3x3.0.bias", "dsn_head.0.bias",
                                                                                                   "spatial_ocr_head.conv_bn_dropout.0.bias"):
            continue  # conv biases in front of a train-mode BN: the true gradient is exactly zero, both sides hold rounding noise
        e = float((g - g_ref.double()).norm() / rn)
        checked += 1
        if e > worst[0]:
            worst = (e, name)
        if name.endswith(("bn1.weight", "bn1.bias", "bn3.weight", "bn3.bias", ".1.weight", ".1.bias")):
            # BN affine parameters: the advisor's round-1 finding was a factor `world` here
            assert e <= (1e-4 if prec == "fp32" else 5e-2), (name, e)
    print(f"{kind}/{prec}/clamp={clamp}: loss {loss2:.6f} vs {ref_loss:.6f}; {checked} gradient tensors, worst rel-L2 {worst[0]:.2e} ({worst[1]})")
    # fp32 arm: the two runs differ by summation order only.  bf16x3: the hi/lo split of an activation depends on the last bit of
    # the BN statistics, so the two runs differ by 1e-6 in the forward and by the ReLU-flip floor in the gradients (see
    # test_gpu_models.py::test_mid_size_gradients_match_reference_per_tensor); a wrong world-size factor would show as >= 0.5
    assert checked > 100 and worst[0] <= (1e-4 if prec == "fp32" else 5e-2), worst
    # running statistics: global mean / unbiased global variance on every rank
    sd_ref = ref_m.state_dict()
    for r in range(world):
        sd = reps[r].state_dict()
        w = max(C.rel_err(sd[k].cpu(), sd_ref[k].cpu()) for k in sd if k.endswith(("running_mean", "running_var")))
        assert w <= 1e-4, (r, w)
    del params


def test_two_nccl_ranks_equal_the_single_device_global_batch(E, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "dist_syncbn.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "dist_syncbn_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0
    import json
    res = json.load(open(out))
    assert res.pop("overlap_rel") < 1e-5  # second step: chunks all-reduced during the backward pass, same averaged gradients
    for mode, v in res.items():
        print(mode, v)
        # fp32 arm: summation order only; bf16x3: the ReLU-flip floor (see test_two_emulated_ranks_*); a world-size factor is >= 0.5
        gate = 1e-4 if mode.endswith("fp32") else 5e-2
        assert v["loss_rel"] <= 2e-5 and v["worst_grad_rel_l2"] <= gate and v["worst_running"] <= 1e-4, (mode, v)
        assert v["bn_affine_worst"] <= gate, (mode, v)
