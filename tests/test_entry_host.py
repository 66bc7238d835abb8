"""CPU: host logic of the entry points — yacs-compatible config, Evaluator / VC metric against the reference's
utils.py, flag surface against the reference's argparse, and the N>1 path (clip sharding + gradient bucket
all-reduce) on gloo with world_size 2."""
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _entry(name):
    """Import this repo's entry point by path (another test puts the reference tree, which has files of the same
    name, in front of sys.path)."""
    import importlib.util
    if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(ROOT):
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def test_cfg_defaults_merge_and_overrides(tmp_path):
    from cvpr2021_vspw_implement_b200.config import get_defaults
    c = get_defaults()
    assert c.TRAIN.seed == 304 and c.TRAIN.deep_sup_scale == 0.4 and c.MODEL.fc_dim == 2048
    c.merge_from_file(os.path.join(ROOT, "config", "vsp-resnet101dilated-ppm_deepsup_clip.yaml"))
    assert c.MODEL.arch_encoder == "resnet101dilated" and c.MODEL.arch_decoder == "ppm_deepsup_clip"
    assert c.DATASET.imgSizes == (300, 375, 450, 525, 600)          # "(300, ...)" string literal-evaluated like yacs
    assert c.TRAIN.weight_decay == pytest.approx(1e-4)              # YAML gives the string '1e-4'
    c.merge_from_list(["TRAIN.fix_bn", "True", "MODEL.fc_dim", "512", "DIR", "ckpt/x"])
    assert c.TRAIN.fix_bn is True and c.MODEL.fc_dim == 512 and c.DIR == "ckpt/x"
    with pytest.raises(AssertionError):
        c.merge_from_list(["TRAIN.nope", "1"])
    with pytest.raises(ValueError):
        c.merge_from_list(["MODEL.fc_dim", "abc"])
    bad = tmp_path / "bad.yaml"
    bad.write_text("MODEL:\n  unknown_key: 1\n")
    with pytest.raises(KeyError):
        c.merge_from_file(str(bad))
    assert "arch_encoder: resnet101dilated" in str(c)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_reference_yamls_load_through_the_shim():
    from cvpr2021_vspw_implement_b200.config import get_defaults
    for name in ("vsp-resnet101dilated-ppm_deepsup_clip.yaml", "vsp-resnet18dilated-ppm_deepsup.yaml"):
        c = get_defaults()
        c.merge_from_file(os.path.join(REF, "config", name))
        mine = get_defaults()
        mine.merge_from_file(os.path.join(ROOT, "config", name))
        assert c.MODEL == mine.MODEL


def _load_ref_utils():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_evaluator_and_vc_match_reference():
    from cvpr2021_vspw_implement_b200 import utils as U
    R = _load_ref_utils()
    rng = np.random.RandomState(0)
    k = 124
    a, b = U.Evaluator(k), R.Evaluator(k)
    for _ in range(3):
        gt = rng.randint(0, k, size=(2, 40, 50))
        gt[rng.rand(*gt.shape) < 0.05] = 255
        pr = np.where(rng.rand(*gt.shape) < 0.6, np.minimum(gt, k - 1), rng.randint(0, k, size=gt.shape))
        a.add_batch(gt, pr)
        b.add_batch(gt, pr)
    assert np.array_equal(a.confusion_matrix, b.confusion_matrix)
    for m in ("Pixel_Accuracy", "Pixel_Accuracy_Class", "Mean_Intersection_over_Union", "Frequency_Weighted_Intersection_over_Union"):
        assert getattr(a, m)() == pytest.approx(getattr(b, m)(), rel=1e-12), m
    gts = [rng.randint(0, 3, size=(20, 30)) for _ in range(12)]
    prs = [np.where(rng.rand(20, 30) < 0.8, g, 0) for g in gts]
    assert U.get_common(gts, prs, 4, 20, 30) == pytest.approx(R.get_common(gts, prs, 4, 20, 30))
    c = U.Evaluator(k)
    c.add_confusion(a.confusion_matrix)
    assert c.Mean_Intersection_over_Union() == pytest.approx(a.Mean_Intersection_over_Union())


def test_parse_devices():
    from cvpr2021_vspw_implement_b200.utils import parse_devices
    assert parse_devices("0-3") == ["gpu0", "gpu1", "gpu2", "gpu3"]
    assert parse_devices("0,2,gpu2") == ["gpu0", "gpu2"]
    with pytest.raises(NotImplementedError):
        parse_devices("tpu0")


def _ref_flags(path):
    src = open(path).read()
    return set(re.findall(r"add_argument\(\s*\"(--[a-z_0-9]+)\"", src))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_cli_flags_cover_the_reference():
    train_clip2 = _entry("train_clip2")
    test_clip2 = _entry("test_clip2")
    for mod, ref in ((train_clip2, "train_clip2.py"), (test_clip2, "test_clip2.py")):
        mine = {s for a in mod.make_parser()._actions for s in a.option_strings}
        missing = _ref_flags(os.path.join(REF, ref)) - mine
        assert not missing, (ref, missing)
    a = train_clip2.make_parser().parse_args(["--method", "clip_psp", "--clip_num", "4", "--dilation2", "3,6,9", "--psp_weight", "False"])
    assert a.method == "clip_psp" and a.clip_num == 4 and a.psp_weight is False and a.lr == 0.02 and a.cropsize == 531
    with pytest.raises(SystemExit):
        train_clip2.make_parser().parse_args(["--method", "not_a_method"])


def test_other_methods_raise_not_implemented():
    import argparse
    train_clip2 = _entry("train_clip2")
    from cvpr2021_vspw_implement_b200.config import get_defaults
    c = get_defaults()
    c.MODEL.arch_encoder = "resnet18dilated"
    with pytest.raises(NotImplementedError):
        train_clip2.build_module(c, argparse.Namespace(method="netwarp", num_class=124, psp_weight=False, use_memory=False, memory_num=8, clipocr_all=False))


def test_poly_lr_schedule_and_groups():
    import argparse
    train_clip2 = _entry("train_clip2")
    from cvpr2021_vspw_implement_b200.config import get_defaults
    import cases as C
    c = get_defaults()
    kind, arch, T, n, H, W, mseed, dseed = C.CASES["clip_psp"]
    m = C.build(kind, arch, mseed)
    args = argparse.Namespace(fix=False, lr=0.002, keep_duplicate_params=False)
    opt = train_clip2.create_optimizers(m, c, args)
    assert [g["weight_decay"] for g in opt.param_groups] == [1e-4, 1e-4, 0, 0]
    assert sum(len(g["params"]) for g in opt.param_groups) == len(list(m.parameters()))
    train_clip2.adjust_learning_rate(opt, 50, c, 100, args)
    lr = 0.002 * 0.5 ** 0.9
    assert [g["lr"] for g in opt.param_groups] == pytest.approx([lr * 0.1, lr, lr * 0.1, lr])


def test_synthetic_datasets_follow_the_loader_contract():
    import argparse
    from cvpr2021_vspw_implement_b200.data import SyntheticClipTest, SyntheticClipTrain
    a = argparse.Namespace(clip_num=4, num_class=124, cropsize=64, dilation2="3,6,9")
    tr = SyntheticClipTrain(a, length=4, height=48, width=64)
    loader = torch.utils.data.DataLoader(tr, batch_size=2, drop_last=True)
    imgs, gts = next(iter(loader))
    assert len(imgs) == 4 and tuple(imgs[0].shape) == (2, 3, 48, 64) and tuple(gts[0].shape) == (2, 1, 48, 64)
    vals = torch.unique(gts[0])
    assert gts[0].dtype == torch.float32 and ((vals < 124) | (vals == 255)).all()
    te = SyntheticClipTest(a, "v", frames=5, height=48, width=64)
    img, gt, clip, _, names = next(iter(torch.utils.data.DataLoader(te, batch_size=2)))
    assert tuple(img.shape) == (2, 3, 48, 64) and len(clip) == 3 and names[0].endswith(".png")


# ---- N > 1 path on gloo ---------------------------------------------------------------------------
def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from cvpr2021_vspw_implement_b200 import parallel as P
    w, r, _ = P.init_from_env(backend="gloo")
    assert (w, r) == (world, rank) and P.is_parallel()
    lo, hi = P.shard_clips(8, world, rank)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    if rank == 1:  # ranks start from different weights; broadcast must fix that
        for p in net.parameters():
            p.data.add_(1.0)
    P.broadcast_parameters(net)
    x = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    loss = net(x[lo:hi]).pow(2).mean()
    loss.backward()
    if rank == 1:
        net[1].bias.grad = None  # a parameter untouched on one rank contributes zero
    bucket = P.GradBucket(net.parameters())
    bucket.all_reduce_mean()
    m = P.mean_scalar(loss.detach())
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in net.parameters()], "loss": m, "range": (lo, hi)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_all_reduce_matches_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    x = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    l0, l1 = net(x[:4]).pow(2).mean(), net(x[4:]).pow(2).mean()
    g0 = torch.autograd.grad(l0, list(net.parameters()))
    g1 = torch.autograd.grad(l1, list(net.parameters()))
    assert got["range"] == (0, 4)
    assert float(got["loss"]) == pytest.approx(float((l0 + l1) / 2), rel=1e-6)
    for i, (a, b, g) in enumerate(zip(g0, g1, got["grads"])):
        want = (a + b) / 2 if i != 3 else a / 2   # net[1].bias grad was dropped on rank 1
        assert torch.allclose(g, want, rtol=1e-5, atol=1e-7), i


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from cvpr2021_vspw_implement_b200 import parallel as P
    P.init_from_env(backend="gloo")
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((4, 3), (5,), (2, 6), (7,), (3, 3), (2,))]
    bucket = P.GradBucket(params, overlap=True, chunk_elems=[10, 12])
    assert [(a, b) for a, b, _, _ in bucket._chunks] == [(0, 1), (1, 3), (3, 5), (5, 6)]
    unused = 3  # a parameter that never receives a gradient: the optimizer must see grad = None for it
    record, launches = [], []
    real_launch = bucket._launch
    bucket._launch = lambda lo, hi: (launches.append((len(record), lo, hi)), real_launch(lo, hi))[1]
    for step in range(3):
        bucket.zero_grad()
        record.clear()
        del launches[:]
        g = torch.Generator().manual_seed(100 * step + rank)
        for i in reversed(range(len(params))):       # the backward pass: last parameter first, one tape node per parameter
            if i == unused:
                continue
            dst = bucket.destination(params[i])
            assert dst is not None and dst.data_ptr() == params[i].grad.data_ptr()
            dst.copy_(torch.randn(params[i].shape, generator=g))
            record.append(i)
            bucket.node_done()
        during = len(launches)
        bucket.all_reduce_mean()
        # step 0 learns the pattern (everything reduced at the end: one launch); afterwards every chunk goes out as soon as
        # its last expected parameter has been written (after 1, 2, 4 and 5 parameters), only the flags at the end
        assert during == (0 if step == 0 else 4) and len(launches) == (1 if step == 0 else 5), (step, launches)
        if step:
            assert [n for n, _, _ in launches[:4]] == [1, 2, 4, 5]
        assert params[unused].grad is None
        if rank == 0:
            torch.save([None if p.grad is None else p.grad.clone() for p in params], out + f".{step}")
    with pytest.raises(RuntimeError, match="changed since the previous step"):
        bucket.zero_grad()
        for i in reversed(range(len(params))):
            bucket.destination(params[i])   # parameter 3 now receives a gradient after its chunk was scheduled
            bucket.node_done()
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_chunked_all_reduce_matches_the_mean(tmp_path):
    """GradBucket(overlap=True) on gloo, world 2: chunks are reduced during the emulated backward pass from the second step on,
    the result is the plain mean of the two ranks' gradients, unused parameters stay None."""
    port = 29500 + (os.getpid() + 7) % 2000
    out = str(tmp_path / "g")
    mp.spawn(_overlap_worker, args=(2, port, out), nprocs=2, join=True)
    shapes = ((4, 3), (5,), (2, 6), (7,), (3, 3), (2,))
    for step in range(3):
        got = torch.load(out + f".{step}")
        gens = [torch.Generator().manual_seed(100 * step + r) for r in range(2)]
        for i in reversed(range(len(shapes))):
            if i == 3:
                assert got[i] is None
                continue
            want = sum(torch.randn(shapes[i], generator=g) for g in gens) / 2
            assert torch.allclose(got[i], want, rtol=1e-6, atol=1e-7), (step, i)


def test_shard_clips_errors():
    from cvpr2021_vspw_implement_b200 import parallel as P
    assert P.shard_clips(16, 8, 3) == (6, 8)
    with pytest.raises(ValueError):
        P.shard_clips(10, 4, 0)


def test_reference_optimizer_checkpoint_with_duplicate_params_is_remapped():
    """a reference opt_epoch_E.pth repeats every parameter per group (quirk Q10); resume must map its
    momentum buffers onto the de-duplicated groups instead of failing on the group-size mismatch."""
    import train_clip2 as T
    torch.manual_seed(0)
    a, b, c = (torch.nn.Parameter(torch.randn(3)) for _ in range(3))
    opt = torch.optim.SGD([{"params": [a, b], "lr": 0.1}, {"params": [c], "lr": 1.0}], lr=0.1, momentum=0.9)
    # what torch writes for the reference's groups [a, b, a, b] and [c, c, c]: positions counted with repeats
    bufs = {0: torch.full((3,), 1.0), 1: torch.full((3,), 2.0), 4: torch.full((3,), 3.0)}
    ref = {"state": {k: {"momentum_buffer": v} for k, v in bufs.items()},
           "param_groups": [dict(opt.state_dict()["param_groups"][0], params=[0, 1, 0, 1]),
                            dict(opt.state_dict()["param_groups"][1], params=[4, 4, 4])]}
    opt.load_state_dict(T.remap_reference_optimizer_state(ref, opt))
    assert torch.equal(opt.state[a]["momentum_buffer"], bufs[0]) and torch.equal(opt.state[b]["momentum_buffer"], bufs[1])
    assert torch.equal(opt.state[c]["momentum_buffer"], bufs[4])
    own = opt.state_dict()
    assert T.remap_reference_optimizer_state(own, opt) is own  # our own checkpoints pass through untouched
    bad = {"state": {}, "param_groups": [dict(own["param_groups"][0], params=[0, 1, 2, 0]), own["param_groups"][1]]}
    with pytest.raises(ValueError, match="distinct parameters"):
        T.remap_reference_optimizer_state(bad, opt)
