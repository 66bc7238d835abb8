"""GPU: every C-ABI kernel family against the plain fp32 CPU op it replaces (through the ctypes boundary)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases as C
import tcb_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cvpr2021_vspw_implement_b200 import engine
    return engine


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


CONV_GEOMS = [
    # n, cin, h, w, cout, k, stride, pad, dil, bias
    (2, 3, 33, 45, 64, 3, 2, 1, 1, False),      # stem conv1 (scalar path)
    (4, 3, 95, 129, 64, 3, 2, 1, 1, False),     # stem conv1, > 4096 output pixels: the small-NW weight-gradient kernel
    (2, 64, 17, 23, 64, 3, 1, 1, 1, False),
    (2, 128, 13, 19, 128, 3, 2, 1, 1, False),   # layer2 strided 3x3
    (2, 256, 13, 19, 512, 1, 2, 0, 1, False),   # strided 1x1 downsample
    (3, 256, 9, 13, 256, 3, 1, 2, 2, False),    # dilated d2
    (2, 512, 9, 13, 512, 3, 1, 4, 4, False),    # dilated d4
    (2, 1024, 7, 9, 256, 1, 1, 0, 1, False),
    (2, 512, 7, 9, 124, 1, 1, 0, 1, True),      # classifier with bias, Cout not multiple of 128
    (2, 2048, 6, 6, 1, 1, 1, 0, 1, False),      # pspweight conv, Cout=1
    (2, 96, 5, 7, 40, 3, 1, 1, 1, True),        # odd channel counts
    (2, 2048, 1, 1, 512, 1, 1, 0, 1, False),    # PPM scale-1 map (skinny-M kernel, M = 2)
    (2, 2048, 3, 3, 512, 1, 1, 0, 1, False),    # PPM scale-3 map (skinny-M kernel, M = 18: partial row group)
    (3, 20, 3, 3, 7, 1, 1, 0, 1, True),         # skinny-M kernel with ragged Cout and bias; dgrad K = 7 stays on the tiled kernel
    (1, 64, 11, 12, 36, 1, 1, 0, 1, True),      # M = 132: just above the skinny-M limit
]


@pytest.mark.parametrize("geom", CONV_GEOMS)
def test_conv2d_fwd_dgrad_wgrad(E, geom):
    n, cin, h, w, cout, k, stride, pad, dil, has_bias = geom
    g = torch.Generator().manual_seed(hash(geom) % 1000)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g) if has_bias else None
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if has_bias else None
    yr = F.conv2d(xr, wr, br, stride=stride, padding=pad, dilation=dil)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)

    tape = E.Tape(True)
    wp = torch.nn.Parameter(wt.cuda())
    bp = torch.nn.Parameter(b.cuda()) if has_bias else None
    xv = E.Var(nhwc(x).cuda(), needs_grad=True)
    with E.precision("fp32"):
        yv = E.conv2d(tape, xv, wp, bp, stride, pad, dil)
        assert C.rel_err(nchw(yv.data.cpu()), yr.detach()) <= 2e-5
        yv.grad = nhwc(gy).cuda()
        tape.backward()
    assert C.rel_err(nchw(xv.grad.cpu()), xr.grad) <= 2e-5
    assert C.rel_err(tape.param(wp).grad.cpu(), wr.grad) <= 5e-5
    if has_bias:
        assert C.rel_err(tape.param(bp).grad.cpu(), br.grad) <= 2e-5


@pytest.mark.parametrize("c,relu,res,drop,train", [(64, True, False, False, True), (256, True, True, False, True),
                                                   (512, True, False, True, True), (128, False, False, False, True),
                                                   (256, True, True, False, False), (2048, True, False, False, True)])
def test_batchnorm_act(E, c, relu, res, drop, train):
    g = torch.Generator().manual_seed(c + relu + 2 * res)
    n, h, w = 3, 7, 9
    y = torch.randn(n, c, h, w, generator=g) * 2 + 0.5
    r = torch.randn(n, c, h, w, generator=g) if res else None
    mask = ((torch.rand(n, c, generator=g) > 0.3).float() / 0.7) if drop else None
    bn_ref = torch.nn.BatchNorm2d(c)
    bn_ref.weight.data = torch.rand(c, generator=g) + 0.5
    bn_ref.bias.data = torch.randn(c, generator=g)
    bn_ref.running_mean.data = torch.randn(c, generator=g) * 0.1
    bn_ref.running_var.data = torch.rand(c, generator=g) + 0.5
    from cvpr2021_vspw_implement_b200.models.sync_batchnorm import BatchNorm2d
    bn = BatchNorm2d(c)
    bn.load_state_dict(bn_ref.state_dict())
    bn = bn.cuda()
    bn_ref.train(train)
    yr = y.clone().requires_grad_(True)
    rr = r.clone().requires_grad_(True) if res else None
    o = bn_ref(yr)
    if res:
        o = o + rr
    if relu:
        o = F.relu(o)
    if drop:
        o = o * mask[:, :, None, None]
    go = torch.randn(o.shape, generator=g)
    o.backward(go)

    tape = E.Tape(True)
    yv = E.Var(nhwc(y).cuda(), needs_grad=True)
    rv = E.Var(nhwc(r).cuda(), needs_grad=True) if res else None
    ov = E.batchnorm_act(tape, yv, bn, relu=relu, residual=rv, chan_scale=mask.cuda() if drop else None, training=train)
    assert C.rel_err(nchw(ov.data.cpu()), o.detach()) <= 1e-5
    ov.grad = nhwc(go).cuda()
    tape.backward()
    assert C.rel_err(nchw(yv.grad.cpu()), yr.grad) <= 5e-5
    if res:
        assert C.rel_err(nchw(rv.grad.cpu()), rr.grad) <= 1e-6
    if train:
        assert C.rel_err(tape.param(bn.weight).grad.cpu(), bn_ref.weight.grad) <= 5e-5
        assert C.rel_err(tape.param(bn.bias).grad.cpu(), bn_ref.bias.grad) <= 5e-5
        assert C.rel_err(bn.running_mean.cpu(), bn_ref.running_mean) <= 1e-5
        assert C.rel_err(bn.running_var.cpu(), bn_ref.running_var) <= 1e-5


def test_bn_golden_formula(E):
    """The reference's own BN numeric pin (tests/golden/bn_formula.npz) through the CUDA kernels."""
    gd = C.golden("bn_formula")
    from cvpr2021_vspw_implement_b200.models.sync_batchnorm import BatchNorm2d
    bn = BatchNorm2d(8)
    bn.weight.data = torch.from_numpy(gd["w"])
    bn.bias.data = torch.from_numpy(gd["b"])
    bn = bn.cuda().train()
    tape = E.Tape(False)
    out = E.batchnorm_act(tape, E.Var(nhwc(torch.from_numpy(gd["x"])).cuda()), bn, relu=False)
    assert np.allclose(nchw(out.data.cpu()).numpy(), gd["y"], atol=1e-5)
    assert np.allclose(bn.running_mean.cpu().numpy(), gd["running_mean"], atol=1e-6)
    assert np.allclose(bn.running_var.cpu().numpy(), gd["running_var"], atol=1e-6)


def test_bn_train_single_value_raises(E):
    from cvpr2021_vspw_implement_b200.models.sync_batchnorm import BatchNorm2d
    bn = BatchNorm2d(8).cuda().train()
    with pytest.raises(ValueError):
        E.batchnorm_act(E.Tape(False), E.Var(torch.zeros(1, 1, 1, 8, device="cuda")), bn)


@pytest.mark.parametrize("h,w", [(33, 45), (32, 44), (7, 5)])
def test_maxpool(E, h, w):
    g = torch.Generator().manual_seed(h * w)
    x = F.relu(torch.randn(2, 64, h, w, generator=g))
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    tape = E.Tape(True)
    xv = E.Var(nhwc(x).cuda(), needs_grad=True)
    yv = E.maxpool3x3s2(tape, xv)
    assert torch.equal(nchw(yv.data.cpu()), yr.detach())
    yv.grad = nhwc(gy).cuda()
    tape.backward()
    # ties (zeros after ReLU) may pick another element of the window; compare where the input is positive
    pos = x > 0
    assert C.rel_err(nchw(xv.grad.cpu())[pos], xr.grad[pos]) <= 1e-6


@pytest.mark.parametrize("T,n,h,w,c", [(3, 2, 7, 9, 64), (5, 2, 13, 22, 128), (1, 3, 6, 6, 32), (2, 1, 60, 107, 16)])
def test_tcb_pool(E, T, n, h, w, c):
    g = torch.Generator().manual_seed(T * 100 + h)
    feat = torch.randn(T * n, c, h, w, generator=g)
    fr = feat.clone().requires_grad_(True)
    scales = (1, 2, 3, 6)
    refs = O.tcb_pool(list(torch.split(fr, n, dim=0)), scales)
    gs = [torch.randn(r.shape, generator=g) for r in refs]
    sum((r * gg).sum() for r, gg in zip(refs, gs)).backward()
    tape = E.Tape(True)
    fv = E.Var(nhwc(feat).cuda(), needs_grad=True)
    outs = E.tcb_pool(tape, fv, T, n, scales)
    for o, r, gg in zip(outs, refs, gs):
        assert C.rel_err(nchw(o.data.cpu()), r.detach()) <= 2e-5
        o.grad = nhwc(gg).cuda()
    tape.backward()
    assert C.rel_err(nchw(fv.grad.cpu()), fr.grad) <= 2e-5


def test_ppm_concat(E):
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 13, 22
    base = torch.randn(n, 64, h, w, generator=g)
    pyr = [torch.randn(n, 32, s, s, generator=g) for s in (1, 2, 3, 6)]
    br = base.clone().requires_grad_(True)
    pr = [p.clone().requires_grad_(True) for p in pyr]
    cat = torch.cat([br] + [F.interpolate(p, (h, w), mode="bilinear", align_corners=False) for p in pr], 1)
    gc = torch.randn(cat.shape, generator=g)
    cat.backward(gc)
    tape = E.Tape(True)
    bv = E.Var(nhwc(base).cuda(), needs_grad=True)
    pv = [E.Var(nhwc(p).cuda(), needs_grad=True) for p in pyr]
    cv = E.ppm_concat(tape, bv, pv)
    assert C.rel_err(nchw(cv.data.cpu()), cat.detach()) <= 1e-5
    cv.grad = nhwc(gc).cuda()
    tape.backward()
    assert C.rel_err(nchw(bv.grad.cpu()), br.grad) <= 1e-6
    for a, b in zip(pv, pr):
        assert C.rel_err(nchw(a.grad.cpu()), b.grad) <= 2e-5


@pytest.mark.parametrize("n,h,w,H,W,k", [(2, 7, 9, 49, 65, 124), (3, 6, 11, 41, 83, 124), (1, 60, 107, 480, 854, 124), (2, 5, 5, 5, 5, 19)])
def test_loss_tail(E, n, h, w, H, W, k):
    g = torch.Generator().manual_seed(n * h)
    logits = torch.randn(n, k, h, w, generator=g) * 3
    _, labs = O.synthetic_clip(1, n, H, W, k, seed=h * w, block=8)
    lab = labs[0]
    lr = logits.clone().requires_grad_(True)
    loss, lp, lb = O.nll_up(lr, lab, 255)
    acc = O.pixel_acc(lp, lb)
    (loss * 1.7).backward()
    tape = E.Tape(True)
    lv = E.Var(nhwc(logits).cuda(), needs_grad=True)
    term = E.nll_term(tape, lv, lab.cuda(), 255, want_acc=True)
    out_loss, out_acc, gslot = E.loss_combine(tape, term, None, 0.0)
    assert abs(out_loss.item() - loss.item()) <= 2e-5 * abs(loss.item())
    assert abs(out_acc.item() - acc.item()) <= 2e-5
    gslot["g"] = torch.tensor(1.7, device="cuda")
    tape.backward()
    assert C.rel_err(nchw(lv.grad.cpu()), lr.grad) <= 5e-5
    # aux-only (thread-per-pixel kernel) must agree with the warp-per-pixel one
    term2 = E.nll_term(E.Tape(False), E.Var(nhwc(logits).cuda()), lab.cuda(), 255, want_acc=False)
    torch.cuda.synchronize()
    assert abs(float(term2.acc[0] / term2.acc[1]) - loss.item()) <= 2e-5 * abs(loss.item())


def test_loss_all_ignored_is_nan(E):
    lv = E.Var(torch.randn(1, 4, 4, 8, device="cuda"))
    lab = torch.full((1, 1, 16, 16), 255.0, device="cuda")
    term = E.nll_term(E.Tape(False), lv, lab, 255, want_acc=True)
    loss, acc, _ = E.loss_combine(E.Tape(False), term, None, 0.0)
    assert torch.isnan(loss).item() and acc.item() == 0.0


@pytest.mark.parametrize("n,h,w,H,W,k", [(2, 7, 9, 49, 65, 124), (1, 9, 14, 70, 107, 124), (2, 4, 4, 33, 31, 150)])
def test_up_softmax(E, n, h, w, H, W, k):
    g = torch.Generator().manual_seed(H)
    logits = torch.randn(n, k, h, w, generator=g) * 3
    ref = F.softmax(F.interpolate(logits, (H, W), mode="bilinear", align_corners=False), dim=1)
    probs, pred = E.up_softmax(E.Var(nhwc(logits).cuda()), H, W, want_pred=True)
    assert C.rel_err(probs.cpu(), ref) <= 1e-5
    assert (pred.cpu() == ref.argmax(1)).float().mean().item() >= 0.9999


def test_region_gather_and_attention(E):
    g = torch.Generator().manual_seed(9)
    T, n, h, w, c, K, kc = 3, 2, 7, 9, 64, 24, 32
    feats = torch.randn(T * n, c, h, w, generator=g)
    dsn = torch.randn(T * n, K, h, w, generator=g) * 2
    fr, dr = feats.clone().requires_grad_(True), dsn.clone().requires_grad_(True)
    ctx = O.region_gather(fr, dr, T)  # (n, c, K, 1)
    gc = torch.randn(ctx.shape, generator=g)
    ctx.backward(gc)
    tape = E.Tape(True)
    fv, dv = E.Var(nhwc(feats).cuda(), needs_grad=True), E.Var(nhwc(dsn).cuda(), needs_grad=True)
    cv = E.region_gather(tape, fv, dv, T, n)
    assert C.rel_err(nchw(cv.data.cpu()), ctx.detach()) <= 2e-5
    cv.grad = nhwc(gc).cuda()
    tape.backward()
    assert C.rel_err(nchw(fv.grad.cpu()), fr.grad) <= 5e-5
    assert C.rel_err(nchw(dv.grad.cpu()), dr.grad) <= 5e-5

    q = torch.randn(n, kc, h, w, generator=g)
    key = torch.randn(n, kc, K, 1, generator=g)
    val = torch.randn(n, kc, K, 1, generator=g)
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, key, val))
    sim = F.softmax(kc ** -0.5 * torch.matmul(qr.reshape(n, kc, -1).permute(0, 2, 1), kr.reshape(n, kc, -1)), dim=-1)
    out = torch.matmul(sim, vr.reshape(n, kc, -1).permute(0, 2, 1)).permute(0, 2, 1).reshape(n, kc, h, w)
    go = torch.randn(out.shape, generator=g)
    out.backward(go)
    tape = E.Tape(True)
    qv, kv, vv = (E.Var(nhwc(t).cuda(), needs_grad=True) for t in (q, key, val))
    ov = E.object_attention(tape, qv, kv, vv, kc)
    assert C.rel_err(nchw(ov.data.cpu()), out.detach()) <= 2e-5
    ov.grad = nhwc(go).cuda()
    tape.backward()
    for a, b in ((qv, qr), (kv, kr), (vv, vr)):
        assert C.rel_err(nchw(a.grad.cpu()), b.grad) <= 5e-5


def test_permute_and_layout(E):
    x = torch.randn(3, 5, 7, 11)
    out = torch.empty(3, 7, 11, 5, device="cuda")
    E.permute4d(x.cuda(), out, (3, 5, 7, 11), (0, 2, 3, 1))
    assert torch.equal(out.cpu(), x.permute(0, 2, 3, 1))
    back = E.nhwc_to_nchw(out)
    assert torch.equal(back.cpu(), x)
    out2 = torch.empty(11, 3, 7, 5, device="cuda")
    E.permute4d(x.cuda(), out2, (3, 5, 7, 11), (3, 0, 2, 1))
    assert torch.equal(out2.cpu(), x.permute(3, 0, 2, 1))
    # OHWI -> OIHW of a 3x3 weight gradient: the short-middle-axis transpose (9 taps), ragged last 256-channel slab
    for co, ci in ((5, 64), (3, 300), (2, 513)):
        w = torch.randn(co, 3, 3, ci)
        o = torch.empty(co, ci, 3, 3, device="cuda")
        E.permute4d(w.cuda(), o, (co, 3, 3, ci), (0, 3, 1, 2))
        assert torch.equal(o.cpu(), w.permute(0, 3, 1, 2))


def test_confusion_matrix_matches_evaluator(E):
    import ctypes
    from cvpr2021_vspw_implement_b200._lib import lib
    g = torch.Generator().manual_seed(1)
    pred = torch.randint(0, 124, (2, 33, 41), generator=g, dtype=torch.int32)
    _, labs = O.synthetic_clip(1, 2, 33, 41, 124, seed=2, block=8)
    conf = torch.zeros(124, 124, dtype=torch.int64, device="cuda")
    pred_d, lab_d = pred.cuda(), labs[0].cuda()
    lib.call("vspw_confusion_add", ctypes.c_void_p(pred_d.data_ptr()), ctypes.c_void_p(lab_d.data_ptr()),
             ctypes.c_void_p(conf.data_ptr()), pred.numel(), 124, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    ev = O.Evaluator(124)
    ev.add_batch(labs[0].squeeze(1).numpy(), pred.numpy())
    assert np.array_equal(conf.cpu().numpy(), ev.confusion_matrix.astype(np.int64))


@pytest.mark.parametrize("shape", [(128, 64, 1, 1), (192, 320, 1, 1), (64, 128, 3, 3), (96, 160, 3, 3), (124, 512, 1, 1)])
@pytest.mark.parametrize("x3", [True, False])
def test_conv_weight_prep_and_zero_insert(E, shape, x3):
    """vspw_conv_weight_prep: OIHW fp32 -> OHWI / IHWO bf16 hi(/lo) planes, bit-exact against torch's permute + bf16
    rounding; vspw_zero_insert2_bf16: the stride-2 gradient laid on the input grid."""
    import ctypes
    from cvpr2021_vspw_implement_b200._lib import lib
    co, ci, kh, kw = shape
    g = torch.Generator().manual_seed(co + ci)
    w = torch.randn(co, ci, kh, kw, generator=g).cuda()
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    mk = lambda *s: torch.full(s, float("nan"), device="cuda", dtype=torch.bfloat16)
    oh, ih = mk(co, kh, kw, ci), mk(ci, kh, kw, co)
    ol, il = (mk(co, kh, kw, ci), mk(ci, kh, kw, co)) if x3 else (None, None)
    lib.call("vspw_conv_weight_prep", P(w), P(oh), P(ol), P(ih), P(il), co, ci, kh, kw, st)
    ref_o = w.permute(0, 2, 3, 1).contiguous()
    ref_i = w.permute(1, 2, 3, 0).contiguous()
    for got_hi, got_lo, ref in ((oh, ol, ref_o), (ih, il, ref_i)):
        hi = ref.to(torch.bfloat16)
        assert torch.equal(got_hi, hi)
        if x3:
            assert torch.equal(got_lo, (ref - hi.float()).to(torch.bfloat16))
    # zero insertion
    n, h, wd, c = 2, 7, 9, 16
    ho, wo = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
    src = torch.randn(n, ho, wo, c, generator=g).to(torch.bfloat16).cuda()
    dst = mk(n, h, wd, c)
    lib.call("vspw_zero_insert2_bf16", P(src), P(dst), n, ho, wo, c, h, wd, st)
    ref = torch.zeros(n, h, wd, c, dtype=torch.bfloat16, device="cuda")
    ref[:, ::2, ::2] = src
    assert torch.equal(dst, ref)


@pytest.mark.parametrize("x3", [True, False])
def test_weight_prep_plan_batched_launch_tracks_parameter_updates(E, x3):
    """engine._WeightPrepPlan: the second tape's first weight request rebuilds the planes of EVERY registered weight with
    vspw_conv_weight_prep_multi (one launch per tile kind) and they are bit-identical to torch's permute + bf16 rounding
    of the CURRENT values (i.e. after an in-place optimizer update), for mixed 1x1 / 3x3 / ragged shapes and a 5-D
    Conv3d 1x1x1 weight; a parameter whose storage moved is re-registered."""
    from cvpr2021_vspw_implement_b200._lib import lib
    g = torch.Generator().manual_seed(11)
    shapes = [(128, 64, 1, 1), (64, 128, 3, 3), (192, 320, 1, 1), (96, 160, 3, 3), (124, 512, 1, 1), (256, 256, 1, 1, 1), (33, 65, 1, 1)]
    params = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    plan = E._WeightPrepPlan(torch.device("cuda", torch.cuda.current_device()), x3)

    def check(pl, p):
        w = p.data.view(p.shape[0], p.shape[1], *(p.shape[2:4] if p.dim() == 4 else (1, 1)))
        for key, perm in (("ohwi", (0, 2, 3, 1)), ("ihwo", (1, 2, 3, 0))):
            ref = w.permute(*perm).contiguous()
            hi = ref.to(torch.bfloat16)
            assert torch.equal(pl[key][0], hi)
            if x3:
                assert torch.equal(pl[key][1], (ref - hi.float()).to(torch.bfloat16))
            else:
                assert pl[key][1] is None

    t1 = E.Tape(False)
    for p in params:
        check(plan.planes(t1, t1.param(p)), p)
    assert len(plan.entries) == len(params) and plan.tables is None
    with torch.no_grad():  # an optimizer step: same storage, new values
        for p in params:
            p.add_(torch.randn(p.shape, generator=g).cuda())
    n0 = lib.launches
    t2 = E.Tape(False)
    first = plan.planes(t2, t2.param(params[0]))
    assert lib.launches - n0 == 2  # one batched launch per tile kind, nothing per weight
    assert {t[0] for t in plan.tables} == {64, 32}
    for p in params:
        pl = plan.planes(t2, t2.param(p))
        check(pl, p)
    assert lib.launches - n0 == 2 and first is plan.planes(t2, t2.param(params[0]))
    # storage moved: the entry is dropped at the next refresh and the weight registers again on use
    params[1].data = params[1].data.clone() * 2
    t3 = E.Tape(False)
    for p in params:
        check(plan.planes(t3, t3.param(p)), p)
    assert plan.tables is None and len(plan.entries) == len(params)
    t4 = E.Tape(False)
    for p in params:
        check(plan.planes(t4, t4.param(p)), p)


@pytest.mark.parametrize("frames,clip_num", [(12, 8), (9, 2), (5, 5), (3, 8), (20, 3)])
def test_vc_counts_device_matches_get_common(E, frames, clip_num):
    """utils.get_common_device (vspw_vc_counts) == utils.get_common (reference utils.py:37-53) bit for bit:
    same windows (the reference's range(len - clip_num)), same ratios, nan where no pixel keeps its label."""
    import numpy as np
    from cvpr2021_vspw_implement_b200.utils import get_common, get_common_device
    g = torch.Generator().manual_seed(frames * 10 + clip_num)
    h, w = 37, 53
    # slowly changing labels / predictions so that constant runs of every length occur
    gt = torch.randint(0, 5, (1, h, w), generator=g).float().repeat(frames, 1, 1)
    pr = torch.randint(0, 5, (1, h, w), generator=g).repeat(frames, 1, 1)
    for f in range(1, frames):
        flip = torch.rand(h, w, generator=g) < 0.08
        gt[f:][:, flip] = torch.randint(0, 5, (int(flip.sum()),), generator=g).float()
        flip = torch.rand(h, w, generator=g) < 0.1
        pr[f:][:, flip] = torch.randint(0, 5, (int(flip.sum()),), generator=g)
    gt[gt == 4] = 255.0
    if frames == 9:  # a window in which no pixel keeps its label: 0/0 -> nan on both sides
        gt[3] = gt[2] + 1.0
    want = get_common([x.numpy() for x in gt], [x.numpy() for x in pr], clip_num, h, w)
    got = get_common_device(gt.cuda(), pr.cuda(), clip_num)
    assert len(got) == len(want) == max(frames - clip_num, 0)
    for a, b in zip(got, want):
        assert (a == b) or (a != a and b != b)


def test_evaluator_device_path_matches_host_path(E):
    """utils.Evaluator.add_batch_device (vspw_confusion_add on the device) == add_batch (reference
    utils.py:86-99 on NumPy), including ignore labels 255."""
    from cvpr2021_vspw_implement_b200.utils import Evaluator
    g = torch.Generator().manual_seed(3)
    gt = torch.randint(0, 124, (3, 1, 37, 53), generator=g).float()
    gt[torch.rand(gt.shape, generator=g) < 0.1] = 255.0
    pred = torch.randint(0, 124, (3, 37, 53), generator=g)
    a, b = Evaluator(124), Evaluator(124)
    a.add_batch(gt.squeeze(1).numpy(), pred.numpy())
    for _ in range(2):
        b.add_batch_device(gt.cuda(), pred.cuda())
    b.sync_device()
    assert np.array_equal(2 * a.confusion_matrix, b.confusion_matrix)
    b.reset()
    b.add_batch_device(gt.cuda(), pred.cuda())
    b.sync_device()
    assert abs(a.Mean_Intersection_over_Union() - b.Mean_Intersection_over_Union()) < 1e-12


def test_fused_sgd_matches_torch_sgd(E):
    """optim.FusedSGD (one vspw_sgd_momentum_step launch) against torch.optim.SGD — the reference's optimizer — over several
    steps: per-group lr / weight decay, momentum buffers, odd sizes, a parameter without gradient, state_dict exchange."""
    from cvpr2021_vspw_implement_b200.optim import FusedSGD
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 3, 3, 3), (5000,), (17,), (128, 64, 1, 1), (3, 4097), (1,)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    mk = lambda ps: [{"params": ps[:3], "lr": 0.02, "weight_decay": 1e-4}, {"params": ps[3:], "lr": 0.002, "weight_decay": 0.0}]
    oa, ob = torch.optim.SGD(mk(pa), lr=0.02, momentum=0.9, weight_decay=1e-4), FusedSGD(mk(pb), lr=0.02, momentum=0.9, weight_decay=1e-4)
    for step in range(4):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 2 and step == 1:
                a.grad = b.grad = None       # a parameter the step did not touch
                continue
            gr = torch.randn(a.shape, generator=g).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()
        for o in (oa, ob):
            o.param_groups[0]["lr"] = 0.02 * (1 - step / 10) ** 0.9   # poly schedule edits the groups in place
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert C.rel_err(b.detach().cpu(), a.detach().cpu()) <= 2e-6
        assert C.rel_err(ob.state[b]["momentum_buffer"].cpu(), oa.state[a]["momentum_buffer"].cpu()) <= 2e-6
    # checkpoints travel both ways (opt_epoch_E.pth, train_clip2.py:186-188)
    oc = FusedSGD(mk(pb), lr=0.02, momentum=0.9, weight_decay=1e-4)
    oc.load_state_dict(oa.state_dict())
    od = torch.optim.SGD(mk(pa), lr=0.02, momentum=0.9, weight_decay=1e-4)
    od.load_state_dict(ob.state_dict())
    with pytest.raises(NotImplementedError):
        FusedSGD(pb, lr=0.1, momentum=0.9, nesterov=True)


@pytest.mark.parametrize("n,h,w,c0,cp,co,k", [(2, 13, 17, 128, 64, 64, 3), (3, 11, 23, 64, 128, 192, 3), (2, 60, 107, 2048, 512, 512, 3),
                                              (2, 16, 16, 128, 64, 128, 1)])
def test_ppm_conv_fused_matches_concat_conv(E, n, h, w, c0, cp, co, k):
    """conv(cat([x] + [bilinear_up(P_s)]), W) without the concat (csrc/ppm.cu, PPM_conv.forward clip_psp.py:45-56): forward,
    BN statistics of the output, and the gradients of x, every P_s and W against torch in fp64 (F.interpolate + cat + conv2d).
    (2, 60, 107, 2048, 512 -> 512) is the head of BASELINE configs[1]."""
    scales = (1, 2, 3, 6)
    pad = (k - 1) // 2
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + h)
    x = torch.randn(n, c0, h, w, generator=g, device="cuda")
    ps = [torch.randn(n, cp, s, s, generator=g, device="cuda") for s in scales]
    ctot = c0 + len(scales) * cp
    wt = torch.randn(co, ctot, k, k, generator=g, device="cuda") / (ctot * k * k) ** 0.5
    xr = x.double().requires_grad_(True)
    pr = [p.double().requires_grad_(True) for p in ps]
    wr = wt.double().requires_grad_(True)
    cat = torch.cat([xr] + [F.interpolate(p, size=(h, w), mode="bilinear", align_corners=False) for p in pr], 1)
    yr = F.conv2d(cat, wr, None, 1, pad, 1)
    gy = torch.randn(n, co, h, w, generator=g, device="cuda")
    yr.backward(gy.double())

    with E.precision("bf16x3"):
        assert E.ppm_fused_supported((n, h, w, c0), [(n, s, s, cp) for s in scales], (co, ctot, k, k), pad, 1)
        tape = E.Tape(True)
        wp = torch.nn.Parameter(wt.clone())
        xv = E.Var(nhwc(x), needs_grad=True)
        pv = [E.Var(nhwc(p), needs_grad=True) for p in ps]
        yv = E.ppm_conv_fused(tape, xv, pv, wp, pad, 1, want_stats=True)
        yv.grad = nhwc(gy)
        tape.backward()
    torch.cuda.synchronize()
    assert C.rel_err(nchw(yv.data).cpu(), yr.detach().cpu()) <= 1e-4
    yd = yv.data.double().reshape(-1, co)
    assert C.rel_err(yv.stats[0].cpu(), yd.sum(0).cpu()) <= 1e-5 and C.rel_err(yv.stats[1].cpu(), (yd * yd).sum(0).cpu()) <= 1e-5
    assert C.rel_err(nchw(xv.grad).cpu(), xr.grad.cpu()) <= 1e-4
    for a, b, s in zip(pv, pr, scales):
        assert C.rel_err(nchw(a.grad).cpu(), b.grad.cpu()) <= 1e-4, s
    dw = tape.param(wp).grad
    e_base = C.rel_err(dw[:, :c0].cpu(), wr.grad[:, :c0].cpu())
    e_pyr = C.rel_err(dw[:, c0:].cpu(), wr.grad[:, c0:].cpu())
    print(f"ppm fused {n}x{h}x{w} {c0}+4x{cp}->{co} k{k}: dW base {e_base:.2e} pyramid {e_pyr:.2e}")
    assert e_base <= 1e-4 and e_pyr <= 1e-4


@pytest.mark.parametrize("T,n,h,w,K", [(5, 2, 60, 107, 124), (3, 2, 13, 21, 124), (2, 3, 16, 24, 97)])
@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("bf16", 2e-2)])
def test_ocr_tensor_core_gather_and_fused_attention(E, T, n, h, w, K, prec, tol):
    """csrc/ocr_tc.cu at the TCB-OCR geometry (K = 124 regions, 512 feature channels, 256 key channels; 60x107 = BASELINE
    configs[2]): the region gather on the tcgen05 weight-gradient kernel and the single-kernel pixel->region attention
    (Q.K^T -> softmax -> .V with scores and probabilities on chip), against torch in fp64; backward through the recorded tape."""
    c, kc = 512, 256
    g = torch.Generator(device="cuda").manual_seed(T * 100 + h)
    feats = torch.randn(T * n, c, h, w, generator=g, device="cuda")
    dsn = torch.randn(T * n, K, h, w, generator=g, device="cuda") * 2
    fr, dr = feats.double().requires_grad_(True), dsn.double().requires_grad_(True)
    ctx = O.region_gather(fr, dr, T)  # (n, c, K, 1)
    gc = torch.randn(ctx.shape, generator=g, device="cuda")
    ctx.backward(gc.double())
    with E.precision(prec):
        tape = E.Tape(True)
        fv, dv = E.Var(nhwc(feats), needs_grad=True), E.Var(nhwc(dsn), needs_grad=True)
        E.conv_profile_begin()
        cv = E.region_gather(tape, fv, dv, T, n)
        assert E.conv_profile_end()["tc_launches"] == 1, "the gather must run on tcgen05"
        e_ctx = C.rel_err(nchw(cv.data).cpu(), ctx.detach().cpu())
        cv.grad = nhwc(gc)
        tape.backward()
    e_df, e_dd = C.rel_err(nchw(fv.grad).cpu(), fr.grad.cpu()), C.rel_err(nchw(dv.grad).cpu(), dr.grad.cpu())

    q = torch.randn(n, kc, h, w, generator=g, device="cuda")
    key = torch.randn(n, kc, K, 1, generator=g, device="cuda")
    val = torch.randn(n, kc, K, 1, generator=g, device="cuda")
    qr, kr, vr = (t.double().requires_grad_(True) for t in (q, key, val))
    sim = F.softmax(kc ** -0.5 * torch.matmul(qr.reshape(n, kc, -1).permute(0, 2, 1), kr.reshape(n, kc, -1)), dim=-1)
    out = torch.matmul(sim, vr.reshape(n, kc, -1).permute(0, 2, 1)).permute(0, 2, 1).reshape(n, kc, h, w)
    go = torch.randn(out.shape, generator=g, device="cuda")
    out.backward(go.double())
    with E.precision(prec):
        tape = E.Tape(True)
        qv, kv, vv = (E.Var(nhwc(t), needs_grad=True) for t in (q, key, val))
        E.conv_profile_begin()
        ov = E.object_attention(tape, qv, kv, vv, kc)
        assert E.conv_profile_end()["tc_launches"] == 1, "the attention must run as the fused tcgen05 kernel"
        e_out = C.rel_err(nchw(ov.data).cpu(), out.detach().cpu())
        # the operand planes the kernel wrote for the f_up conv reproduce its fp32 output
        hi, lo = ov.planes
        rec = hi.float() + (lo.float() if lo is not None else 0)
        assert C.rel_err(rec.cpu(), ov.data.cpu()) <= (2e-5 if lo is not None else 8e-3)
        ov.grad = nhwc(go)
        tape.backward()
    errs = [C.rel_err(nchw(a.grad).cpu(), b.grad.cpu()) for a, b in ((qv, qr), (kv, kr), (vv, vr))]
    print(f"{prec} T={T} n={n} {h}x{w} K={K}: gather {e_ctx:.2e} (dF {e_df:.2e}, ddsn {e_dd:.2e}); attention {e_out:.2e} (dQ {errs[0]:.2e}, dK {errs[1]:.2e}, dV {errs[2]:.2e})")
    assert e_ctx <= tol and e_out <= tol
    assert e_df <= max(tol, 5e-5) and e_dd <= max(tol, 5e-5) and max(errs) <= max(tol, 5e-5)
    # without gradients (inference) no probability tensor is produced and the result is the same
    with E.precision(prec), torch.no_grad():
        tape = E.Tape(False)
        ov2 = E.object_attention(tape, E.Var(nhwc(q)), E.Var(nhwc(key)), E.Var(nhwc(val)), kc)
    assert torch.equal(ov2.data, ov.data)
