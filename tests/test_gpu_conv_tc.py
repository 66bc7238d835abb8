"""GPU: the tcgen05 implicit-GEMM convolution (conv_tc.cu) against fp32 F.conv2d on the CPU and against the
exact-fp32 CUDA-core arm.  bf16x3 (parity mode) must sit far inside the 1e-3 north-star tolerance; single-pass
bf16 (fast mode) is reported and loosely gated."""
import pytest
import torch
import torch.nn.functional as F

import cases as C

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


TC_GEOMS = [
    # n, cin, h, w, cout, k, pad, dil, bias
    (2, 64, 40, 56, 64, 1, 0, 1, False),       # single k-block, Cout < BN (tile wider than the matrix)
    (2, 256, 30, 53, 256, 3, 2, 2, False),     # layer3-style dilated 3x3, ragged patch edges (53 = 3*16+5)
    (3, 512, 24, 40, 128, 1, 0, 1, True),      # 1x1 with bias
    (2, 128, 33, 47, 192, 3, 1, 1, False),     # Cout = 1.5 tiles
    (1, 512, 60, 107, 512, 3, 4, 4, False),    # layer4 geometry at the real 480p map size
    (10, 256, 12, 20, 1024, 1, 0, 1, False),   # many images, 8 n-tiles
    (2, 128, 31, 45, 128, 3, 1, 1, False, 2),  # layer2.0.conv2: 3x3 stride 2 (odd map: 31x45 -> 16x23), TMA elementStrides
    (2, 256, 30, 44, 512, 1, 0, 1, False, 2),  # layer2.0.downsample: 1x1 stride 2 (even map)
    (1, 64, 60, 107, 128, 3, 1, 1, True, 2),   # stride 2 with bias on the real 480p map width
    (2, 64, 33, 47, 64, 3, 1, 1, False),       # layer1 3x3 64->64: the 64-wide (UMMA 128x64x16, 4-stage) kernel, fwd and dgrad
    (2, 256, 24, 40, 64, 1, 0, 1, True),       # layer1 1x1 256->64 with bias: narrow forward, pair-kernel dgrad
    # multi-wave 1x1 maps on the pair kernel with short reductions
    (8, 128, 60, 80, 512, 1, 0, 1, False),     # K = 128, 2 channel tiles; dgrad is 512 -> 128
    (6, 64, 60, 107, 256, 1, 0, 1, True),      # K = 64 (one k-block per tile), bias, odd patch count
    (3, 192, 60, 107, 768, 1, 0, 1, False),    # K = 192, 3 channel tiles
    # wide few-channel 3x3 maps (wgrad: filter-row CTAs with the shared x box; fwd / dgrad: the fused hi|lo 64-wide kernel)
    (2, 64, 9, 130, 64, 3, 1, 1, True),        # 64->64 with bias, 130 = 128 + 2: second segment almost empty; wgrad: shared x box
    (1, 128, 7, 200, 128, 3, 2, 2, False),     # 128->128 dilation 2, two k-blocks per tap row
    (1, 128, 5, 100, 64, 3, 4, 4, False),      # 128->64 dilation 4 (widest box: 136 pixels); its dgrad is 64->128
    (2, 64, 6, 97, 128, 3, 1, 1, False),       # stem conv3 shape 64->128; its dgrad 128->64 walks the taps right to left
]


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("geom", TC_GEOMS)
def test_conv_tc_fwd_dgrad(E, geom, prec, tol):
    from cvpr2021_vspw_implement_b200._lib import ConvDesc, lib, PREC_BF16X3
    n, cin, h, w, cout, k, pad, dil, has_bias = geom[:9]
    stride = geom[9] if len(geom) > 9 else 1
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    assert lib.tc_supported(ConvDesc(n, h, w, cin, cout, k, k, stride, pad, dil, ho, wo, PREC_BF16X3))
    g = torch.Generator().manual_seed(abs(hash(geom)) % 1000)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g) if has_bias else None
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, b, stride=stride, padding=pad, dilation=dil)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)

    tape = E.Tape(True)
    wp = torch.nn.Parameter(wt.cuda())
    bp = torch.nn.Parameter(b.cuda()) if has_bias else None
    xv = E.Var(nhwc(x).cuda(), needs_grad=True)
    with E.precision(prec):
        E.conv_profile_begin()
        yv = E.conv2d(tape, xv, wp, bp, stride, pad, dil, want_stats=True)
        yv.grad = nhwc(gy).cuda()
        tape.backward()
        prof = E.conv_profile_end()
    assert prof["tc_launches"] >= 2, "the tcgen05 kernel must be the one that ran (fwd + dgrad)"
    # the epilogue's fused BN statistics against the fp32 output it wrote
    yd = yv.data.double().reshape(-1, cout)
    assert C.rel_err(yv.stats[0].cpu(), yd.sum(0).cpu()) <= 1e-5 and C.rel_err(yv.stats[1].cpu(), (yd * yd).sum(0).cpu()) <= 1e-5
    e_fwd = C.rel_err(nchw(yv.data.cpu()), yr.detach())
    e_dx = C.rel_err(nchw(xv.grad.cpu()), xr.grad)
    e_dw = C.rel_err(tape.param(wp).grad.cpu(), wr.grad)
    print(f"{prec} {geom}: fwd {e_fwd:.2e} dgrad {e_dx:.2e} wgrad {e_dw:.2e}")
    assert e_fwd <= tol and e_dx <= tol
    assert e_dw <= max(tol, 5e-5)


def test_conv_tc_zero_padding_is_exact(E):
    """All-ones input and weights: every output equals the number of in-bounds taps times Cin (exact in bf16),
    so any error in the TMA out-of-bounds fill or the tap offsets shows up as an integer mismatch."""
    n, c, h, w = 1, 64, 37, 61
    x = torch.ones(n, c, h, w)
    wt = torch.ones(64, c, 3, 3)
    ref = F.conv2d(x, wt, None, 1, 4, 4)
    tape = E.Tape(False)
    with E.precision("bf16x3"):
        y = E.conv2d(tape, E.Var(nhwc(x).cuda()), torch.nn.Parameter(wt.cuda()), None, 1, 4, 4)
    assert torch.equal(nchw(y.data.cpu()), ref)


@pytest.mark.parametrize("n,cin,h,w,cout", [(10, 512, 60, 107, 124), (2, 512, 60, 107, 124), (3, 256, 13, 21, 124), (2, 128, 20, 30, 100)])
def test_classifier_head_on_tensor_cores(E, n, cin, h, w, cout):
    """The 124-class 1x1 heads (clip_psp.py:40,79; clip_ocr.py:56-62): forward on the tcgen05 conv kernel with the missing classes
    zero-filled by the TMA load and clipped by the TMA store (no padded logits tensor), backward on 128-pitch operand planes of
    the logit gradient (weight gradient on the weight-gradient kernel, input gradient as a 1x1 GEMM).  Against torch in fp64."""
    g = torch.Generator(device="cuda").manual_seed(cin + cout + n)
    x = torch.randn(n, cin, h, w, generator=g, device="cuda")
    wt = torch.randn(cout, cin, 1, 1, generator=g, device="cuda") / cin ** 0.5
    b = torch.randn(cout, generator=g, device="cuda")
    xr, wr, br = x.double().requires_grad_(True), wt.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.conv2d(xr, wr, br)
    gy = torch.randn(yr.shape, generator=g, device="cuda")
    yr.backward(gy.double())
    tape = E.Tape(True)
    wp, bp = torch.nn.Parameter(wt.clone()), torch.nn.Parameter(b.clone())
    xv = E.Var(nhwc(x), needs_grad=True)
    with E.precision("bf16x3"):
        E.conv_profile_begin()
        yv = E.conv2d(tape, xv, wp, bp, 1, 0, 1)
        yv.grad = nhwc(gy)
        tape.backward()
        prof = E.conv_profile_end()
    assert prof["tc_launches"] == 3 and not prof["fp32_arm"], "fwd, dgrad and wgrad of the head must run on tcgen05"
    assert tuple(yv.data.shape) == (n, h, w, cout)
    e = [C.rel_err(nchw(yv.data).cpu(), yr.detach().cpu()), C.rel_err(nchw(xv.grad).cpu(), xr.grad.cpu()),
         C.rel_err(tape.param(wp).grad.cpu(), wr.grad.cpu()), C.rel_err(tape.param(bp).grad.cpu(), br.grad.cpu())]
    print(f"head {n}x{h}x{w} {cin}->{cout}: fwd {e[0]:.2e} dgrad {e[1]:.2e} wgrad {e[2]:.2e} dbias {e[3]:.2e}")
    assert max(e) <= 1e-4
