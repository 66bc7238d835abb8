"""GPU: the tcgen05 conv kernels at the BENCHMARK geometries (BASELINE configs[1]: 10 frames, 60x107 stride-8 maps,
120x214 stride-4 maps) against torch in FP64 on the same device (cuDNN's fp32 Winograd kernels for the non-dilated 3x3 shapes are themselves 1e-4 off).

Why these sizes: the persistent kernels of conv_tc.cu give every CTA (pair) several tiles here — 251..1004 pair-tiles on
74 CTA pairs — so the TMEM double-buffer phase wrap, the smem ring wrapping across tiles, the weight-gradient split-K
(`red.global.add`) and the dgrad fan-in path all run the way they run in the bench step; the small geometries of
test_gpu_conv_tc.py finish in one wave.  Reference call sites: models/resnet.py:61-66 (Bottleneck convs), clip_psp.py:35-41,
74-79 (PPM_conv.conv_last_, deepsup), clip_ocr.py:43 (conv_3x3)."""
import pytest
import torch
import torch.nn.functional as F

import cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from cvpr2021_vspw_implement_b200 import engine
    return engine


GEOMS = [
    # name, n, h, w, cin, cout, k, dil, stride, fan_in
    ("l3 3x3 d2 256->256", 10, 60, 107, 256, 256, 3, 2, 1, False),
    ("l3 1x1 256->1024", 10, 60, 107, 256, 1024, 1, 1, 1, False),
    ("l3 1x1 1024->256 (+fan-in dgrad)", 10, 60, 107, 1024, 256, 1, 1, 1, True),
    ("l4 3x3 d4 512->512", 10, 60, 107, 512, 512, 3, 4, 1, False),
    ("deepsup 3x3 1024->512", 10, 60, 107, 1024, 512, 3, 1, 1, False),
    ("l4 1x1 512->2048", 10, 60, 107, 512, 2048, 1, 1, 1, False),
    ("ppm 3x3 4096->512", 2, 60, 107, 4096, 512, 3, 1, 1, False),
    ("ocr 3x3 2048->512", 10, 60, 107, 2048, 512, 3, 1, 1, False),
    ("l1 3x3 64->64", 10, 120, 214, 64, 64, 3, 1, 1, False),
    ("l1 1x1 64->256", 10, 120, 214, 64, 256, 1, 1, 1, False),
    ("l2 3x3 s2 128->128", 10, 120, 214, 128, 128, 3, 1, 2, False),
    ("l2 1x1 s2 256->512", 10, 120, 214, 256, 512, 1, 1, 2, False),
    ("stem 3x3 64->128", 10, 240, 427, 64, 128, 3, 1, 1, False),
]


@pytest.mark.parametrize("geom", GEOMS, ids=[g[0] for g in GEOMS])
@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4)])
def test_conv_tc_benchmark_geometry(E, geom, prec, tol):
    name, n, h, w, cin, cout, k, dil, stride, fan_in = geom
    pad = dil * (k - 1) // 2
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout + k)
    x = torch.randn(n, cin, h, w, generator=g, device="cuda")
    wt = torch.randn(cout, cin, k, k, generator=g, device="cuda") / (cin * k * k) ** 0.5
    xr = x.double().requires_grad_(True)
    wr = wt.double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, stride=stride, padding=pad, dilation=dil)
    gy = torch.randn(yr.shape, generator=g, device="cuda")
    yr.backward(gy.double())
    pre = torch.randn(x.shape, generator=g, device="cuda") if fan_in else None

    tape = E.Tape(True)
    wp = torch.nn.Parameter(wt.clone())
    xv = E.Var(x.permute(0, 2, 3, 1).contiguous(), needs_grad=True)
    if fan_in:
        xv.grad = pre.permute(0, 2, 3, 1).contiguous()  # another consumer already deposited its share: dx += dgrad
    with E.precision(prec):
        E.conv_profile_begin()
        yv = E.conv2d(tape, xv, wp, None, stride, pad, dil, want_stats=True)
        yv.grad = gy.permute(0, 2, 3, 1).contiguous()
        tape.backward()
        prof = E.conv_profile_end()
    assert prof["tc_launches"] == 3, "fwd, dgrad and wgrad must all run on tcgen05"
    y = yv.data.permute(0, 3, 1, 2)
    yd = yv.data.double().reshape(-1, cout)
    assert C.rel_err(yv.stats[0].cpu(), yd.sum(0).cpu()) <= 1e-5
    assert C.rel_err(yv.stats[1].cpu(), (yd * yd).sum(0).cpu()) <= 1e-5
    dx_ref = xr.grad + pre if fan_in else xr.grad
    e_fwd = C.rel_err(y.cpu(), yr.detach().cpu())
    e_dx = C.rel_err(xv.grad.permute(0, 3, 1, 2).cpu(), dx_ref.cpu())
    e_dw = C.rel_err(tape.param(wp).grad.cpu(), wr.grad.cpu())
    l2 = C.rel_l2(tape.param(wp).grad.cpu(), wr.grad.cpu())
    # signed bias of the forward: mean of (ours - exact) projected on the exact value.  tcgen05 accumulates in fp32 with
    # truncation, so a long accumulation shrinks every output by ~2.5e-8 per MMA into the same accumulator (measured: the
    # max-abs error grows LINEARLY with the accumulation length: 1e-5 at K = 2304, 1.4e-4 at K = 36 864, 2.8e-4 for a weight
    # gradient over 64 200 pixels in one CTA); the gates below scale with it.
    yd64 = yr.detach()
    bias = float(((y.double() - yd64) * yd64).sum() / (yd64 * yd64).sum())
    acc_fwd = 3 * k * k * cin / 16            # MMAs into one accumulator: forward
    acc_dx = 3 * k * k * cout / 16            # ... dgrad
    acc_dw = 3 * n * (h // stride) * (w // stride) / 16 / max(1, prof.get("wgrad_splits", 1))
    print(f"{prec} {name}: fwd {e_fwd:.2e} (signed bias {bias:+.2e}, {acc_fwd:.0f} MMAs/accumulator) dgrad {e_dx:.2e} wgrad {e_dw:.2e} (rel-L2 {l2:.2e})")
    assert e_fwd <= max(tol, 3e-8 * acc_fwd) and e_dx <= max(tol, 3e-8 * acc_dx) and e_dw <= max(tol, 3e-8 * acc_dw)
