"""GPU vs the CPU oracle, live, on configurations the golden fixtures do not cover: other clip lengths (T = 1, 2, 5),
one clip per device in eval mode, odd and non-multiple-of-8 frame sizes (the 479-pixel training crop, 853-wide frames),
another class count (`--lesslabel`: 42 classes), psp_weight with T = 4.  Same seeded weights on both sides (conditioned as
in the fixtures), bf16x3 parity mode, forward quantities within the north-star's 1e-3."""
import pytest
import torch

import cases as C
import tcb_oracle as O
from cvpr2021_vspw_implement_b200 import models as M

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cvpr2021_vspw_implement_b200 import engine
    return engine


def _build(kind, k, seed, **kw):
    torch.manual_seed(seed)
    crit = torch.nn.NLLLoss(ignore_index=255)
    enc = M.ModelBuilder.build_encoder("resnet50dilated")
    cls = M.Clip_PSP if kind == "psp" else M.ClipOCRNet
    m = cls(enc, crit, C.ns(num_class=k, **kw), deep_sup_scale=0.4)
    sd = m.state_dict()
    O.condition_weights(sd)
    m.load_state_dict(sd)
    return C.no_dropout(m)


SWEEP = [
    # kind, T, n, H, W, classes, extra args
    ("psp", 1, 2, 41, 57, 124, {}),                       # a "clip" of the current frame only
    ("psp", 5, 2, 33, 49, 124, {}),                       # the benchmark's clip length
    ("psp", 4, 2, 47, 53, 124, {"psp_weight": True}),     # the scripts' CLIPNUM=4 with the learned frame weights
    ("psp", 2, 3, 59, 43, 42, {}),                        # --lesslabel (42 classes), three clips, portrait, odd sizes
    ("ocr", 1, 2, 41, 57, 124, {}),
    ("ocr", 5, 2, 33, 49, 124, {}),
    ("ocr", 2, 3, 59, 43, 42, {}),
]


@pytest.mark.parametrize("spec", SWEEP, ids=lambda s: f"{s[0]}-T{s[1]}-n{s[2]}-{s[3]}x{s[4]}-k{s[5]}" + ("-pspw" if s[6] else ""))
def test_train_and_eval_match_the_oracle(E, spec):
    kind, T, n, H, W, k, extra = spec
    seed = 100 + T * 7 + n
    m = _build(kind, k, seed, **extra)
    sd = {key: v.clone() for key, v in m.state_dict().items()}
    imgs, labs = O.synthetic_clip(T, n, H, W, k, seed=seed + 1, block=8)
    fr, lb = C.oracle_order(imgs, labs)
    fwd = O.clip_psp_forward if kind == "psp" else O.clip_ocr_forward
    okw = {"args_psp_weight": True} if extra.get("psp_weight") else {}
    # ---- train step (batch statistics) ----
    ref = fwd({key: v.clone() for key, v in sd.items()}, fr, lb, train=True, **okw)
    g = m.cuda().train()
    C.no_dropout(g)
    with E.precision("bf16x3"), E.capturing() as cap:
        loss, acc = g(C.feed(imgs, labs, True, "cuda"))
        loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref["loss"].item()) <= TOL * abs(ref["loss"].item())
    assert abs(acc.item() - ref["acc"].item()) <= TOL
    assert C.rel_err(cap["logits"].permute(0, 3, 1, 2).cpu(), ref["logits"].detach()) <= TOL
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in g.parameters())
    # ---- inference with the ORIGINAL running statistics (the train step above updated the module's) ----
    g.load_state_dict(sd)
    g.eval()
    with torch.no_grad(), E.precision("bf16x3"):
        probs = g(C.feed(imgs, labs, False, "cuda"), segSize=(H, W))
        refp = fwd({key: v.clone() for key, v in sd.items()}, fr, None, train=False, seg_size=(H, W), **okw)["probs"]
    assert tuple(probs.shape) == (n, k, H, W)
    assert C.rel_err(probs.cpu(), refp) <= TOL
    assert (probs.argmax(1).cpu() == refp.argmax(1)).float().mean() >= 0.999


def test_eval_with_one_clip_per_device(E):
    """n = 1 is legal at inference (test_clip2.py default batch sizes) although Clip_PSP cannot TRAIN with it (quirk Q12)."""
    m = _build("psp", 124, 5)
    sd = {key: v.clone() for key, v in m.state_dict().items()}
    imgs, labs = O.synthetic_clip(3, 1, 40, 56, 124, seed=6, block=8)
    fr, _ = C.oracle_order(imgs, labs)
    g = m.cuda().eval()
    with torch.no_grad():
        probs = g(C.feed(imgs, labs, False, "cuda"), segSize=(40, 56))
        refp = O.clip_psp_forward(sd, fr, None, train=False, seg_size=(40, 56))["probs"]
    assert C.rel_err(probs.cpu(), refp) <= TOL
