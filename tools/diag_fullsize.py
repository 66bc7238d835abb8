"""Diagnostic (GPU box): where does the full-size eval/train error of the CUDA path against the GPU fp32 oracle come from?
Prints logits / context / probability errors for TCB-PSP and TCB-OCR in bf16x3 and in the exact-fp32 arm."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cases as C  # noqa: E402
import tcb_oracle as O  # noqa: E402
from cvpr2021_vspw_implement_b200 import engine as E  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
T, N, K = 5, 2, 124
H, W = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "480x854").split("x"))
KINDS = {"psp": ("Clip_PSP", O.clip_psp_forward), "ocr": ("ClipOCRNet", O.clip_ocr_forward)}


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


for kind in ("psp", "ocr"):
    for gamma in (0.25,):
        m = C.no_dropout(C.build(KINDS[kind][0], "resnet101dilated", 21))
        imgs, labs = O.synthetic_clip(T, N, H, W, K, seed=304)
        sd = {k: v.detach().clone().cuda() for k, v in m.state_dict().items()}
        fr, lb = C.oracle_order([i.cuda() for i in imgs], [l.cuda() for l in labs])
        O.BN_MOMENTUM = 1.0
        with torch.no_grad():
            tr = KINDS[kind][1](sd, fr, lb, train=True)
        O.BN_MOMENTUM = 0.1
        m.load_state_dict({k: v.cpu() for k, v in sd.items()})
        with torch.no_grad():
            ref = KINDS[kind][1](sd, fr, train=False, seg_size=(H, W))
        m = m.cuda().eval()
        for prec in ("bf16x3", "fp32"):
            with torch.no_grad(), E.precision(prec), E.capturing() as cap:
                probs = m(C.feed(imgs, labs, False, "cuda"), segSize=(H, W))
            torch.cuda.synchronize()
            lg = cap["logits"].permute(0, 3, 1, 2)
            line = f"{kind} eval {prec}: logits {rel(lg, ref['logits']):.2e} (max|logit| {float(ref['logits'].abs().max()):.2f}), probs/maxprob {rel(probs, ref['probs']):.2e}, " \
                   f"argmax agree {float((probs.argmax(1) == ref['probs'].argmax(1)).float().mean()):.5f}"
            if "context" in cap:
                line += f", context {rel(cap['context'].permute(0, 3, 1, 2), ref['context']):.2e}"
            for name in ("feats", "x_dsn", "attn_ctx"):
                if name in cap and name in ref:
                    line += f", {name} {rel(cap[name].permute(0, 3, 1, 2), ref[name]):.2e}"
            print(line, flush=True)
        # top-2 margin of the oracle: how fragile is argmax on this fixture?
        top2 = ref["probs"].topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1]) / top2[:, 0]
        print(f"   oracle top-2 relative margin: median {float(margin.median()):.3e}, fraction below 1e-2: {float((margin < 1e-2).float().mean()):.4f}, below 1e-3: {float((margin < 1e-3).float().mean()):.4f}")
        del sd, ref, tr
        torch.cuda.empty_cache()
