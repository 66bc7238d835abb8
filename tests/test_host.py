"""CPU: host-side logic — C-ABI library exports, plugin-surface parity with the reference, error paths."""
import ctypes
import os
import re
import sys

import pytest
import torch

import cases as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_library_builds_loads_and_exports_every_declared_symbol():
    from cvpr2021_vspw_implement_b200 import build, _lib
    path = build.build_library()
    dll = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "vspw_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(vspw_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in include/vspw_b200.h but not exported"
    for name in _lib.EXPORTED_SYMBOLS:
        assert name in declared, f"{name} bound by ctypes but not declared in the header"
    assert _lib.lib.version() >= 100


def test_conv_desc_matches_header_layout():
    from cvpr2021_vspw_implement_b200._lib import ConvDesc
    # 13 int32 geometry fields + cin_pitch (the weight's channel pitch of a conv over a channel concat, 0 = cin)
    assert ctypes.sizeof(ConvDesc) == 14 * 4
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "vspw_b200.h")).read()
    body = hdr[hdr.index("typedef struct vspw_conv_desc {"):hdr.index("} vspw_conv_desc;")]
    fields = [f.strip() for decl in re.findall(r"int32_t ([^;]+);", body) for f in decl.split(",")]
    assert fields == [f for f, _ in ConvDesc._fields_]


def test_argument_errors_surface_as_exceptions_without_a_gpu():
    from cvpr2021_vspw_implement_b200._lib import ConvDesc, VspwError, lib
    d = ConvDesc(1, 8, 8, 16, 16, 3, 3, 1, 1, 1, 7, 7, 0)  # wrong ho/wo
    with pytest.raises(VspwError, match="do not match geometry"):
        lib.call("vspw_conv2d_fwd", ctypes.byref(d), ctypes.c_void_p(16), ctypes.c_void_p(16), None, ctypes.c_void_p(16), None)
    with pytest.raises(VspwError, match="null"):
        lib.call("vspw_fill", None, 0.0, 4, None)


def test_engine_refuses_cpu_tensors():
    from cvpr2021_vspw_implement_b200 import engine as E
    with pytest.raises(E.VspwError, match="CUDA"):
        E.input_from_frames([torch.zeros(1, 3, 8, 8)])


def test_builder_errors_match_reference():
    from cvpr2021_vspw_implement_b200.models import ModelBuilder
    with pytest.raises(Exception, match="Architecture undefined"):
        ModelBuilder.build_encoder("vgg16")
    with pytest.raises(NotImplementedError):
        ModelBuilder.build_encoder("resnet34")
    with pytest.raises(Exception, match="Architecture undefined"):
        ModelBuilder.build_decoder("nope")


def test_lr_group_generators_keep_reference_duplicates():
    kind, arch, T, n, H, W, mseed, dseed = C.CASES["clip_psp"]
    m = C.build(kind, arch, mseed)
    ys = list(m.get_1x_lr_params()) + list(m.get_10x_lr_params()) + list(m.get_1x_lr_params_bias()) + list(m.get_10x_lr_params_bias())
    uniq = {id(p) for p in ys}
    assert len(uniq) == len(list(m.parameters()))
    assert len(ys) > 2 * len(uniq)  # quirk Q10: nested named_modules x named_parameters yields duplicates


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_state_dict_and_init_identical_to_reference():
    sys.path.insert(0, REF)
    sys.path.insert(1, os.path.join(REF, "RAFT_core"))
    sys.dont_write_bytecode = True
    import importlib
    ref = importlib.import_module("models")
    from cvpr2021_vspw_implement_b200 import models as M
    crit = torch.nn.NLLLoss(ignore_index=255)
    for arch in ("resnet18dilated", "resnet50dilated", "resnet101dilated", "resnet50"):
        torch.manual_seed(0); a = M.ModelBuilder.build_encoder(arch)
        torch.manual_seed(0); b = ref.ModelBuilder.build_encoder(arch)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
        for (na, ma), (nb, mb) in zip(a.named_modules(), b.named_modules()):
            if isinstance(ma, torch.nn.Conv2d):
                assert (ma.stride, ma.padding, ma.dilation) == (mb.stride, mb.padding, mb.dilation), na
    for name in ("Clip_PSP", "ClipOCRNet"):
        torch.manual_seed(1); a = getattr(M, name)(M.ModelBuilder.build_encoder("resnet50dilated"), crit, C.ns(), deep_sup_scale=0.4)
        torch.manual_seed(1); b = getattr(ref, name)(ref.ModelBuilder.build_encoder("resnet50dilated"), crit, C.ns(), deep_sup_scale=0.4)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
        for gname in ("get_1x_lr_params", "get_10x_lr_params", "get_1x_lr_params_bias", "get_10x_lr_params_bias"):
            la, lb = list(getattr(a, gname)()), list(getattr(b, gname)())
            assert [tuple(x.shape) for x in la] == [tuple(x.shape) for x in lb]
    torch.manual_seed(2)
    a = M.SegmentationModule(M.ModelBuilder.build_encoder("resnet18dilated"), M.ModelBuilder.build_decoder("ppm_deepsup", fc_dim=512, num_class=124), crit, 0.4)
    torch.manual_seed(2)
    b = ref.SegmentationModule(ref.ModelBuilder.build_encoder("resnet18dilated"), ref.ModelBuilder.build_decoder("ppm_deepsup", fc_dim=512, num_class=124), crit, 0.4)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    # image-model decoder family: same keys, shapes and init RNG consumption as the reference builders
    for dec, fc in (("c1", 512), ("c1_deepsup", 512), ("ppm", 512), ("upernet", 2048), ("upernet_lite", 2048), ("ocrnet_deepsup", 2048)):
        torch.manual_seed(3); a = M.ModelBuilder.build_decoder(dec, fc_dim=fc, num_class=124)
        torch.manual_seed(3); b = ref.ModelBuilder.build_decoder(dec, fc_dim=fc, num_class=124)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb), dec
        assert all(torch.equal(sa[k], sb[k]) for k in sa), dec


def test_clip_window_and_sublists_follow_the_reference_rule():
    """vspw_data.clip_window / dilation_sublists (dataset2.py:143-151, 276-300) against a literal restatement of the
    reference's branches for every (length, index, clip_num); the synthetic window dataset serves the same windows."""
    import argparse
    from cvpr2021_vspw_implement_b200.data import SyntheticWindowTest
    from cvpr2021_vspw_implement_b200.vspw_data import clip_window, dilation_sublists

    def ref(length, imgindex, clip_num):
        add = int(clip_num / 2) if clip_num % 2 == 0 else int((clip_num - 1) / 2)
        addleft, addright = add, (add - 1 if clip_num % 2 == 0 else add)
        if imgindex - addleft < 0:
            start, end = 0, clip_num
            if end >= length:
                end = length
        elif imgindex + addright >= length:
            end = length
            start = max(end - clip_num, 0)
        else:
            start = imgindex - addleft
            end = start + clip_num
        return start, end

    for length in range(1, 15):
        for clip_num in range(1, 9):
            for i in range(length):
                assert clip_window(length, i, clip_num) == ref(length, i, clip_num)
    names = [f"{i:03d}" for i in range(11)]
    subs = dilation_sublists(names, 2)
    assert subs == [names[0::3], names[1::3], names[2::3]] and sorted(sum(subs, [])) == names
    args = argparse.Namespace(clip_num=4, num_class=5, dilation2="1,2,3", dilation_num=1, method="nonlocal3d")
    ds = SyntheticWindowTest(args, "v", frames=9, height=8, width=8, seed=1)
    for i in range(9):
        sub = list(range(i % 2, 9, 2))
        s, e = ref(len(sub), sub.index(i), 4)
        assert ds[i][4] == [f"{k:08d}.png" for k in sub[s:e]] and len(ds[i][2]) == e - s
    args.method = "clip_psp"
    ds = SyntheticWindowTest(args, "v", frames=9, height=8, width=8, seed=1)
    assert ds[4][4] == "00000004.png" and len(ds[4][2]) == 3


def test_environment_switches_are_documented():
    """Every VSPW_* environment variable the package, the kernels or the entry points read appears in INTEGRATION.md's table,
    and the table lists nothing that no longer exists (A/B switches come and go with the experiments: profiles/r2_ab_table.md)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = set(re.findall(r"`(VSPW_[A-Z0-9_]+)`", open(os.path.join(root, "INTEGRATION.md")).read()))
    code = set()
    files = glob.glob(os.path.join(root, "cvpr2021_vspw_implement_b200", "**", "*.py"), recursive=True)
    files += glob.glob(os.path.join(root, "cvpr2021_vspw_implement_b200", "csrc", "*.cu*"))
    files += [os.path.join(root, f) for f in ("bench.py", "train_clip2.py", "test_clip2.py")]
    for f in files:
        t = open(f).read()
        code |= set(re.findall(r'getenv\("(VSPW_[A-Z0-9_]+)"\)', t))
        code |= set(re.findall(r'environ(?:\.get)?[\(\[]"(VSPW_[A-Z0-9_]+)"', t))
    assert code, "no switches found: the patterns above no longer match the code"
    assert code - doc == set(), f"undocumented switches: {sorted(code - doc)}"
    assert doc - code == set(), f"documented switches that no longer exist: {sorted(doc - code)}"
