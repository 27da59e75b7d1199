"""CPU tests of the host-side logic: model construction from yaml (parameter counts, state_dict keys identical
to the reference's), config overrides, C-ABI export list, and the loud failure when no GPU is present."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]

PARAMS = {"n": 2624080, "s": 9458752, "m": 20114688}  # reference cfg/yolo11.yaml:8-10


@pytest.mark.parametrize("scale", ["n", "s", "m"])
def test_model_matches_reference_structure(golden, scale):
    from yololite.nn.tasks import DetectionModel

    m = DetectionModel(f"yolo11{scale}.yaml", verbose=False)
    assert sum(p.numel() for p in m.parameters()) == PARAMS[scale]
    g = golden(f"model_yolo11{scale}.npz")
    ref = {str(k): tuple(int(v) for v in s[: int(nd)]) for k, s, nd in zip(g["keys"], g["shapes"], g["ndims"])}
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == ref                                   # same keys, same shapes, same order-independent set
    assert list(mine) == list(ref)                       # and same registration order
    assert m.stride.tolist() == [8.0, 16.0, 32.0]
    assert m.save == [4, 6, 10, 13, 16, 19, 22]
    assert all(bn.eps == 1e-3 for bn in m.modules() if isinstance(bn, torch.nn.BatchNorm2d))


def test_module_signatures_and_attrs():
    from yololite.nn.modules import C2PSA, C3k2, DFL, SPPF, Attention, Bottleneck, Conv, Detect, DWConv

    c = Conv(16, 32, 3, 2)
    assert c.conv.stride == (2, 2) and c.conv.padding == (1, 1) and isinstance(c.act, torch.nn.SiLU)
    assert isinstance(Conv(8, 8, 1, act=False).act, torch.nn.Identity)
    assert DWConv(64, 64, 3).conv.groups == 64
    b = Bottleneck(32, 32)
    assert b.add and b.cv1.conv.out_channels == 16
    k = C3k2(64, 128, 1, False, 0.25)
    assert k.c == 32 and k.cv2.conv.in_channels == 96
    assert type(C3k2(128, 128, 1, True).m[0]).__name__ == "C3k"
    a = Attention(128, num_heads=2)
    assert (a.head_dim, a.key_dim, a.qkv.conv.out_channels) == (64, 32, 256) and a.pe.conv.groups == 128
    assert C2PSA(256, 256).m[0].attn.num_heads == 2
    assert SPPF(256, 256).cv2.conv.in_channels == 512
    d = Detect(80, (64, 128, 256))
    assert (d.no, d.nl, d.reg_max) == (144, 3, 16) and d.cv3[0][0][0].conv.groups == 64
    assert DFL(16).conv.weight.flatten().tolist() == list(range(16))


def test_modules_refuse_cpu():
    from yololite.nn.modules import Conv

    with pytest.raises(RuntimeError, match="CUDA"):
        Conv(8, 8, 1).eval()(torch.zeros(1, 8, 4, 4))


def test_get_cfg_overrides():
    from yololite.cfg import get_cfg

    a = get_cfg(overrides={"conf": 0.1, "iou": 0.5, "max_det": 10, "classes": [0, 1], "epochs": 3})
    assert (a.conf, a.iou, a.max_det, a.classes) == (0.1, 0.5, 10, [0, 1])
    with pytest.raises(SyntaxError):
        get_cfg(overrides={"cnof": 0.1})
    with pytest.raises(ValueError):
        get_cfg(overrides={"conf": 1.5})
    with pytest.raises(TypeError):
        get_cfg(overrides={"max_det": 1.5})


def test_letterbox_matches_reference_geometry():
    from yololite.data import LetterBox

    im = np.zeros((1080, 1920, 3), np.uint8)
    out = LetterBox((640, 640), auto=True, stride=32)(image=im)
    assert out.shape == (384, 640, 3)                     # boats.jpg -> 384x640
    out = LetterBox((640, 640), auto=False)(image=im)
    assert out.shape == (640, 640, 3) and out[0, 0, 0] == 114


def test_box_helpers():
    from yololite.utils import ops

    b = torch.tensor([[10.0, 20.0, 4.0, 6.0]])
    assert ops.xywh2xyxy(b).tolist() == [[8.0, 17.0, 12.0, 23.0]]
    assert ops.xyxy2xywh(ops.xywh2xyxy(b)).tolist() == b.tolist()
    boxes = torch.tensor([[100.0, 50.0, 700.0, 400.0]])
    out = ops.scale_boxes((384, 640), boxes.clone(), (1080, 1920))
    assert out[0, 2] == 1920.0 and out[0, 0] == 300.0     # gain 1/3, pad (0, 12)


def test_c_abi_exports_every_declared_symbol():
    """include/yl11.h vs libyl11.so vs the ctypes prototypes: all three must agree (no compute calls)."""
    from yololite import _C

    header = (ROOT / "include" / "yl11.h").read_text()
    declared = set(re.findall(r"^(?:int|size_t|const char\*)\s+(yl_\w+)\(", header, re.M))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(str(_C.lib_path()))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in yl11.h but not exported"
    assert declared == set(_C.EXPORTS)
    assert _C.load().yl_version() == 200


def test_no_gpu_fails_loudly():
    from yololite import _C

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_C.YLError):
        _C.init(0)
    rc = _C.load().yl_init(0)
    assert rc == -5 and b"no CPU path" in _C.load().yl_last_error_string()


def test_product_never_imports_oracle():
    pkg = ROOT / "yolo-lite_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
    for f in pkg.rglob("*.cu"):
        assert "oracle/" not in f.read_text().replace("oracle/ ", "")


def test_plan_dependency_tracking_on_channel_slices():
    """Builder._track derives RAW / WAR / WAW edges from channel-slice accesses: writers of disjoint concat
    slices are independent, a reader depends on every overlapping writer, an overwrite waits for its readers."""
    import torch

    from yololite._ops import View
    from yololite._plan import Builder

    b = Builder.__new__(Builder)
    b._access = {}
    buf = torch.zeros(1, 2, 2, 64)
    other = torch.zeros(1, 2, 2, 16)
    assert b._track(0, (), (View(buf, 0, 32),)) == ()                  # producer of slice [0, 32)
    assert b._track(1, (), (View(buf, 32, 32),)) == ()                 # producer of slice [32, 64): independent
    assert b._track(2, (View(buf, 16, 32),), (View(other, 0, 16),)) == (0, 1)   # reads across both slices
    assert b._track(3, (View(buf, 0, 16),), ()) == (0,)                # reads only the first producer's slice
    assert b._track(4, (View(other, 0, 16),), (View(buf, 0, 64),)) == (0, 1, 2, 3)  # overwrite: WAW + WAR
    assert b._track(5, (View(buf, 40, 8),), ()) == (1, 4)              # every overlapping earlier writer


def test_conv_chain_grouping_and_dependency_remap():
    """Builder._fuse_chains (host logic only, stub library): runs of consecutive chainable convs of ONE lane become one
    call; a run is broken by any other launch and by a lane change; the chain inherits its members' outside dependencies
    and later calls' dependencies are renumbered onto it."""
    import ctypes as C
    import types

    import torch

    from yololite import _C, _plan

    built = []

    def build(arr, n, ptr, nbytes, out, stream):
        built.append(n)
        return 0

    # (a ctypes library hands out ONE function object per symbol; the plan compares them by identity)
    lib = types.SimpleNamespace(yl_conv_bn_act=lambda *a: 0, yl_other=lambda *a: 0, yl_conv_chain_supported=lambda a: 1,
                                yl_conv_chain_desc_bytes=lambda n: 64 * n, yl_conv_chain_build=build,
                                yl_conv_chain_run=lambda *a: 0)
    b = _plan.Builder.__new__(_plan.Builder)
    b.lib, b.device, b.buffers = lib, torch.device("cpu"), []
    b.chain_enabled, b.chain_min_batch, b.chain_max_hw = True, 1, 1600

    def conv_call(hw=20, n=4):
        a = _C.ConvArgs()
        a.x.n, a.x.h, a.x.w, a.y.h, a.y.w = n, hw, hw, hw, hw
        return (lib.yl_conv_bn_act, (C.byref(a),), (a, None, None)), {"kind": "conv_tc", "bytes": 10, "flops": 100, "desc": f"c{hw}"}

    other = ((lib.yl_other, (), ()), {"kind": "pool", "bytes": 1, "flops": 0, "desc": "pool"})
    seq = [conv_call(80), conv_call(), conv_call(), conv_call(), other, conv_call(), conv_call(), conv_call(), conv_call()]
    b.calls = [c for c, _ in seq]
    b.meta = [m for _, m in seq]
    b.lanes = [0, 0, 0, 0, 0, 0, 0, 1, 1]
    #            0    1     2     3       4     5     6     7       8
    b.deps = [(), (0,), (1,), (1, 2), (3,), (4,), (5,), (4, 6), (7,)]
    import yololite._C as cmod

    orig = cmod.stream_ptr
    cmod.stream_ptr = lambda *a, **k: 0
    try:
        b._fuse_chains()
    finally:
        cmod.stream_ptr = orig
    kinds = [m["kind"] for m in b.meta]
    # call 0 (80x80: too large) stays; 1-3 chain; pool; 5-6 chain (lane 0); 7-8 chain (lane 1)
    assert kinds == ["conv_tc", "conv_chain", "pool", "conv_chain", "conv_chain"], kinds
    assert built == [3, 2, 2]
    assert b.lanes == [0, 0, 0, 0, 1]
    assert b.deps == [(), (0,), (1,), (2,), (2, 3)], b.deps
    assert [len(m.get("members", [])) for m in b.meta] == [0, 3, 0, 2, 2]
    assert b.meta[1]["bytes"] == 30 and b.meta[1]["flops"] == 300
