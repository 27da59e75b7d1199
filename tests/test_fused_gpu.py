"""GPU parity of the fused kernels (through the C-ABI): the C3k2 tail (Bottleneck 3x3 + 3x3 + cv2 1x1) and the
fused stem (ingest + layer 0 + layer 1), each against the layer-by-layer kernels and the fp32 oracle, incl. ragged
tile edges, channel-slice destinations and the no-shortcut variant."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]

pytestmark = pytest.mark.gpu


def _randomise_bn(m, seed):
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
        elif isinstance(mod, torch.nn.Conv2d):
            mod.weight.data.copy_(torch.randn(mod.weight.shape, generator=g) * (1.5 / np.sqrt(mod.weight[0].numel())))
    return m


def _ref_conv(cv, x):
    """conv -> BN (eval) -> SiLU in fp32 on the CPU (reference Conv.forward, nn/modules/conv.py:47-49)."""
    y = F.conv2d(x, cv.conv.weight.float().cpu(), None, cv.conv.stride, cv.conv.padding)
    bn = cv.bn
    y = (y - bn.running_mean.cpu()[None, :, None, None]) / torch.sqrt(bn.running_var.cpu()[None, :, None, None] + bn.eps)
    y = y * bn.weight.cpu()[None, :, None, None] + bn.bias.cpu()[None, :, None, None]
    return F.silu(y)


@pytest.mark.parametrize("c1,c2,e,shortcut,n,h,w", [
    (32, 64, 0.25, True, 2, 24, 40),      # c = 16 (thin variant), ragged: 24 = 3 x 8, 40 = 2.5 x 16
    (64, 128, 0.25, True, 1, 17, 23),     # c = 32, odd sizes: partial tiles both ways
    (64, 64, 0.5, False, 2, 16, 16),      # c = 32, C2 = 64, no shortcut
    (32, 32, 0.5, True, 3, 8, 16),        # c = 16, C2 = 32 (one channel group), exactly one tile per image
])
def test_c3k2_tail_fused_vs_layerwise_and_fp32(c1, c2, e, shortcut, n, h, w):
    from yololite.nn.modules import C3k2

    m = _randomise_bn(C3k2(c1, c2, 1, False, e, 1, shortcut), 3).eval().cuda()
    assert m.c in (16, 32)
    m.fuse_tail_max_c = 32          # the plan only fuses c = 16 by default (it is faster there); the kernel covers both
    x = torch.rand(n, c1, h, w, generator=torch.Generator().manual_seed(5)) * 2 - 1
    os.environ["YL_C3K2_FUSE"] = "1"
    m._yl_invalidate()
    y_f = m(x.cuda()).float().cpu()
    os.environ["YL_C3K2_FUSE"] = "0"
    m._yl_invalidate()
    y_u = m(x.cuda()).float().cpu()
    os.environ["YL_C3K2_FUSE"] = "1"
    # fp32 reference of the block (block.py:231-235, 330-343)
    t = _ref_conv(m.cv1, x)
    y0, y1 = t.chunk(2, 1)
    b = m.m[0]
    y2 = _ref_conv(b.cv2, _ref_conv(b.cv1, y1))
    y2 = y1 + y2 if b.add else y2
    ref = _ref_conv(m.cv2, torch.cat([y0, y1, y2], 1))
    tol = lambda d, r: (d.abs() <= 4e-2 + 2e-2 * r.abs()).all()   # noqa: E731  (bf16 feature-map tolerance)
    assert tol(y_f - ref, ref), float((y_f - ref).abs().max())
    assert tol(y_u - ref, ref)
    # fused and layer-by-layer round h and y2 to bf16 at the same points: they differ only by accumulation order
    assert float((y_f - y_u).abs().max()) <= 3e-2, float((y_f - y_u).abs().max())


@pytest.mark.parametrize("c1,c2,e,shortcut,n,h,w", [
    (32, 64, 0.25, True, 2, 24, 40), (64, 128, 0.25, True, 1, 17, 23), (64, 64, 0.5, False, 2, 16, 16),
    (32, 32, 0.5, True, 3, 8, 16), (64, 128, 0.25, True, 2, 80, 80),
])
def test_c3k2_tail_tcgen05_version_vs_layerwise(c1, c2, e, shortcut, n, h, w):
    """The tcgen05 version of the fused tail (smem-resident intermediates, epilogue -> next MMA's A operand; opt-in, see
    profiles/r02_c3k2_tc.md) agrees with the layer-by-layer path up to accumulation order, incl. ragged tiles."""
    from yololite.nn.modules import C3k2

    m = _randomise_bn(C3k2(c1, c2, 1, False, e, 1, shortcut), 3).eval().cuda()
    m.fuse_tail_max_c, m.tail_impl = 32, "tc"
    x = torch.rand(n, c1, h, w, generator=torch.Generator().manual_seed(5)) * 2 - 1
    y_tc = m(x.cuda()).float().cpu()
    kinds = [md["kind"] for md in next(iter(m.__dict__["_yl_plans"].values()))[0].meta]
    assert "c3k2_tail" in kinds
    m.fuse_tail = False
    m._yl_invalidate()
    y_u = m(x.cuda()).float().cpu()
    assert float((y_tc - y_u).abs().max()) <= 3e-2, float((y_tc - y_u).abs().max())


def test_c3k2_tail_is_used_by_the_plan():
    from yololite import _plan
    from yololite.nn.modules import C3k2

    m = _randomise_bn(C3k2(32, 64, 1, False, 0.25), 1).eval().cuda()
    g = _plan.Builder(torch.device("cuda", 0))
    x = g.alloc(1, 16, 16, 32)
    m._emit(g, x)
    assert [md["kind"] for md in g.meta] == ["conv_tc", "c3k2_tail"]


@pytest.mark.parametrize("n,h,w", [(2, 64, 64), (1, 96, 160), (3, 32, 32), (1, 36, 52)])
def test_stem_fused_vs_layerwise_and_fp32(n, h, w):
    from yololite import _plan
    from yololite.nn.modules import Conv
    from yololite.nn.modules._emit import packed

    l0 = _randomise_bn(Conv(3, 16, 3, 2), 7).eval().cuda()
    l1 = _randomise_bn(Conv(16, 32, 3, 2), 8).eval().cuda()
    x = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(11))
    xc = x.cuda()
    dev = torch.device("cuda", 0)

    g = _plan.Builder(dev)
    y_f = g.stem_fused(g.input_nchw(xc), packed(l0.conv, l0.bn, l0), packed(l1.conv, l1.bn, l1), True, True)
    g.finish().run_eager()
    g2 = _plan.Builder(dev)
    y_u = l1._emit(g2, l0._emit(g2, g2.input_nchw(xc)))
    g2.finish().run_eager()
    torch.cuda.synchronize()
    f = y_f.torch_nhwc().float().cpu().permute(0, 3, 1, 2)
    u = y_u.torch_nhwc().float().cpu().permute(0, 3, 1, 2)
    # the product rounds the image and the layer-0 map to bf16; the reference keeps fp32
    ref = _ref_conv(l1, _ref_conv(l0, x))
    assert f.shape == ref.shape == u.shape
    assert ((f - ref).abs() <= 4e-2 + 2e-2 * ref.abs()).all(), float((f - ref).abs().max())
    assert float((f - u).abs().max()) <= 3e-2, float((f - u).abs().max())


def test_stem_fused_writes_into_a_channel_slice():
    from yololite import _plan
    from yololite.nn.modules import Conv
    from yololite.nn.modules._emit import packed

    l0 = _randomise_bn(Conv(3, 16, 3, 2), 7).eval().cuda()
    l1 = _randomise_bn(Conv(16, 32, 3, 2), 8).eval().cuda()
    xc = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)).cuda()
    dev = torch.device("cuda", 0)
    g = _plan.Builder(dev)
    big = g.alloc(1, 16, 16, 96)
    big.buf.fill_(7.5)
    g.stem_fused(g.input_nchw(xc), packed(l0.conv, l0.bn, l0), packed(l1.conv, l1.bn, l1), True, True, out=big.slice(32, 32))
    g.finish().run_eager()
    g2 = _plan.Builder(dev)
    ref = g2.stem_fused(g2.input_nchw(xc), packed(l0.conv, l0.bn, l0), packed(l1.conv, l1.bn, l1), True, True)
    g2.finish().run_eager()
    torch.cuda.synchronize()
    assert torch.equal(big.buf[..., 32:64], ref.buf)
    assert (big.buf[..., :32] == 7.5).all() and (big.buf[..., 64:] == 7.5).all()


def test_model_uses_fused_stem_and_tails():
    from bench import randomise_model_
    from yololite.nn.tasks import DetectionModel

    m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
    x = torch.rand(1, 3, 64, 64, device="cuda")
    m.infer(x)
    kinds = [md["kind"] for md in m._get_plan(x.shape, x.device)[0].meta]
    assert kinds[0] == "stem_fused" and kinds.count("c3k2_tail") == 1, kinds[:8]    # only the c = 16 block (layer 2)


# ------------------------------------------------------------------------------------------------ conv chains
def _chain_fixture(n, hw, seed):
    """A small net of plain convs that exercises every member kind of a chain: 1x1, 3x3, stride 2, residual, channel
    slices (concat by aliasing), no activation, and a result stored twice (conv resolution + 2x upsampled)."""
    from yololite import _ops

    g = torch.Generator().manual_seed(seed)

    def pc(ci, co, k):
        w = torch.randn(co, ci, k, k, generator=g) * (1.5 / (ci * k * k) ** 0.5)
        return _ops.pack_conv(w, bn=None, conv_bias=torch.randn(co, generator=g) * 0.1)

    x = torch.randn(n, hw, hw, 64, generator=g).to(torch.bfloat16).cuda()
    packs = {"a": pc(64, 128, 1), "b": pc(64, 64, 3), "c": pc(64, 64, 3), "d": pc(192, 256, 1), "e": pc(256, 128, 3),
             "f": pc(128, 96, 1), "g": pc(96, 32, 3), "h": pc(32, 64, 1)}
    return x, packs


def _emit_chain_net(g, x, packs):
    from yololite._ops import View
    from yololite._plan import DualDest

    xv = View(x, 0, 64)
    cat = g.alloc(xv.n, xv.h, xv.w, 192)                               # [a (128) | c (64)]
    a = g.conv(xv, packs["a"], 1, True, out=cat.slice(0, 128))
    b = g.conv(a.slice(64, 64), packs["b"], 1, True)
    g.conv(b, packs["c"], 1, True, out=cat.slice(128, 64), res=cat.slice(64, 64))     # residual = upper half of a
    d = g.conv(cat, packs["d"], 1, True)
    e = g.conv(d, packs["e"], 2, True)                                 # stride 2: hw -> hw / 2
    dual = DualDest(None, None)
    f = g.conv(e, packs["f"], 1, False, out=dual)                      # no activation, stored at hw/2 and at hw
    gg = g.conv(f, packs["g"], 1, True)
    h = g.conv(gg, packs["h"], 1, True, res=None)
    return [cat, d, e, f, dual.up_view, gg, h]


@pytest.mark.parametrize("n,hw", [(3, 20), (2, 40), (1, 24), (5, 10)])
def test_conv_chain_is_bit_identical_to_its_layers(n, hw):
    """yl_conv_chain (one cluster per image walks all layers) against the same layers launched one by one: every
    intermediate and final tensor bit for bit, incl. ragged tiles (24 = 128-pixel tiles do not divide the map) and maps
    smaller than one tile (10x10 -> 5x5)."""
    from yololite import _plan

    dev = torch.device("cuda", 0)
    x, packs = _chain_fixture(n, hw, 100 + hw)
    outs = {}
    for mode in ("layers", "chain"):
        g = _plan.Builder(dev)
        g.chain_enabled = mode == "chain"
        g.chain_min_batch = 1
        views = _emit_chain_net(g, x, packs)
        plan = g.finish()
        kinds = [md["kind"] for md in plan.meta]
        if mode == "chain":
            assert kinds == ["conv_chain"], kinds
            assert len(plan.meta[0]["members"]) == 8
        else:
            assert kinds == ["conv_tc"] * 8, kinds
        plan.run_eager()
        plan.run_eager()            # idempotent: a second pass over the same buffers gives the same result
        torch.cuda.synchronize()
        outs[mode] = [v.torch_nhwc().clone() for v in views]
    for i, (a, b) in enumerate(zip(outs["layers"], outs["chain"])):
        assert a.shape == b.shape
        assert torch.equal(a, b), f"tensor {i}: max diff {float((a.float() - b.float()).abs().max())}"
    assert float(outs["chain"][-1].float().abs().max()) > 0


def test_model_with_chains_equals_model_without(monkeypatch):
    """yolo11n at 640x640: the plan with conv chains (forced on at this small batch) gives the same prediction, bit for
    bit, as the layer-by-layer plan; and the chains swallow the 20x20 / 40x40 convolutions (<= 12 launches remain on
    those maps)."""
    from bench import randomise_model_
    from yololite.nn.tasks import DetectionModel

    x = torch.rand(3, 3, 640, 640, generator=torch.Generator().manual_seed(9)).cuda()
    ys, metas = {}, {}
    for mode in ("0", "1"):
        monkeypatch.setenv("YL_CHAIN", mode)
        monkeypatch.setenv("YL_CHAIN_MIN_BATCH", "1")
        torch.manual_seed(4)            # same conv weights in both builds
        m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
        y, _ = m.infer(x)
        torch.cuda.synchronize()
        ys[mode] = y.clone()
        metas[mode] = m._get_plan(x.shape, x.device)[0].meta
    assert torch.equal(ys["0"], ys["1"]), float((ys["0"] - ys["1"]).abs().max())
    chains = [md for md in metas["1"] if md["kind"] == "conv_chain"]
    assert chains and sum(len(c["members"]) for c in chains) >= 40, [len(c["members"]) for c in chains]
    assert not [md for md in metas["0"] if md["kind"] == "conv_chain"]
    small = [md for md in metas["1"] if md["kind"] == "conv_tc" and ("20x20" in md["desc"] or "40x40" in md["desc"])]
    assert len(small) <= 8, [md["desc"] for md in small]      # what remains: the Detect head's last 1x1 convs


# ------------------------------------------------------------------------------------------------ DWConv 3x3 + Conv 1x1
@pytest.mark.parametrize("c,co,n,h,w,coff", [
    (64, 80, 2, 24, 40, 0),      # 80x80-level shape class: one channel block, ragged tile rows / columns
    (80, 80, 1, 17, 23, 0),      # a 64 + 16 channel split, odd sizes: partial tiles both ways
    (128, 80, 2, 40, 40, 16),    # two channel blocks, destination is a channel slice
    (256, 80, 3, 20, 20, 0),     # four channel blocks (the 20x20 level)
    (16, 32, 1, 8, 16, 0),       # one narrow block, exactly one tile
    (96, 128, 1, 12, 20, 0),     # 64 + 32 split, 128 output channels
])
def test_dw_pw_fused_vs_layerwise_and_fp32(c, co, n, h, w, coff):
    """yl_dw_pw_conv (depthwise 3x3 inside the A-operand producer of the 1x1 tcgen05 GEMM) against the two layers launched
    one by one and against fp32 PyTorch on the CPU; bytes around a channel-slice destination stay untouched."""
    from yololite import _plan
    from yololite._ops import View
    from yololite.nn.modules import Conv, DWConv
    from yololite.nn.modules.head import Detect

    seq = torch.nn.Sequential(_randomise_bn(DWConv(c, c, 3), 21 + c), _randomise_bn(Conv(c, co, 1), 22 + co)).eval().cuda()
    x = torch.rand(n, c, h, w, generator=torch.Generator().manual_seed(13)) * 2 - 1
    xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    dev = torch.device("cuda", 0)
    outs = {}
    for mode in ("fused", "layers"):
        g = _plan.Builder(dev)
        g.dwpw_enabled = mode == "fused"
        big = g.alloc(n, h, w, co + 2 * coff)
        big.buf.fill_(7.5)
        dest = big.slice(coff, co)
        if mode == "fused":
            dw, pw = seq[0], seq[1]
            from yololite.nn.modules._emit import act_flag, packed

            y = g.dwpw(View(xb, 0, c), packed(dw.conv, dw.bn, dw), act_flag(dw.act), packed(pw.conv, pw.bn, pw),
                       act_flag(pw.act), out=dest)
            assert y is not None, "the fused kernel rejected a supported shape"
            assert [md["kind"] for md in g.meta] == ["dwpw_tc"]
        else:
            t = seq[0]._emit(g, View(xb, 0, c))
            seq[1]._emit(g, t, out=dest)
            assert [md["kind"] for md in g.meta] == ["dwconv3x3", "conv_tc"]
        plan = g.finish()
        plan.run_eager()
        plan.run_eager()
        torch.cuda.synchronize()
        outs[mode] = big.buf.clone()
        if coff:
            assert (big.buf[..., :coff] == 7.5).all() and (big.buf[..., coff + co:] == 7.5).all()
    f = outs["fused"][..., coff:coff + co].float().cpu().permute(0, 3, 1, 2)
    u = outs["layers"][..., coff:coff + co].float().cpu().permute(0, 3, 1, 2)

    def ref_conv(cv, t):
        yy = F.conv2d(t, cv.conv.weight.float().cpu(), None, cv.conv.stride, cv.conv.padding, 1, cv.conv.groups)
        bn = cv.bn
        yy = (yy - bn.running_mean.cpu()[None, :, None, None]) / torch.sqrt(bn.running_var.cpu()[None, :, None, None] + bn.eps)
        return F.silu(yy * bn.weight.cpu()[None, :, None, None] + bn.bias.cpu()[None, :, None, None])

    ref = ref_conv(seq[1], ref_conv(seq[0], xb.float().cpu().permute(0, 3, 1, 2)))
    assert ((f - ref).abs() <= 4e-2 + 2e-2 * ref.abs()).all(), float((f - ref).abs().max())
    assert float((f - u).abs().max()) <= 3e-2, float((f - u).abs().max())


def test_detect_class_branch_uses_the_fused_dw_pw_kernel():
    from bench import randomise_model_
    from yololite.nn.tasks import DetectionModel

    m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
    x = torch.rand(1, 3, 64, 64, device="cuda")
    m.infer(x)
    kinds = [md["kind"] for md in m._get_plan(x.shape, x.device)[0].meta]
    assert kinds.count("dwpw_tc") == 6, kinds           # 3 levels x 2 stages of the class branch
    assert kinds.count("dwconv3x3") == 2, kinds          # what remains: Attention.pe of C2PSA


@pytest.mark.parametrize("c,co,n,h,w,xoff,act_dw,act_pw", [
    (64, 64, 1, 5, 7, 0, True, True),       # a map smaller than one 16 x 8 tile
    (64, 80, 2, 9, 33, 16, True, True),     # input is a channel slice (offset 16 of a 96-channel buffer), one column over 2 tiles
    (32, 48, 70, 8, 16, 0, True, False),    # many images of exactly one tile, no activation on the 1x1
    (80, 80, 1, 16, 16, 8, False, True),    # no activation on the depthwise conv, sliced 64 + 16 input
])
def test_dw_pw_fused_edge_shapes(c, co, n, h, w, xoff, act_dw, act_pw):
    """Edge shapes of yl_dw_pw_conv against fp32 PyTorch: maps smaller than a tile, channel-slice inputs (TMA base not at
    the start of a pixel), large batches, either activation switched off."""
    from yololite import _ops, _plan
    from yololite._ops import View

    g0 = torch.Generator().manual_seed(1000 + c + co + h)
    wd = torch.randn(c, 1, 3, 3, generator=g0) * 0.4
    bd = torch.randn(c, generator=g0) * 0.1
    wp = torch.randn(co, c, 1, 1, generator=g0) * (1.5 / c ** 0.5)
    bp = torch.randn(co, generator=g0) * 0.1
    pdw = _ops.pack_conv(wd, bn=None, conv_bias=bd)
    ppw = _ops.pack_conv(wp, bn=None, conv_bias=bp)
    x = torch.rand(n, c, h, w, generator=g0) * 2 - 1
    buf = torch.full((n, h, w, c + 2 * xoff), 3.25, dtype=torch.bfloat16, device="cuda")
    buf[..., xoff:xoff + c] = x.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    gb = _plan.Builder(torch.device("cuda", 0))
    y = gb.dwpw(View(buf, xoff, c), pdw, act_dw, ppw, act_pw)
    assert y is not None
    gb.finish().run_eager()
    torch.cuda.synchronize()
    got = y.torch_nhwc().float().cpu().permute(0, 3, 1, 2)
    xr = buf[..., xoff:xoff + c].float().cpu().permute(0, 3, 1, 2)
    t = F.conv2d(xr, wd.to(torch.bfloat16).float(), bd, 1, 1, 1, c)
    t = (F.silu(t) if act_dw else t).to(torch.bfloat16).float()
    ref = F.conv2d(t, wp.to(torch.bfloat16).float(), bp)
    ref = F.silu(ref) if act_pw else ref
    assert got.shape == ref.shape
    assert ((got - ref).abs() <= 4e-2 + 2e-2 * ref.abs()).all(), float((got - ref).abs().max())


def test_back_to_back_class_tail_equals_the_two_launches(monkeypatch):
    """yl_dw_pw_det (DWConv + Conv + final class conv with its Detect epilogue as two back-to-back GEMMs in one launch)
    against yl_dw_pw_conv followed by the head conv: the prediction of `infer` (class decode) and the detections of
    `infer_nms` (class filter) must be bit-identical, also at a confidence low enough that every anchor is a candidate."""
    from bench import randomise_model_
    from yololite.nn.tasks import DetectionModel

    x = torch.rand(3, 3, 320, 448, generator=torch.Generator().manual_seed(21)).cuda()
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("YL_DWPW_DET", mode)
        torch.manual_seed(7)
        m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
        y, _ = m.infer(x)
        torch.cuda.synchronize()
        out = [y.clone()]
        for conf in (0.001, 0.25):
            d, c = m.infer_nms(x, conf=conf, iou=0.7)
            torch.cuda.synchronize()
            out += [d.clone(), c.clone()]
        d, c = m.infer_nms(x, conf=0.25, iou=0.7)          # replay of the captured plan
        torch.cuda.synchronize()
        out += [d.clone(), c.clone()]
        res[mode] = out
        descs = [md["desc"] for md in m._get_plan(x.shape, x.device)[0].meta if md["kind"] in ("conv_tc", "dwpw_tc")]
        fused_tail = [d for d in descs if d.startswith("[dw3x3") and "+decode]" in d]
        plain_cls = [d for d in descs if "->80 k1s1" in d and d.endswith("+decode") and d.startswith("80->")]
        assert (len(fused_tail), len(plain_cls)) == ((3, 0) if mode == "1" else (0, 3)), descs[-12:]
    for a, b in zip(res["0"], res["1"]):
        assert torch.equal(a, b), float((a.float() - b.float()).abs().max())
    assert int(res["1"][2].sum()) > 0 and int(res["1"][4].sum()) > 0


def test_back_to_back_box_tail_equals_the_two_launches(monkeypatch):
    """yl_conv_b2b_det (last 3x3 conv of the Detect box branch + the final 1x1 conv + DFL / dist2bbox decode as two
    back-to-back GEMMs in one launch of the conv_tc back-to-back instantiation) against the two launches: `infer` and
    `infer_nms` must be bit-identical; covers the halo-patch (80x80, 40x40 maps) and the streamed-tap (small map) modes."""
    from bench import randomise_model_
    from yololite.nn.tasks import DetectionModel

    x = torch.rand(2, 3, 640, 640, generator=torch.Generator().manual_seed(23)).cuda()
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("YL_CONV_DET", mode)
        torch.manual_seed(7)
        m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
        y, _ = m.infer(x)
        torch.cuda.synchronize()
        out = [y.clone()]
        d, c = m.infer_nms(x, conf=0.25, iou=0.7)
        torch.cuda.synchronize()
        out += [d.clone(), c.clone()]
        y2, _ = m.infer(x)                                   # replay
        torch.cuda.synchronize()
        out += [y2.clone()]
        res[mode] = out
        descs = [md["desc"] for md in m._get_plan(x.shape, x.device)[0].meta if md["kind"] == "conv_tc"]
        fused = [d for d in descs if d.startswith("[64->64 k3s1, 64->64 k1 +decode]")]
        plain = [d for d in descs if d.startswith("64->64 k1s1") and d.endswith("+decode")]
        # (at this small batch the 20x20 level takes the small-problem N split, which the back-to-back kernel does not
        # support: that level falls back to the two launches)
        assert len(fused) + len(plain) == 3 and (len(fused) >= 2 if mode == "1" else len(fused) == 0), descs[-12:]
    for a, b in zip(res["0"], res["1"]):
        assert torch.equal(a, b), float((a.float() - b.float()).abs().max())
    assert torch.equal(res["1"][0], res["1"][3])
    assert int(res["1"][2].sum()) > 0
