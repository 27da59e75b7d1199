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
