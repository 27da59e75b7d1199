"""Block modules of the YOLO11 path: DFL, SPPF, C2f, C3, C3k, C3k2, Bottleneck, Attention, PSABlock, C2PSA.

Constructor signatures, attribute names and state_dict keys follow reference yololite/nn/modules/block.py
(DFL :51-69, SPPF :165-184, C2f :220-242, C3 :245-259, Bottleneck :330-343, C3k2 :720-728, C3k :731-739,
Attention :863-916, PSABlock :919-953, C2PSA :999-1038).  What runs is different: each block records launches
of the fused sm_100a kernels; `chunk`/`cat` become channel-slice aliasing, residual adds ride in conv
epilogues, the SPPF pool chain is one kernel and the attention core is one kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._emit import YLModule, emit_any, packed
from .conv import Conv

__all__ = ("DFL", "SPPF", "C2f", "C3", "C3k", "C3k2", "Bottleneck", "Attention", "PSABlock", "C2PSA")


class DFL(YLModule):
    """Distribution-focal-loss integral: softmax over c1 bins, expectation with frozen weights 0..c1-1."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        """x: (b, 4*c1, a) -> (b, 4, a).  Inside Detect this is fused into the decode kernel."""
        from ... import _ops

        if not x.is_cuda:
            raise RuntimeError("yololite modules run on CUDA (sm_100) only")
        return _ops.dfl_expectation(x, self.c1)


class Bottleneck(YLModule):
    """x + cv2(cv1(x)); the add is fused into cv2's epilogue."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def _emit(self, g, x, out=None):
        return self.cv2._emit(g, self.cv1._emit(g, x), out=out, res=x if self.add else None)


class C2f(YLModule):
    """CSP bottleneck with 2 convs; chunk/cat are slices of one (2+n)*c channel buffer."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))

    #: run [Bottleneck + cv2] as ONE kernel when the block is thin enough; YL_C3K2_FUSE=0 disables
    fuse_tail = True
    #: ... i.e. c <= this.  Measured after the round-2 conv_tc epilogue work (gpurun_out/ab2, yolo11n bs=64): the c = 16
    #: block at 160x160 is 139 us fused vs 180 us layer by layer, but the c = 32 blocks at 80x80 are 96 / 79 us fused vs
    #: 79 / 67 us layer by layer (the tcgen05 1x1 now streams at 5.7 TB/s), so only the thinnest block stays fused
    fuse_tail_max_c = 16
    #: "auto" = the mma.sync kernel (or whatever YL_C3K2_TC selects); "tc" = the tcgen05 version (A/B, parity tests)
    tail_impl = "auto"

    def _tail_fusable(self, out):
        import os

        from ... import _C, _plan
        from ._emit import act_flag

        if not self.fuse_tail or os.environ.get("YL_C3K2_FUSE", "1") == "0" or len(self.m) != 1:
            return False
        b = self.m[0]
        if type(b) is not Bottleneck or isinstance(out, _plan.DualDest):
            return False
        convs = (b.cv1, b.cv2, self.cv2)
        if any(type(cv) is not Conv or not hasattr(cv, "bn") or not act_flag(cv.act) or cv.conv.groups != 1
               or cv.conv.stride != (1, 1) for cv in convs):
            return False
        if b.cv1.conv.kernel_size != (3, 3) or b.cv2.conv.kernel_size != (3, 3) or self.cv2.conv.kernel_size != (1, 1):
            return False
        c = self.c
        if c > self.fuse_tail_max_c:
            return False
        if b.cv1.conv.in_channels != c or b.cv1.conv.out_channels != c // 2 or b.cv2.conv.out_channels != c:
            return False
        return bool(_C.load().yl_c3k2_tail_supported(c, self.cv2.conv.out_channels))

    def _emit(self, g, x, out=None):
        c, n = self.c, len(self.m)
        if self._tail_fusable(out):
            t = self.cv1._emit(g, x)                 # [y0 | y1]: the only intermediate that touches HBM
            b = self.m[0]
            return g.c3k2_tail(t, packed(b.cv1.conv, b.cv1.bn, b.cv1), packed(b.cv2.conv, b.cv2.bn, b.cv2),
                               packed(self.cv2.conv, self.cv2.bn, self.cv2), b.add, out=out, impl=self.tail_impl)
        cat = g.alloc(x.n, x.h, x.w, (2 + n) * c)
        self.cv1._emit(g, x, out=cat.slice(0, 2 * c))
        prev = cat.slice(c, c)
        for i, m in enumerate(self.m):
            prev = m._emit(g, prev, out=cat.slice((2 + i) * c, c))
        return self.cv2._emit(g, cat, out=out)

    forward_split = YLModule.forward


class C3(YLModule):
    """CSP bottleneck with 3 convs: cv3(cat(m(cv1(x)), cv2(x)))."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=((1, 1), (3, 3)), e=1.0) for _ in range(n)))

    def _merged_pack(self):
        """cv2 and cv1 read the same tensor: ONE 1x1 conv with the stacked weights [cv2; cv1] produces both, or
        None when the two convs are not alike."""
        from ... import _ops

        a, b = self.cv1, self.cv2
        ca, cb = a.conv, b.conv
        same = (ca.kernel_size == cb.kernel_size == (1, 1) and ca.stride == cb.stride == (1, 1)
                and ca.groups == cb.groups == 1 and ca.in_channels == cb.in_channels
                and ca.out_channels == cb.out_channels and ca.out_channels % 8 == 0 and ca.bias is None
                and cb.bias is None and type(a.act) is type(b.act) and hasattr(a, "bn") and hasattr(b, "bn")
                and a.bn.eps == b.bn.eps)
        if not same:
            return None
        pc = self.__dict__.get("_yl_pc12")
        dev = ca.weight.device
        if pc is None or pc.w.device != dev:
            cat = lambda f: torch.cat([f(b), f(a)])  # noqa: E731  (cv2's rows first)
            with torch.cuda.device(dev):
                pc = _ops.pack_conv(cat(lambda m: m.conv.weight),
                                    (cat(lambda m: m.bn.weight), cat(lambda m: m.bn.bias),
                                     cat(lambda m: m.bn.running_mean), cat(lambda m: m.bn.running_var), a.bn.eps),
                                    None, device=dev)
            self.__dict__["_yl_pc12"] = pc
        return pc

    def _emit(self, g, x, out=None):
        from ._emit import act_flag

        c_ = self.cv1.conv.out_channels
        blocks = list(self.m)
        pc12 = self._merged_pack() if blocks else None
        if pc12 is not None:
            # channels [0, c_) = m(cv1(x)), [c_, 2c_) = cv2(x), [2c_, 3c_) = cv1(x): the merged conv fills the
            # upper two slices in one launch, cv3 reads the lower two
            buf = g.alloc(x.n, x.h, x.w, 3 * c_)
            g.conv(g.mat(x), pc12, 1, act=act_flag(self.cv1.act), out=buf.slice(c_, 2 * c_))
            y = buf.slice(2 * c_, c_)
            for i, b in enumerate(blocks):
                y = b._emit(g, y, out=buf.slice(0, c_) if i == len(blocks) - 1 else None)
            return self.cv3._emit(g, buf.slice(0, 2 * c_), out=out)
        cat = g.alloc(x.n, x.h, x.w, 2 * c_)
        y = self.cv1._emit(g, x, out=None if blocks else cat.slice(0, c_))
        for i, b in enumerate(blocks):
            y = b._emit(g, y, out=cat.slice(0, c_) if i == len(blocks) - 1 else None)
        self.cv2._emit(g, x, out=cat.slice(c_, c_))
        return self.cv3._emit(g, cat, out=out)


class C3k(C3):
    """C3 with k x k bottlenecks (e = 1.0)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))


class C3k2(C2f):
    """C2f whose inner blocks are C3k (c3k=True) or plain Bottlenecks."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.m = nn.ModuleList(
            C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g) for _ in range(n)
        )


class SPPF(YLModule):
    """Spatial pyramid pooling (fast): cv1, three chained k x k max-pools in one kernel, cv2."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def _emit(self, g, x, out=None):
        c_ = self.cv1.conv.out_channels
        k = self.m.kernel_size if isinstance(self.m.kernel_size, int) else self.m.kernel_size[0]
        cat = g.alloc(x.n, x.h, x.w, 4 * c_)
        self.cv1._emit(g, x, out=cat.slice(0, c_))
        g.sppf_pool(cat.slice(0, c_), cat.slice(c_, c_), cat.slice(2 * c_, c_), cat.slice(3 * c_, c_), k)
        return self.cv2._emit(g, cat, out=out)


class Attention(YLModule):
    """Multi-head self-attention over the H*W positions with a depth-wise positional term.

    qkv 1x1 conv -> fused attention kernel (softmax(q^T k * scale) applied to v, never materialised in HBM)
    -> `+ pe(v)` folded into the depth-wise kernel's epilogue -> proj 1x1 conv (+ optional residual)."""

    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim**-0.5
        nh_kd = self.key_dim * num_heads
        h = dim + nh_kd * 2
        self.qkv = Conv(dim, h, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def _pe_packs(self):
        """Per-head packs of the depth-wise `pe` weights: head h owns channels [h*hd, (h+1)*hd)."""
        from ... import _ops

        packs = self.__dict__.get("_yl_pe")
        dev = self.pe.conv.weight.device
        if packs is None or packs[0].w.device != dev:
            hd, bn = self.head_dim, self.pe.bn
            packs = []
            for h in range(self.num_heads):
                s = slice(h * hd, (h + 1) * hd)
                packs.append(_ops.pack_conv(self.pe.conv.weight[s],
                                            (bn.weight[s], bn.bias[s], bn.running_mean[s], bn.running_var[s], bn.eps),
                                            None, device=dev))
            self.__dict__["_yl_pe"] = packs
        return packs

    def _emit(self, g, x, out=None, res=None):
        nh, kd, hd = self.num_heads, self.key_dim, self.head_dim
        qkv = self.qkv._emit(g, x)
        att = g.attention(qkv, nh, kd, hd, self.scale)
        mixed = g.alloc(x.n, x.h, x.w, nh * hd)
        per = 2 * kd + hd
        for h, pc in enumerate(self._pe_packs()):
            # v of head h is a contiguous channel run of qkv; pe(v) + attention output in one pass
            g.conv(qkv.slice(h * per + 2 * kd, hd), pc, 1, act=False, out=mixed.slice(h * hd, hd),
                   res=att.slice(h * hd, hd))
        return self.proj._emit(g, mixed, out=out, res=res)


class PSABlock(YLModule):
    """x + attn(x), then x + ffn(x); both adds are conv-epilogue residuals."""

    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True) -> None:
        super().__init__()
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))
        self.add = shortcut

    def _emit(self, g, x, out=None):
        x1 = self.attn._emit(g, x, res=x if self.add else None)
        h = self.ffn[0]._emit(g, x1)
        return self.ffn[1]._emit(g, h, out=out, res=x1 if self.add else None)


class C2PSA(YLModule):
    """cv1 -> split (a, b) -> PSABlocks on b -> cv2(cat(a, b)); split/cat are slices of one buffer."""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def _emit(self, g, x, out=None):
        c = self.c
        cat = g.alloc(x.n, x.h, x.w, 2 * c)
        self.cv1._emit(g, x, out=cat)
        b = cat.slice(c, c)
        blocks = list(self.m)
        for i, blk in enumerate(blocks):
            b = blk._emit(g, b, out=cat.slice(c, c) if i == len(blocks) - 1 else None)
        return self.cv2._emit(g, cat, out=out)
