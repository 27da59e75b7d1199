"""Convolution modules of the YOLO11 path: Conv, DWConv, Concat (+ autopad).

API-compatible with reference yololite/nn/modules/conv.py (autopad :26-32, Conv :35-53, DWConv :100-105,
Concat :321-331): same constructor signatures, attribute names (`conv`, `bn`, `act`, `d`) and state_dict keys.
Execution differs completely: BN is folded into the weights once, and `Conv` is one launch of the NHWC
implicit-GEMM tcgen05 kernel with bias + SiLU (+ residual, + concat placement) fused in its epilogue.
The reference's other conv variants (Conv2, LightConv, Focus, GhostConv, RepConv, CBAM ...) are not used by
cfg/yolo11.yaml and are out of scope.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ._emit import YLModule, act_flag, packed

__all__ = ("autopad", "Conv", "DWConv", "Concat")


def autopad(k, p=None, d=1):
    """'same' padding for kernel k with dilation d (reference conv.py:26-32)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (v - 1) + 1 for v in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [v // 2 for v in k]
    return p


class Conv(YLModule):
    """conv -> BatchNorm -> activation as ONE fused kernel. Args: (c1, c2, k, s, p, g, d, act)."""

    default_act = nn.SiLU()
    _takes_nchw = True

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        if act is True:
            self.act = self.default_act
        elif isinstance(act, nn.Module):
            self.act = act
        else:
            self.act = nn.Identity()

    def _emit(self, g, x, out=None, res=None, upsample=False, out_dtype=torch.bfloat16):
        pc = packed(self.conv, getattr(self, "bn", None), self)
        return g.conv(x, pc, self.conv.stride[0], act=act_flag(self.act), out=out, res=res, upsample=upsample,
                      out_dtype=out_dtype)

    def forward_fuse(self, x):
        """Reference API (conv.py:51-53): identical here, BN is always folded."""
        return self.forward(x)


class DWConv(Conv):
    """Depth-wise convolution (groups = gcd(c1, c2)); 3x3 stride-1 instances run the dedicated DW kernel."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), d=d, act=act)


class Concat(YLModule):
    """Channel concatenation (reference conv.py:321-331).

    Inside a model plan this module emits nothing: producers were already told to write into slices of one
    buffer (see DetectionModel).  Standalone it copies its inputs into a fresh buffer."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def _emit(self, g, xs, out=None):
        assert self.d == 1, "only channel concatenation is on the YOLO11 path"
        # already adjacent slices of one buffer?  then the concat is a no-op view
        first = xs[0]
        adjacent = all(v.buf is first.buf for v in xs)
        off = first.coff
        for v in xs:
            adjacent = adjacent and v.coff == off
            off += v.c
        from ..._ops import View

        total = sum(v.c for v in xs)
        if adjacent and out is None:
            return View(first.buf, first.coff, total)
        dst = g._out(out, first.n, first.h, first.w, total)
        o = 0
        for v in xs:
            g.copy(v, dst.slice(o, v.c))
            o += v.c
        return dst
