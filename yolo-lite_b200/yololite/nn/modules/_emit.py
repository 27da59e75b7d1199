"""Shared machinery of the yololite nn modules.

Every module is a parameter container with the reference's attribute names (so state_dicts and pickled
checkpoints interchange) plus an `_emit(g, x, out=None)` method that records its kernel launches
into a `_plan.Builder`.  `forward()` on a CUDA tensor builds (once per input shape) and replays such a plan
for the module alone, converting NCHW fp32 <-> the internal NHWC bf16 at the boundary, so each class is
individually usable and testable like the reference's.  There is no CPU or eager-PyTorch execution path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lazy


class YLModule(nn.Module):
    """Base: plan-cached standalone forward + cache invalidation when parameters move or reload."""

    #: True for modules whose _emit accepts the raw NCHW fp32 input (the fused stem conv)
    _takes_nchw = False

    def _emit(self, g, x, out=None):  # pragma: no cover - abstract
        raise NotImplementedError

    # -- standalone execution ------------------------------------------------------------------------
    def forward(self, x):
        return _lazy.run_standalone(self, x)

    # -- cache hygiene ---------------------------------------------------------------------------------
    def _yl_invalidate(self):
        """Drop every cached plan / packed weight / pipeline state.  Their buffers may still be in use by launches in
        flight on side streams (asynchronous predict), and freeing them would let the caching allocator reuse the
        memory under those launches: drain the device first."""
        if torch.cuda.is_available() and torch.cuda.is_initialized() and any(
                k.startswith("_yl_") for m in self.modules() for k in m.__dict__):
            torch.cuda.synchronize()
        for m in self.modules():
            for k in [k for k in m.__dict__ if k.startswith("_yl_")]:
                del m.__dict__[k]

    def _apply(self, fn, *a, **kw):
        self._yl_invalidate()
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._yl_invalidate()
        return super().load_state_dict(*a, **kw)

    def __getstate__(self):
        st = self.__dict__.copy()
        for k in [k for k in st if k.startswith("_yl_")]:
            del st[k]
        return st


def packed(conv2d: nn.Conv2d, bn: nn.BatchNorm2d | None, owner: nn.Module, tag: str = "pc"):
    """PackedConv of (conv2d [+ bn]) cached on `owner` (BN fold per utils/torch_utils.py:182-209)."""
    from ... import _ops

    key = f"_yl_{tag}"
    pc = owner.__dict__.get(key)
    dev = conv2d.weight.device
    if pc is None or pc.w.device != dev:
        if dev.type != "cuda":
            raise RuntimeError("yololite modules run on CUDA (sm_100) only: move the module with .cuda() first")
        assert conv2d.dilation == (1, 1), "dilated convolutions are not part of the YOLO11 path"
        kh, kw = conv2d.kernel_size
        assert kh == kw and conv2d.padding == (kh // 2, kh // 2), "only 'same' square kernels (autopad) are supported"
        g = conv2d.groups
        assert g == 1 or (g == conv2d.in_channels == conv2d.out_channels), "groups must be 1 or depthwise"
        bn_t = None
        if bn is not None:
            bn_t = (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
        with torch.cuda.device(dev):
            pc = _ops.pack_conv(conv2d.weight, bn_t, conv2d.bias, device=dev)
        owner.__dict__[key] = pc
    return pc


def act_flag(act: nn.Module) -> bool:
    if isinstance(act, nn.SiLU):
        return True
    if isinstance(act, nn.Identity):
        return False
    raise NotImplementedError(f"activation {type(act).__name__} is not on the YOLO11 path (SiLU / Identity only)")


def emit_any(g, m: nn.Module, x, out=None, out_dtype=torch.bfloat16):
    """Emit a yololite module, a plain nn.Conv2d (head.py:39,48: bias, no BN/act) or an nn.Sequential of them."""
    if not getattr(m, "_takes_nchw", False) and not isinstance(m, nn.Sequential):
        x = g.mat(x)
    if isinstance(m, nn.Sequential):
        mods = list(m)
        for i, sub in enumerate(mods):
            last = i == len(mods) - 1
            x = emit_any(g, sub, x, out if last else None, out_dtype if last else torch.bfloat16)
        return x
    if isinstance(m, nn.Conv2d):
        pc = packed(m, None, m)
        return g.conv(x, pc, m.stride[0], act=False, out=out, out_dtype=out_dtype)
    if isinstance(m, nn.Identity):
        return x if out is None else g.copy(x, out)
    if isinstance(m, nn.Upsample):
        assert m.mode == "nearest" and float(m.scale_factor) == 2.0, "only nearest x2 upsampling is supported"
        return g.upsample2x(x, out)
    if hasattr(m, "_emit"):
        if out_dtype is not torch.bfloat16:
            return m._emit(g, x, out=out, out_dtype=out_dtype)
        return m._emit(g, x, out=out)
    raise NotImplementedError(f"module {type(m).__name__} has no B200 kernel path")
