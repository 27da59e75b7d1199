"""Standalone execution of a single nn module through a cached launch plan (see nn/modules/_emit.py)."""
from __future__ import annotations

import torch

from . import _plan

#: plans kept per standalone module (LRU over input shapes)
MAX_MODULE_PLANS = 8


def _as_f32_cuda(t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or t.dim() != 4:
        raise TypeError("expected a 4-D NCHW tensor")
    if not t.is_cuda:
        raise RuntimeError("yololite modules run on CUDA (sm_100) only; got a CPU tensor (there is no CPU fallback)")
    return t


def run_standalone(module, x):
    """x: NCHW tensor or list of them (Concat / Detect). Returns NCHW fp32 tensor(s) like the reference module."""
    multi = isinstance(x, (list, tuple))
    xs = [_as_f32_cuda(t) for t in (x if multi else [x])]
    dev = xs[0].device
    key = (tuple(tuple(t.shape) for t in xs), dev.index)
    plans = module.__dict__.setdefault("_yl_plans", {})
    entry = plans.get(key)
    if entry is not None:
        plans[key] = plans.pop(key)
    else:
        if len(plans) >= MAX_MODULE_PLANS:          # LRU: see BaseModel._get_plan
            torch.cuda.synchronize(dev)
            plans.pop(next(iter(plans)))
        with torch.cuda.device(dev):
            g = _plan.Builder(dev)
            statics = [torch.empty(tuple(t.shape), dtype=torch.float32, device=dev) for t in xs]
            views = [g.input_nchw(s) for s in statics]
            if not getattr(module, "_takes_nchw", False):
                views = g.mat(views)
            res = module._emit(g, views if multi else views[0])
            post = getattr(module, "_yl_export", None)
            outs = post(g, res) if post is not None else g.to_nchw(res)
            entry = (g.finish(), statics, outs)
        plans[key] = entry
    plan, statics, outs = entry
    with torch.cuda.device(dev):
        for s, t in zip(statics, xs):
            s.copy_(t)
        plan.run()
    return _clone(outs)


def _clone(o):
    if isinstance(o, torch.Tensor):
        return o.clone()
    if isinstance(o, (list, tuple)):
        return type(o)(_clone(v) for v in o)
    return o
