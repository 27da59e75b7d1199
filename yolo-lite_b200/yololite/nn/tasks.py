"""Model graph of the YOLO11 detector: yaml -> module list -> static B200 launch plan.

Public surface follows reference yololite/nn/tasks.py: `DetectionModel(cfg, ch, nc, verbose)` with
`.model` (nn.Sequential, modules carry `.i .f .type .np`), `.save`, `.stride`, `.names`, `.yaml`, `.forward /
.predict`; `parse_model` (:525-664), `yaml_model_load` (:667-680), `guess_model_scale` (:683-698),
`attempt_load_one_weight / attempt_load_weights` (:461-522).  The layer loop of `_predict_once` (:118-145)
exists only at plan-build time: per input shape the whole forward is compiled into a fixed launch sequence
(optionally a CUDA graph), with skip connections and `Concat` resolved to buffer aliasing.
Training-side members (loss, criterion, augment/TTA, profiling, Ensemble) are out of scope.
"""
from __future__ import annotations

import ast
import math
import re
from copy import deepcopy
from pathlib import Path

import torch
import torch.nn as nn

from .. import _plan
from ..cfg import CFG_DIR, DEFAULT_CFG_DICT
from ..utils import LOGGER, yaml_load
from .modules import (C2PSA, C3, SPPF, Bottleneck, C2f, C3k, C3k2, Concat, Conv, Detect, DWConv)
from .modules._emit import YLModule, emit_any

_MODULES = {m.__name__: m for m in (Conv, DWConv, Bottleneck, SPPF, C2f, C3, C3k, C3k2, C2PSA, Concat, Detect)}
_REPEATABLE = {C2f, C3, C3k2, C2PSA}            # repeats become the module's own `n` argument
_CHANNEL_ARGS = {Conv, DWConv, Bottleneck, SPPF, C2f, C3, C3k2, C2PSA}


def make_divisible(x, divisor):
    """Smallest multiple of divisor >= x (reference utils/ops.py:101-114)."""
    if isinstance(divisor, torch.Tensor):
        divisor = int(divisor.max())
    return math.ceil(x / divisor) * divisor


def guess_model_scale(model_path) -> str:
    m = re.search(r"yolo[v]?\d+([nslmx])", Path(model_path).stem)
    return m.group(1) if m else ""


def yaml_model_load(path) -> dict:
    """'yolo11s.yaml' -> cfg/yolo11.yaml + scale 's' (reference tasks.py:667-680). Bare names resolve to the
    packaged cfg directory."""
    path = Path(path)
    unified = Path(re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", str(path)))
    for cand in (unified, path, CFG_DIR / unified.name, CFG_DIR / path.name):
        if cand.exists():
            d = yaml_load(cand)
            break
    else:
        raise FileNotFoundError(f"model yaml '{path}' not found (also looked in {CFG_DIR})")
    d["scale"] = guess_model_scale(path)
    d["yaml_file"] = str(path)
    return d


def parse_model(d: dict, ch: int, verbose: bool = True):
    """Build the nn.Sequential of a model dict. Returns (model, sorted save list)."""
    nc, scales = d.get("nc"), d.get("scales")
    depth, width, max_channels = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0), float("inf")
    scale = d.get("scale")
    if scales:
        if not scale:
            scale = next(iter(scales))
            LOGGER.warning(f"WARNING: no model scale passed, assuming scale='{scale}'")
        depth, width, max_channels = scales[scale]
    if d.get("activation"):
        raise NotImplementedError("custom activations are not on the YOLO11 path (SiLU only)")
    legacy = True
    chs = [ch]
    layers, save = [], []
    for i, (f, n, name, args) in enumerate(d["backbone"] + d["head"]):
        args = list(args)
        if name.startswith("nn."):
            m = getattr(nn, name[3:])
        elif name in _MODULES:
            m = _MODULES[name]
        else:
            raise NotImplementedError(f"layer {i}: module '{name}' is outside the YOLO11 detection path")
        for j, a in enumerate(args):
            if isinstance(a, str):
                if a == "nc":
                    args[j] = nc
                else:
                    try:
                        args[j] = ast.literal_eval(a)
                    except (ValueError, SyntaxError):
                        pass
        n_ = n = max(round(n * depth), 1) if n > 1 else n
        if m in _CHANNEL_ARGS:
            c1, c2 = chs[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in _REPEATABLE:
                args.insert(2, n)
                n = 1
            if m is C3k2:
                legacy = False
                if scale in "mlx":
                    args[3] = True
        elif m is Concat:
            c2 = sum(chs[x] for x in f)
        elif m is Detect:
            args.append([chs[x] for x in f])
            m.legacy = legacy
            c2 = None
        else:  # nn.Upsample, nn.Identity ...
            c2 = chs[f]
        m_ = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
        m_.np = sum(p.numel() for p in m_.parameters())
        m_.i, m_.f, m_.type = i, f, f"{m.__module__}.{m.__name__}".replace("torch.nn.modules.", "torch.nn.")
        if verbose:
            LOGGER.info(f"{i:>3}{str(f):>20}{n_:>3}{m_.np:10.0f}  {m_.type:<45}{str(args):<30}")
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            chs = []
        chs.append(c2)
    return nn.Sequential(*layers), sorted(save)


def initialize_weights(model: nn.Module):
    """Reference utils/torch_utils.py:242-252: BN eps 1e-3 / momentum 0.03, in-place activations."""
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps = 1e-3
            m.momentum = 0.03
        elif isinstance(m, (nn.SiLU, nn.ReLU, nn.LeakyReLU, nn.Hardswish, nn.ReLU6)):
            m.inplace = True


def fused_up_guard(layers):
    """Indices of layers whose output feeds an nn.Upsample (those producers need the dual-destination conv)."""
    out = set()
    for u, m in enumerate(layers):
        if isinstance(m, nn.Upsample):
            f = m.f if isinstance(m.f, int) else m.f[0]
            out.add(u - 1 if f == -1 else f)
    return out


class BaseModel(YLModule):
    """Layer-list model executed through a cached static plan (one per input shape and device)."""

    #: capture each plan into a CUDA graph (one launch per forward); set False to debug launch by launch
    use_cuda_graph = True
    #: plans kept per model (LRU): every (shape, slot, NMS arguments) owns activation buffers + a CUDA graph, so a
    #: stream of differently shaped rect batches or a conf/iou sweep must not grow GPU memory without bound
    max_plans = 16
    #: return fresh tensors from forward() like the reference; engine code uses infer() and skips the copy
    clone_outputs = True

    def forward(self, x, *args, **kwargs):
        if isinstance(x, dict):
            raise NotImplementedError("training/loss is out of scope: yololite is the inference path")
        return self.predict(x, *args, **kwargs)

    def predict(self, x, profile=False, visualize=False, augment=False, embed=None):
        if augment or visualize or embed or profile:
            raise NotImplementedError("augment / visualize / embed / profile are not part of the inference hot path")
        out = self.infer(x, want_raw=True)   # the module-level API returns the raw head maps like the reference
        if self.clone_outputs:
            y, raws = out
            return y.clone(), [r.clone() for r in raws]
        return out

    _predict_once = predict

    # ------------------------------------------------------------------ plan
    def _layer_scales(self):
        """Down-sampling factor of every layer output w.r.t. the image (replaces the 256x256 probe forward of
        the reference, tasks.py:258-268)."""
        sc = []
        for m in self.model:
            f = m.f
            src = (sc[f] if f != -1 else (sc[-1] if sc else 1.0)) if isinstance(f, int) else \
                [(sc[j] if j != -1 else sc[-1]) for j in f]
            if isinstance(m, Conv):
                s = src * m.conv.stride[0]
            elif isinstance(m, nn.Upsample):
                s = src / float(m.scale_factor)
            elif isinstance(m, Concat):
                s = src[0]
            elif isinstance(m, Detect):
                s = src
            else:
                s = src
            sc.append(s)
        return sc

    def _emit(self, g, x, out=None):
        """Record the whole forward. Concat layers whose inputs each feed exactly one Concat become aliasing:
        the producers are handed a slice of the concat buffer as their destination."""
        from .._ops import View

        layers = list(self.model)
        n_layers = len(layers)
        # absolute source indices per layer
        srcs = []
        for i, m in enumerate(layers):
            f = [m.f] if isinstance(m.f, int) else list(m.f)
            srcs.append([(i - 1 if j == -1 else (j if j >= 0 else i + j)) for j in f])
        # placement: producer layer -> (concat layer, channel offset) when it feeds exactly one Concat
        n_cat_uses = [0] * n_layers
        for i, m in enumerate(layers):
            if isinstance(m, Concat):
                for j in srcs[i]:
                    if j >= 0:
                        n_cat_uses[j] += 1
        place, cat_buf = {}, {}
        for i, m in enumerate(layers):
            if isinstance(m, Concat) and all(j >= 0 and n_cat_uses[j] == 1 for j in srcs[i]) \
                    and len(set(srcs[i])) == len(srcs[i]):
                for j in srcs[i]:
                    place[j] = i
        ys: list = [None] * n_layers
        out_ch = {}

        def dest_for(j):
            ci = place.get(j)
            if ci is None:
                return None

            def resolve(n, h, w, c, j=j, ci=ci):
                if ci not in cat_buf:
                    # channel counts of all sources are known statically from the modules
                    chans = [self._out_channels(layers, k, out_ch) for k in srcs[ci]]
                    cat_buf[ci] = (g.alloc(n, h, w, sum(chans)), chans)
                buf, chans = cat_buf[ci]
                off = sum(chans[: srcs[ci].index(j)])
                return buf.slice(off, c)

            return _plan.Dest(resolve)

        # nn.Upsample fused into its producer: when layer j ends in a dense conv, that conv's TMA store also
        # writes the 2x2-replicated copy straight into the upsample's destination (usually a Concat slice)
        n_uses = [0] * n_layers
        for i in range(n_layers):
            for j in srcs[i]:
                if j >= 0:
                    n_uses[j] += 1
        fused_up = {}                    # producer layer j -> upsample layer u
        for u, m in enumerate(layers):
            if isinstance(m, nn.Upsample) and m.mode == "nearest" and float(m.scale_factor) == 2.0:
                j = srcs[u][0]
                if j >= 0 and isinstance(layers[j], (Conv, C2f, C3, SPPF, C2PSA)) and j not in fused_up:
                    fused_up[j] = u
        done_up = {}

        # layers 0 + 1 as one launch (image ingest + two stride-2 3x3 convs) when layer 0 feeds only layer 1
        fuse_stem = self._stem_fusable(g, x, layers, srcs, n_uses)
        cur = x
        for i, m in enumerate(layers):
            if fuse_stem and i == 0:
                continue
            if fuse_stem and i == 1:
                from .modules._emit import act_flag, packed

                l0, l1 = layers[0], layers[1]
                cur = ys[1] = g.stem_fused(x, packed(l0.conv, l0.bn, l0), packed(l1.conv, l1.bn, l1), act_flag(l0.act),
                                           act_flag(l1.act), out=dest_for(1))
                continue
            inp = [ys[j] if j >= 0 else x for j in srcs[i]]
            if isinstance(m, Detect):
                return m._emit(g, inp)
            if isinstance(m, Concat):
                if i in cat_buf and all(place.get(j) == i for j in srcs[i]):
                    buf, chans = cat_buf[i]
                    cur = View(buf.buf, buf.coff, sum(chans))   # sources already live in the buffer
                else:
                    cur = m._emit(g, inp, out=dest_for(i))
            elif i in done_up:
                cur = done_up[i]
            elif i in fused_up:
                u = fused_up[i]
                dd = _plan.DualDest(dest_for(i), dest_for(u))
                cur = emit_any(g, m, inp[0], out=dd)
                assert dd.up_view is not None, f"layer {i}: producer did not honour the fused upsample"
                done_up[u] = dd.up_view
            else:
                cur = emit_any(g, m, inp[0], out=dest_for(i))
            ys[i] = cur
        return cur

    #: run layers 0 and 1 as one kernel when the shapes allow (YL_STEM_FUSE=0 disables)
    fuse_stem = True

    def _stem_fusable(self, g, x, layers, srcs, n_uses):
        import os

        from .. import _C

        if not self.fuse_stem or os.environ.get("YL_STEM_FUSE", "1") == "0" or len(layers) < 3:
            return False
        if not isinstance(x, _plan.NchwInput) or x.view is not None or g.calls:
            return False
        l0, l1 = layers[0], layers[1]
        if type(l0) is not Conv or type(l1) is not Conv or srcs[1] != [0] or n_uses[0] != 1 or 1 in fused_up_guard(layers):
            return False
        for cv in (l0, l1):
            c = cv.conv
            if c.kernel_size != (3, 3) or c.stride != (2, 2) or c.groups != 1 or not hasattr(cv, "bn") or c.bias is not None:
                return False
        if l0.conv.in_channels != x.c:
            return False
        return bool(_C.load().yl_stem_fused_supported(x.c, l0.conv.out_channels, l1.conv.out_channels))

    @staticmethod
    def _out_channels(layers, k, memo):
        if k in memo:
            return memo[k]
        m = layers[k]
        if isinstance(m, Conv):
            c = m.conv.out_channels
        elif isinstance(m, (C2f, SPPF, C2PSA, C3)):
            c = m.cv2.conv.out_channels if not isinstance(m, C3) else m.cv3.conv.out_channels
        elif isinstance(m, Concat):
            f = [m.f] if isinstance(m.f, int) else list(m.f)
            c = sum(BaseModel._out_channels(layers, (k - 1 if j == -1 else j), memo) for j in f)
        elif isinstance(m, nn.Sequential):
            c = BaseModel._out_channels(list(m), len(m) - 1, {})
        else:  # Upsample / Identity: same as its input
            j = m.f if isinstance(m.f, int) else m.f[0]
            c = BaseModel._out_channels(layers, k - 1 if j == -1 else j, memo)
        memo[k] = c
        return c

    def _get_plan(self, shape, dev, want_raw=False, slot=0, nms=None):
        """`nms`: None or the hashable argument tuple of Builder.nms: the plan then ends with the batched NMS and its
        entry carries (dets, counts) instead of the raw head maps."""
        plans = self.__dict__.setdefault("_yl_plans", {})
        key = (tuple(shape), dev.index, bool(self.use_cuda_graph), bool(want_raw), int(slot), nms)
        entry = plans.get(key)
        if entry is not None:
            plans[key] = plans.pop(key)                     # most recently used last (dicts keep insertion order)
        else:
            if len(plans) >= max(int(self.max_plans), 1):
                # evict the least recently used plan; its buffers / graph may still be in flight on some stream, and the
                # caching allocator would hand them to the new plan: drain the device first (rare, build-time cost)
                torch.cuda.synchronize(dev)
                plans.pop(next(iter(plans)))
            with torch.cuda.device(dev):
                g = _plan.Builder(dev)
                g.want_raw = bool(want_raw)   # Detect writes its raw (B, H, W, no) maps only on request
                if nms is not None and not nms[4] and nms[2] is None:
                    g.nms_fuse_conf = float(nms[0])   # single-label, no class filter: the head may filter in its epilogue
                static_in = torch.zeros(tuple(shape), dtype=torch.float32, device=dev)
                xin = g.input_nchw(static_in)
                y, raws = self._emit(g, xin)
                post = g.nms(y, *nms) if nms is not None else None
                plan = g.finish()
                if self.use_cuda_graph:
                    plan.capture(skip=1)     # the image ingest stays eager so it can read the caller's tensor
                raw_views = [r.buf.permute(0, 3, 1, 2) for r in raws]   # (B, no, H, W) views, zero-copy
                entry = (plan, static_in, y, raw_views if post is None else post)
            plans[key] = entry
        return entry

    @torch.no_grad()
    def infer(self, x: torch.Tensor, want_raw: bool = False, slot: int = 0):
        """Engine entry: NCHW float image batch on CUDA -> (y (B, 4+nc, A) fp32, [raw (B, no, H, W) views]).

        `slot` selects one of several independent plans (own activation buffers and CUDA graph) for the same
        shape, so a caller can keep two batches in flight on two streams: the launch-latency-bound small layers
        of one batch then overlap the bandwidth-bound large layers of the other.

        The raw head maps are only needed by callers of the module-level API (`forward`); the engine path
        (`want_raw=False`) gets an empty list and the plan never writes them.  The returned tensors alias the
        plan's static buffers and are overwritten by the next call with the same shape."""
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise TypeError("expected a (B, C, H, W) tensor")
        if not x.is_cuda:
            raise RuntimeError("yololite runs on CUDA (sm_100) only; move the input with .cuda() "
                               "(there is no CPU fallback)")
        dev = x.device
        p0 = next(self.parameters())
        if p0.device != dev:
            raise RuntimeError(f"model is on {p0.device} but input is on {dev}")
        s = int(self.stride.max()) if hasattr(self, "stride") else 32
        if x.shape[2] % s or x.shape[3] % s:
            raise ValueError(f"image size {tuple(x.shape[2:])} must be a multiple of the model stride {s}")
        plan, static_in, y, raws = self._get_plan(x.shape, dev, want_raw, slot)
        self._run_plan(plan, static_in, x, dev)
        return y, raws

    @staticmethod
    def _run_plan(plan, static_in, x, dev):
        """Feed `x` to the plan's ingest.  fp32 is read in place (zero-copy); fp16 is widened and uint8 is taken as
        image bytes (value / 255, the predictor's `/255`, predictor.py:84) — inside the fused stem when the plan has one,
        otherwise by one conversion launch into the plan's static fp32 input."""
        from .. import _C

        with torch.cuda.device(dev):
            yd = _C.INGEST_DTYPES.get(x.dtype)
            if yd is not None and x.is_contiguous() and yd in plan.native_ingest and x.data_ptr() % 16 == 0:
                plan.run(ingest_ptr=x.data_ptr(), ingest_dtype=yd)
            elif yd in (_C.YL_F16, _C.YL_U8) and x.is_contiguous() and x.data_ptr() % 16 == 0:
                _C.check(_C.load().yl_to_f32(x.data_ptr(), yd, static_in.data_ptr(), x.numel(), _C.stream_ptr()), "yl_to_f32")
                plan.run()
            else:
                if x.dtype == torch.uint8:
                    static_in.copy_(x.float() / 255, non_blocking=True)
                else:
                    static_in.copy_(x, non_blocking=True)
                plan.run()

    @torch.no_grad()
    def infer_nms(self, x: torch.Tensor, conf=0.25, iou=0.45, classes=None, agnostic=False, multi_label=False,
                  max_det=300, max_nms=30000, max_wh=7680.0, slot: int = 0):
        """Engine entry for a whole step: NCHW float batch on CUDA -> (dets (B, max_det, 6), counts (B,) int32), the
        padded output of `ops.nms_padded(self.infer(x)[0], ...)`, with the NMS kernels recorded in the same plan /
        CUDA graph as the model (ingest + one graph launch per step).  The returned tensors alias plan buffers and
        are overwritten by the next call with the same shape, arguments and slot."""
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or not x.is_cuda:
            raise RuntimeError("infer_nms expects a (B, C, H, W) CUDA tensor (yololite has no CPU fallback)")
        dev = x.device
        s = int(self.stride.max()) if hasattr(self, "stride") else 32
        if x.shape[2] % s or x.shape[3] % s:
            raise ValueError(f"image size {tuple(x.shape[2:])} must be a multiple of the model stride {s}")
        if classes is not None and not isinstance(classes, (list, tuple)):
            classes = [classes] if isinstance(classes, int) else list(classes)      # `classes=0` is legal in the reference
        nms = (float(conf), float(iou), None if classes is None else tuple(int(c) for c in classes), bool(agnostic),
               bool(multi_label), int(max_det), int(max_nms), float(max_wh))
        plan, static_in, _, post = self._get_plan(x.shape, dev, False, slot, nms)
        self._run_plan(plan, static_in, x, dev)
        return post

    def fuse(self, verbose=True):
        """Reference API (AutoBackend calls model.fuse(), autobackend.py:74; the reference deleted the method
        and crashes there).  BN is always folded at plan build, so this is a no-op returning self."""
        return self


class DetectionModel(BaseModel):
    """YOLO11 detection model built from a yaml dict or path."""

    def __init__(self, cfg="yolo11n.yaml", ch=3, nc=None, verbose=True):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        ch = self.yaml["ch"] = self.yaml.get("ch", ch)
        if nc and nc != self.yaml["nc"]:
            LOGGER.info(f"Overriding model.yaml nc={self.yaml['nc']} with nc={nc}")
            self.yaml["nc"] = nc
        self.model, self.save = parse_model(deepcopy(self.yaml), ch=ch, verbose=verbose)
        self.names = {i: f"{i}" for i in range(self.yaml["nc"])}
        self.inplace = self.yaml.get("inplace", True)
        self.end2end = False
        m = self.model[-1]
        if isinstance(m, Detect):
            m.inplace = self.inplace
            m.stride = torch.tensor([float(s) for s in self._layer_scales()[-1]])
            self.stride = m.stride
            m.bias_init()
        else:
            self.stride = torch.Tensor([32])
        initialize_weights(self)


def torch_safe_load(weight):
    """torch.load of a reference/Ultralytics-format checkpoint: the pickled module classes resolve to this
    package's classes because the module paths are identical (`yololite.nn.tasks.DetectionModel`, ...)."""
    ckpt = torch.load(weight, map_location="cpu", weights_only=False)
    if not isinstance(ckpt, dict):
        ckpt = {"model": ckpt.model if hasattr(ckpt, "model") else ckpt}
    return ckpt, weight


def guess_model_task(model) -> str:
    return "detect"


def attempt_load_one_weight(weight, device=None, inplace=True, fuse=False):
    """Load a single checkpoint -> (model in eval mode, ckpt dict). Reference tasks.py:499-522."""
    ckpt, weight = torch_safe_load(weight)
    args = {**DEFAULT_CFG_DICT, **(ckpt.get("train_args") or {})}
    model = (ckpt.get("ema") or ckpt["model"]).to(device).float()
    model.args = {k: v for k, v in args.items() if k in DEFAULT_CFG_DICT}
    model.pt_path = weight
    model.task = guess_model_task(model)
    if not hasattr(model, "stride"):
        model.stride = torch.tensor([32.0])
    model = model.eval()
    for m in model.modules():
        if hasattr(m, "inplace"):
            m.inplace = inplace
    return model, ckpt


def attempt_load_weights(weights, device=None, inplace=True, fuse=False):
    """Reference tasks.py:461-496 (single-model case; ensembles are out of scope)."""
    ws = weights if isinstance(weights, list) else [weights]
    if len(ws) != 1:
        raise NotImplementedError("model ensembles are out of scope")
    model, _ = attempt_load_one_weight(ws[0], device, inplace, fuse)
    return model
