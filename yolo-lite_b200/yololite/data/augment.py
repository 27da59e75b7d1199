"""LetterBox (reference yololite/data/augment.py:612-681): aspect-preserving resize + constant border.

`LetterBox` keeps the reference's host API (cv2); the predictor uses `letterbox_batch_cuda`, which does the
letterbox AND the rest of `preprocess` (BGR->RGB, HWC->CHW, float, /255; engine/predictor.py:67-85) in one CUDA
kernel on the raw uint8 images, bit-exact against the cv2 path (csrc/preprocess.cu)."""
from __future__ import annotations

import ctypes as C

import cv2
import numpy as np
import torch


class LetterBox:
    def __init__(self, new_shape=(640, 640), auto=False, scaleFill=False, scaleup=True, center=True, stride=32):
        self.new_shape = new_shape
        self.auto = auto
        self.scaleFill = scaleFill
        self.scaleup = scaleup
        self.stride = stride
        self.center = center

    def geometry(self, shape):
        """(new_unpad (w, h), (left, top, right, bottom)) for an image of `shape` = (h, w)."""
        new_shape = (self.new_shape, self.new_shape) if isinstance(self.new_shape, int) else tuple(self.new_shape)
        r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
        if not self.scaleup:
            r = min(r, 1.0)
        new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
        dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
        if self.auto:
            dw, dh = np.mod(dw, self.stride), np.mod(dh, self.stride)
        elif self.scaleFill:
            dw, dh = 0.0, 0.0
            new_unpad = (new_shape[1], new_shape[0])
        if self.center:
            dw /= 2
            dh /= 2
        top, bottom = (int(round(dh - 0.1)) if self.center else 0), int(round(dh + 0.1))
        left, right = (int(round(dw - 0.1)) if self.center else 0), int(round(dw + 0.1))
        return new_unpad, (left, top, right, bottom)

    def __call__(self, labels=None, image=None):
        img = image if image is not None else (labels or {}).get("img")
        if labels:
            raise NotImplementedError("label transformation belongs to training and is out of scope")
        shape = img.shape[:2]
        new_unpad, (left, top, right, bottom) = self.geometry(shape)
        if shape[::-1] != new_unpad:
            img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
        return cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))


_POOL = None


def _pack_pool():
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) // 2)), thread_name_prefix="yl-pack")
    return _POOL


class _Staging:
    """Grow-only pinned host buffer + device mirror for one letterbox batch (descriptors first, then pixels)."""

    def __init__(self):
        self.host = None
        self.dev = None
        self.event = None      # the previous batch's host->device copy: the pinned buffer is free once it fired

    def get(self, nbytes, device):
        if self.event is not None:
            self.event.synchronize()
        if self.host is None or self.host.numel() < nbytes or self.dev.device != device:
            cap = max(int(nbytes * 1.25), 1 << 20)
            self.host = torch.empty(cap, dtype=torch.uint8).pin_memory()
            self.dev = torch.empty(cap, dtype=torch.uint8, device=device)
        return self.host, self.dev


def letterbox_batch_cuda(images, new_shape=(640, 640), auto=False, stride=32, device=None, scaleup=True,
                         center=True, staging: _Staging | None = None, out: torch.Tensor | None = None):
    """LetterBox + `preprocess` of a list of HWC BGR uint8 images on the GPU -> (B, 3, H, W) fp32 in [0, 1].

    One pinned-memory pack of the raw bytes, ONE host->device copy (descriptors + pixels), one kernel.  All
    images must letterbox to the same canvas (true when they share a shape, or when auto=False)."""
    from .. import _C

    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = _C.init(device)
    lb = LetterBox(new_shape, auto=auto, stride=stride, scaleup=scaleup, center=center)
    geo, canvas = [], None
    for im in images:
        if not (isinstance(im, np.ndarray) and im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3):
            raise TypeError("letterbox_batch_cuda expects HWC uint8 BGR numpy images")
        (nw, nh), (left, top, right, bottom) = lb.geometry(im.shape[:2])
        hw = (nh + top + bottom, nw + left + right)
        if canvas is None:
            canvas = hw
        elif canvas != hw:
            raise ValueError(f"images letterbox to different canvases {canvas} vs {hw}; pass auto=False")
        geo.append((nw, nh, left, top))
    n = len(images)
    desc_bytes = (n * C.sizeof(_C.LbImage) + 255) // 256 * 256
    offs, total = [], desc_bytes
    for im in images:
        offs.append(total)
        total += (im.shape[0] * im.shape[1] * 3 + 255) // 256 * 256
    staging = staging or _Staging()
    host, dev = staging.get(total, device)
    hnp = host.numpy()
    descs = (_C.LbImage * n)()
    base = dev.data_ptr()
    for i, (im, (nw, nh, left, top), off) in enumerate(zip(images, geo, offs)):
        sh, sw = im.shape[:2]
        descs[i] = _C.LbImage(base + off, sh, sw, sw * 3, nw, nh, left, top, 0)

    def _pack(i):
        im, off = images[i], offs[i]
        hnp[off:off + im.shape[0] * im.shape[1] * 3] = np.ascontiguousarray(im).reshape(-1)

    # the pack into pinned memory is a plain memcpy (numpy releases the GIL): spread large batches over a few threads
    if n >= 4 and total > (8 << 20):
        list(_pack_pool().map(_pack, range(n)))
    else:
        for i in range(n):
            _pack(i)
    hnp[: n * C.sizeof(_C.LbImage)] = np.frombuffer(descs, dtype=np.uint8)
    with torch.cuda.device(device):
        dev[:total].copy_(host[:total], non_blocking=True)
        staging.event = torch.cuda.Event()
        staging.event.record()
        if out is None:
            out = torch.empty((n, 3, canvas[0], canvas[1]), dtype=torch.float32, device=device)
        assert out.shape == (n, 3, canvas[0], canvas[1]) and out.dtype == torch.float32 and out.is_contiguous()
        _C.check(lib.yl_letterbox_u8(dev.data_ptr(), n, out.data_ptr(), canvas[0], canvas[1], 114, _C.stream_ptr()),
                 "yl_letterbox_u8")
    return out
