"""oracle/nms_ref.py — TEST INFRASTRUCTURE, not product code.

numpy restatement of `ops.non_max_suppression` (reference yololite/utils/ops.py:138-273) on top of the C
restatement of torchvision's CPU NMS (oracle/nms_ref.c).  Differences from the reference, all deliberate:
  * no wall-clock early exit (ops.py:207,269-271) — the reference silently drops images when slow;
  * the max_nms truncation (ops.py:254-255) uses a *stable* descending sort (the reference's argsort is
    unstable, so ties at the cut are implementation-defined there);
  * the caller's tensor is not mutated.
Parity pin: tests/golden/nms_*.npz hold outputs of the reference itself and of torchvision 0.26 CPU
(oracle/gen_golden.py); tests/test_oracle.py checks this file against them.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


def _load():
    global _lib
    if _lib is None:
        so = _HERE / "_build" / "libnmsref.so"
        if not so.exists() or so.stat().st_mtime < (_HERE / "nms_ref.c").stat().st_mtime:
            subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
        lib = ctypes.CDLL(str(so))
        lib.yl_ref_nms.restype = ctypes.c_int64
        lib.yl_ref_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                   ctypes.c_void_p, ctypes.c_int64]
        _lib = lib
    return _lib


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float, max_keep: int = 0) -> np.ndarray:
    """torchvision.ops.nms semantics (ops.py:265). boxes (n,4) xyxy fp32, scores (n,) fp32 -> int64 keep."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = _load().yl_ref_nms(boxes.ctypes.data, scores.ctypes.data, n, float(iou_threshold), keep.ctypes.data,
                           int(max_keep))
    return keep[:k].copy()


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """ops.py:372-389 in fp32: xy -/+ wh/2."""
    x = x.astype(np.float32)
    y = np.empty_like(x)
    half = x[..., 2:4] / np.float32(2)
    y[..., 0:2] = x[..., 0:2] - half
    y[..., 2:4] = x[..., 0:2] + half
    return y


def non_max_suppression(prediction: np.ndarray, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300, nc=0, max_nms=30000, max_wh=7680):
    """ops.py:138-273 for detection (nm = 0, not rotated, no autolabels). prediction: (B, 4+nc, A) fp32.

    Returns a list of (n_i, 6) fp32 arrays [x1, y1, x2, y2, conf, cls], descending confidence.
    """
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1                     # ops.py:186-187
    p = np.asarray(prediction, dtype=np.float32)
    bs = p.shape[0]
    nc = nc or (p.shape[1] - 4)                                             # ops.py:200
    assert p.shape[1] == 4 + nc, "mask channels are out of scope"
    conf32 = np.float32(conf_thres)
    xc = p[:, 4:].max(axis=1) > conf32                                      # ops.py:203 (strict >)
    multi_label = bool(multi_label) and nc > 1                              # ops.py:208
    p = p.transpose(0, 2, 1)                                                # ops.py:210 -> (B, A, 4+nc)
    out = [np.zeros((0, 6), np.float32) for _ in range(bs)]                 # ops.py:218
    for xi in range(bs):                                                    # ops.py:219
        x = p[xi][xc[xi]]                                                   # ops.py:222 (anchor order)
        if not x.shape[0]:
            continue
        box = xywh2xyxy(x[:, :4])                                           # ops.py:213
        cls = x[:, 4:]
        if multi_label:                                                     # ops.py:239-241 (row-major i, j)
            i, j = np.nonzero(cls > conf32)
            x = np.concatenate([box[i], cls[i, j][:, None], j[:, None].astype(np.float32)], 1)
        else:                                                               # ops.py:242-244 (first max wins)
            j = cls.argmax(1)
            conf = cls[np.arange(cls.shape[0]), j]
            x = np.concatenate([box, conf[:, None], j[:, None].astype(np.float32)], 1)[conf > conf32]
        if classes is not None:                                             # ops.py:247-248
            x = x[np.isin(x[:, 5], np.asarray(classes, dtype=np.float32))]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:                                                     # ops.py:254-255
            x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        c = x[:, 5:6] * np.float32(0 if agnostic else max_wh)               # ops.py:258
        boxes = (x[:, :4] + c).astype(np.float32)                           # ops.py:264 (fp32 add)
        keep = nms(boxes, x[:, 4], iou_thres, max_keep=max_det)             # ops.py:265-266
        out[xi] = np.ascontiguousarray(x[keep[:max_det]], dtype=np.float32)  # ops.py:268
    return out
