"""oracle/letterbox_ref.py — CPU restatement of the reference's image preprocess (TEST INFRASTRUCTURE ONLY: imported
by tests/ and __graft_entry__.smoke(), never by the product package).

Path restated (reference file:line):
  LetterBox.__call__                     data/augment.py:612-681   geometry, cv2.resize(INTER_LINEAR), copyMakeBorder(114)
  DetectionPredictor.pre_transform       engine/predictor.py:87-103 (auto = same shapes and .pt model)
  DetectionPredictor.preprocess          engine/predictor.py:67-85  stack, BGR->RGB, HWC->CHW, float, /255

`cv2.resize` is a third-party binary (opencv-python, unpinned by the reference); its 8-bit INTER_LINEAR algorithm is
restated here from its published source (imgproc/resize.cpp: fixed-point 11-bit coefficients, horizontal pass into
int32, vertical pass `((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2`) and PINNED against the installed cv2 by
tests/test_letterbox.py (random shapes) and against tests/golden/letterbox.npz, which oracle/gen_golden.py wrote
by running the unmodified reference LetterBox + preprocess.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _axis_tables(ssize: int, dsize: int):
    """Source index pair and fixed-point weights of every destination coordinate along one axis.

    Horizontal semantics (resize.cpp, resizeGeneric / HResizeLinear): sx < 0 -> (0, fx = 0); sx >= ssize-1 ->
    (ssize-1, fx = 0).  The vertical pass instead clamps the two ROW indices and keeps the weights; `clamp_rows`
    below selects that behaviour."""
    inv_scale = np.float64(dsize) / np.float64(ssize)
    scale = np.float64(1.0) / inv_scale
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _coefs(f: np.ndarray):
    c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)   # cvRound = round half to even
    c1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    return np.clip(c0, -32768, 32767), np.clip(c1, -32768, 32767)


def resize_linear_u8(src: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for HWC uint8, bit-exact."""
    assert src.dtype == np.uint8 and src.ndim == 3
    sh, sw, _ = src.shape
    dw, dh = int(dsize_wh[0]), int(dsize_wh[1])
    # horizontal tables
    sx, fx = _axis_tables(sw, dw)
    lo = sx < 0
    fx = np.where(lo, np.float32(0), fx)
    sx = np.where(lo, 0, sx)
    hi = sx >= sw - 1
    fx = np.where(hi, np.float32(0), fx)
    sx = np.where(hi, sw - 1, sx)
    a0, a1 = _coefs(fx)
    sx1 = np.minimum(sx + 1, sw - 1)          # weight is 0 wherever this clamp acts
    # vertical tables: rows clamped, weights kept
    sy, fy = _axis_tables(sh, dh)
    b0, b1 = _coefs(fy)
    y0 = np.clip(sy, 0, sh - 1)
    y1 = np.clip(sy + 1, 0, sh - 1)
    s32 = src.astype(np.int32)
    # horizontal pass of the two source rows of every destination row
    r0 = s32[y0][:, sx, :] * a0[None, :, None] + s32[y0][:, sx1, :] * a1[None, :, None]
    r1 = s32[y1][:, sx, :] * a0[None, :, None] + s32[y1][:, sx1, :] * a1[None, :, None]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_geometry(shape_hw, new_shape=(640, 640), auto=False, scaleFill=False, scaleup=True, center=True,
                       stride=32):
    """data/augment.py:640-665: returns (new_unpad (w, h), (left, top, right, bottom))."""
    new_shape = (new_shape, new_shape) if isinstance(new_shape, int) else tuple(new_shape)
    r = min(new_shape[0] / shape_hw[0], new_shape[1] / shape_hw[1])
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(shape_hw[1] * r)), int(round(shape_hw[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    elif scaleFill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return new_unpad, (left, top, right, bottom)


def letterbox(img: np.ndarray, new_shape=(640, 640), auto=False, scaleFill=False, scaleup=True, center=True,
              stride=32, value=114) -> np.ndarray:
    """LetterBox.__call__(image=img) for an HWC uint8 BGR image."""
    new_unpad, (left, top, right, bottom) = letterbox_geometry(img.shape[:2], new_shape, auto, scaleFill, scaleup,
                                                               center, stride)
    if img.shape[:2][::-1] != new_unpad:
        img = resize_linear_u8(img, new_unpad)
    h, w, c = img.shape
    out = np.full((h + top + bottom, w + left + right, c), value, dtype=np.uint8)
    out[top:top + h, left:left + w] = img
    return out


def preprocess(images, new_shape=(640, 640), auto=True, stride=32) -> np.ndarray:
    """predictor.preprocess for a list of same-shape HWC BGR uint8 images -> (B, 3, H, W) float32 in [0, 1]."""
    lb = np.stack([letterbox(im, new_shape, auto=auto, stride=stride) for im in images])
    chw = np.ascontiguousarray(lb[..., ::-1].transpose(0, 3, 1, 2))
    return chw.astype(np.float32) / np.float32(255)
