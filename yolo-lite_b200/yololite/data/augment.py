"""LetterBox (reference yololite/data/augment.py:612-681): aspect-preserving resize + constant border.

Host-side (cv2) for now — SURVEY §8f ranks the GPU letterbox as the next component after the hot path."""
from __future__ import annotations

import cv2
import numpy as np


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
