"""Results / Boxes containers (reference yololite/engine/results.py:42, :443-581), data layout
(n, 6) = [x1, y1, x2, y2, conf, cls].  Plotting, saving and JSON export are visualisation/I-O, out of scope."""
from __future__ import annotations

import numpy as np
import torch

from ..utils import ops


class BaseTensor:
    def __init__(self, data, orig_shape):
        assert isinstance(data, (torch.Tensor, np.ndarray))
        self.data = data
        self.orig_shape = orig_shape

    @property
    def shape(self):
        return self.data.shape

    def cpu(self):
        return self if isinstance(self.data, np.ndarray) else self.__class__(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return self if isinstance(self.data, np.ndarray) else self.__class__(self.data.cpu().numpy(), self.orig_shape)

    def cuda(self):
        return self.__class__(torch.as_tensor(self.data).cuda(), self.orig_shape)

    def to(self, *a, **kw):
        return self.__class__(torch.as_tensor(self.data).to(*a, **kw), self.orig_shape)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.__class__(self.data[idx], self.orig_shape)


class Boxes(BaseTensor):
    """Detections of one image: xyxy in original-image pixels, confidence, class."""

    def __init__(self, boxes, orig_shape):
        if boxes.ndim == 1:
            boxes = boxes[None, :]
        assert boxes.shape[-1] in {6, 7}, f"expected 6 or 7 values but got {boxes.shape[-1]}"
        super().__init__(boxes, orig_shape)
        self.is_track = boxes.shape[-1] == 7

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return self.data[:, -3] if self.is_track else None

    @property
    def xywh(self):
        return ops.xyxy2xywh(self.xyxy)

    @property
    def xyxyn(self):
        xyxy = self.xyxy.clone() if isinstance(self.xyxy, torch.Tensor) else np.copy(self.xyxy)
        xyxy[..., [0, 2]] /= self.orig_shape[1]
        xyxy[..., [1, 3]] /= self.orig_shape[0]
        return xyxy

    @property
    def xywhn(self):
        xywh = ops.xyxy2xywh(self.xyxy)
        xywh[..., [0, 2]] /= self.orig_shape[1]
        xywh[..., [1, 3]] /= self.orig_shape[0]
        return xywh


class BatchDetections:
    """Device-resident NMS output of one batch, (B, max_det, 6) + counts (B,): the per-image row counts are read
    from the device on first use (ONE host synchronisation per batch), so `predict()` itself never blocks and
    the next batch's upload overlaps this batch's kernels."""

    def __init__(self, dets: torch.Tensor, counts: torch.Tensor):
        self.dets, self.counts, self._n = dets, counts, None
        self.ready = torch.cuda.Event()      # recorded by the producer once dets / counts are final

    def n(self):
        if self._n is None:
            cur = torch.cuda.current_stream(self.dets.device)
            cur.wait_event(self.ready)       # stream-level wait: only THIS batch's work, not whatever was enqueued later
            self.dets.record_stream(cur)
            self.counts.record_stream(cur)
            self._n = self.counts.tolist()
        return self._n

    def boxes(self, i: int) -> torch.Tensor:
        return self.dets[i, : self.n()[i]]

    def host(self):
        """(dets (B, max_det, 6) float32, counts (B,) int32) as numpy arrays: the whole batch's detections in ONE
        device->host copy per tensor (rows beyond counts[i] are padding).  Cached."""
        if getattr(self, "_host", None) is None:
            cur = torch.cuda.current_stream(self.dets.device)
            cur.wait_event(self.ready)
            self.dets.record_stream(cur)
            self.counts.record_stream(cur)
            d = self.dets.to("cpu", non_blocking=True)
            c = self.counts.to("cpu", non_blocking=True)
            cur.synchronize()
            self._host = (d.numpy(), c.numpy())
            if self._n is None:
                self._n = self._host[1].tolist()
        return self._host


class Results:
    def __init__(self, orig_img, path, names, boxes=None, speed=None, lazy=None):
        """`lazy` = (BatchDetections, image index): `.boxes` materialises on first access."""
        self.orig_img = orig_img
        self.orig_shape = orig_img.shape[:2] if orig_img is not None else None
        self._boxes = Boxes(boxes, self.orig_shape) if boxes is not None else None
        self._lazy = lazy
        self._batch = lazy[0] if lazy is not None else None
        self.masks = self.probs = self.keypoints = self.obb = None
        self.speed = speed or {"preprocess": None, "inference": None, "postprocess": None}
        self.names = names
        self.path = path
        self.save_dir = None

    @property
    def boxes(self):
        if self._boxes is None and self._lazy is not None:
            batch, i = self._lazy
            self._boxes = Boxes(batch.boxes(i), self.orig_shape)
            self._lazy = None
        return self._boxes

    @boxes.setter
    def boxes(self, value):
        self._boxes, self._lazy = value, None

    @property
    def batch(self):
        """The device-resident detections of the whole batch this result belongs to (`BatchDetections`: `.host()` reads
        all of them with one copy), or None for results constructed from explicit boxes."""
        return self._batch

    def __len__(self):
        return len(self.boxes) if self.boxes is not None else 0

    def cpu(self):
        r = Results(self.orig_img, self.path, self.names, speed=self.speed)
        r.boxes = self.boxes.cpu() if self.boxes is not None else None
        return r

    def numpy(self):
        r = Results(self.orig_img, self.path, self.names, speed=self.speed)
        r.boxes = self.boxes.numpy() if self.boxes is not None else None
        return r

    def verbose(self):
        if not len(self):
            return "(no detections), "
        cls = self.boxes.cls
        out = ""
        for c in (cls.unique() if isinstance(cls, torch.Tensor) else np.unique(cls)):
            n = int((cls == c).sum())
            out += f"{n} {self.names[int(c)]}{'s' * (n > 1)}, "
        return out

    def summary(self, decimals=5):
        rows = []
        data = self.boxes.numpy().data if self.boxes is not None else np.zeros((0, 6))
        for r in data:
            rows.append({"name": self.names[int(r[5])], "class": int(r[5]), "confidence": round(float(r[4]), decimals),
                         "box": {k: round(float(v), decimals) for k, v in zip(("x1", "y1", "x2", "y2"), r[:4])}})
        return rows
