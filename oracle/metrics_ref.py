"""oracle/metrics_ref.py — CPU restatement of the validator's matching step (TEST INFRASTRUCTURE ONLY).

  box_iou            utils/metrics.py:51-70
  match_predictions  engine/validator.py:195-233 (non-scipy branch), called per image by _process_batch :410-429

Pinned against tests/golden/val_metrics.npz, written by running the unmodified reference on seeded batches
(oracle/gen_golden.py::gen_val_metrics)."""
from __future__ import annotations

import numpy as np


def box_iou(box1: np.ndarray, box2: np.ndarray, eps=np.float32(1e-7)) -> np.ndarray:
    a1, a2 = box1[:, None, :2].astype(np.float32), box1[:, None, 2:].astype(np.float32)
    b1, b2 = box2[None, :, :2].astype(np.float32), box2[None, :, 2:].astype(np.float32)
    inter = np.clip(np.minimum(a2, b2) - np.maximum(a1, b1), 0, None).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


def match_predictions(pred_cls: np.ndarray, true_cls: np.ndarray, iou: np.ndarray, iouv: np.ndarray) -> np.ndarray:
    """(N,) predicted classes, (M,) label classes, (M, N) IoU -> (N, len(iouv)) bool."""
    correct = np.zeros((pred_cls.shape[0], iouv.shape[0]), bool)
    iou = iou * (true_cls[:, None] == pred_cls[None, :])
    for i, thr in enumerate(iouv.astype(np.float32)):
        m = np.argwhere(iou >= thr)                       # rows (label, detection)
        if not len(m):
            continue
        if len(m) > 1:
            m = m[iou[m[:, 0], m[:, 1]].argsort()[::-1]]
            m = m[np.unique(m[:, 1], return_index=True)[1]]   # best label per detection
            m = m[np.unique(m[:, 0], return_index=True)[1]]   # lowest-index detection per label
        correct[m[:, 1], i] = True
    return correct


def batch_tp(dets, counts, gt_boxes, gt_cls, offsets, iouv) -> np.ndarray:
    B, max_det, _ = dets.shape
    tp = np.zeros((B, max_det, len(iouv)), bool)
    for b in range(B):
        n, s, e = int(counts[b]), int(offsets[b]), int(offsets[b + 1])
        if n == 0 or e == s:
            continue
        iou = box_iou(gt_boxes[s:e], dets[b, :n, :4])
        tp[b, :n] = match_predictions(dets[b, :n, 5], gt_cls[s:e], iou, iouv)
    return tp
