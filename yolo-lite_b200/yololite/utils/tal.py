"""Anchor helpers with the reference's signatures (yololite/utils/tal.py:326-350).  On the inference path
both are fused into csrc/decode.cu; these host versions exist for API compatibility and tests."""
from __future__ import annotations

import torch


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Anchor points (A, 2) in grid units and stride tensor (A, 1), level-major / row-major."""
    points, svec = [], []
    assert feats is not None
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2:] if isinstance(feats, list) else (int(feats[i][0]), int(feats[i][1]))
        sx = torch.arange(w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(h, device=device, dtype=dtype) + grid_cell_offset
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        points.append(torch.stack((gx, gy), -1).view(-1, 2))
        svec.append(torch.full((h * w, 1), float(stride), dtype=dtype, device=device))
    return torch.cat(points), torch.cat(svec)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """(l, t, r, b) distances -> boxes around the anchor points."""
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)
