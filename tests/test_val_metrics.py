"""Validator metric: box_iou + match_predictions + ap_per_class.

CPU: oracle (oracle/metrics_ref.py) and the host `ap_per_class` against goldens written by the unmodified reference.
GPU: yl_match_predictions through the C-ABI, bit-exact (bool) against goldens and oracle; DetectionValidator end to end."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]

from oracle import metrics_ref  # noqa: E402
from oracle.gen_golden import val_case  # noqa: E402

GOLD = np.load(ROOT / "tests" / "golden" / "val_metrics.npz")
IOUV = np.linspace(0.5, 0.95, 10).astype(np.float32)
SEEDS = [int(s) for s in GOLD["seeds"]]


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_matching_equals_reference(seed):
    dets, counts, gtb, gtc, offs = val_case(seed)
    tp = metrics_ref.batch_tp(dets, counts, gtb, gtc, offs, torch.linspace(0.5, 0.95, 10).numpy())
    assert np.array_equal(tp, GOLD[f"tp_{seed}"])
    assert tp.any() and not tp.all()


@pytest.mark.parametrize("seed", SEEDS)
def test_host_ap_per_class_equals_reference(seed):
    from yololite.utils.metrics import ap_per_class

    dets, counts, gtb, gtc, offs = val_case(seed)
    tps = GOLD[f"tp_{seed}"]
    sel = [(b, i) for b in range(dets.shape[0]) for i in range(counts[b])]
    tp = np.asarray([tps[b, i] for b, i in sel]).reshape(-1, 10)
    conf = np.asarray([dets[b, i, 4] for b, i in sel])
    pc = np.asarray([dets[b, i, 5] for b, i in sel])
    r = ap_per_class(tp, conf, pc, gtc)
    for k, name in zip(range(7), ("tpn", "fpn", "p", "r", "f1", "ap", "cls")):
        np.testing.assert_allclose(np.asarray(r[k]), GOLD[f"{name}_{seed}"], rtol=1e-12, atol=1e-12, err_msg=name)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_cuda_match_predictions_equals_reference(seed):
    from yololite.utils.metrics import match_predictions_batched

    dets, counts, gtb, gtc, offs = val_case(seed)
    tp = match_predictions_batched(torch.from_numpy(dets).cuda(), torch.from_numpy(counts).cuda(), torch.from_numpy(gtb),
                                   torch.from_numpy(gtc), offs.tolist(), torch.linspace(0.5, 0.95, 10))
    assert np.array_equal(tp.cpu().numpy(), GOLD[f"tp_{seed}"])


@pytest.mark.gpu
def test_cuda_match_predictions_large_batch_vs_oracle():
    """BASELINE-sized: 64 images x 300 detections, up to 11 labels each, plus a crowded image (400 labels)."""
    from yololite.utils.metrics import match_predictions_batched

    dets, counts, gtb, gtc, offs = val_case(99, B=64, max_det=300, nc=80)
    want = metrics_ref.batch_tp(dets, counts, gtb, gtc, offs, torch.linspace(0.5, 0.95, 10).numpy())
    got = match_predictions_batched(torch.from_numpy(dets).cuda(), torch.from_numpy(counts).cuda(), torch.from_numpy(gtb),
                                    torch.from_numpy(gtc), offs.tolist(), torch.linspace(0.5, 0.95, 10))
    assert np.array_equal(got.cpu().numpy(), want)
    g = np.random.default_rng(5)
    xy = g.integers(0, 600 * 64, (400, 2)) / 64.0
    lab = np.concatenate([xy, xy + g.integers(8 * 64, 60 * 64, (400, 2)) / 64.0], 1).astype(np.float32)
    cl = g.integers(0, 3, (400,)).astype(np.float32)
    d = np.zeros((1, 300, 6), np.float32)
    d[0, :, :4] = lab[g.integers(0, 400, 300)] + (g.integers(-4 * 64, 4 * 64, (300, 4)) / 64.0).astype(np.float32)
    d[0, :, 4] = np.sort(g.random(300).astype(np.float32))[::-1]
    d[0, :, 5] = g.integers(0, 3, 300)
    c = np.asarray([300], np.int32)
    o = np.asarray([0, 400], np.int32)
    want = metrics_ref.batch_tp(d, c, lab, cl, o, torch.linspace(0.5, 0.95, 10).numpy())
    got = match_predictions_batched(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(lab),
                                    torch.from_numpy(cl), o.tolist(), torch.linspace(0.5, 0.95, 10))
    assert np.array_equal(got.cpu().numpy(), want) and want.any()


@pytest.mark.gpu
def test_validator_end_to_end_on_synthetic_batches():
    """YOLOLite.val over an in-memory dataloader: labels = the model's own confident detections, so the metrics
    must come out (near) perfect; checks the whole validator plumbing on the GPU."""
    from oracle.weights import fill_state_dict_
    from yololite import YOLOLite

    yl = YOLOLite("yolo11n.yaml")
    fill_state_dict_(yl.model)
    x = torch.rand(4, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    res = yl.predict(x, imgsz=64, conf=0.05, iou=0.5, verbose=False, device=0)
    cls, boxes, bidx = [], [], []
    for i, r in enumerate(res):
        b = r.boxes.data.cpu()[:6]
        if len(b):
            xyxy = b[:, :4]
            xywh = torch.stack([(xyxy[:, 0] + xyxy[:, 2]) / 2, (xyxy[:, 1] + xyxy[:, 3]) / 2, xyxy[:, 2] - xyxy[:, 0],
                                xyxy[:, 3] - xyxy[:, 1]], 1) / 64.0
            cls.append(b[:, 5:6]); boxes.append(xywh); bidx.append(torch.full((len(b),), float(i)))
    assert cls, "the seeded model produced no detections at conf 0.05"
    batch = {"img": x.clone(), "cls": torch.cat(cls), "bboxes": torch.cat(boxes), "batch_idx": torch.cat(bidx),
             "ori_shape": [(64, 64)] * 4, "ratio_pad": [None] * 4, "im_file": [f"im{i}.jpg" for i in range(4)]}
    m = yl.val(dataloader=[batch], conf=0.001, iou=0.5, verbose=False, device=0)
    d = m.results_dict
    assert set(d) >= {"metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)", "fitness"}
    # labels are the model's own boxes: a matched pair has IoU 1, so every IoU threshold sees the same matches
    assert d["metrics/mAP50(B)"] > 0.5 and abs(d["metrics/mAP50(B)"] - d["metrics/mAP50-95(B)"]) < 1e-6, d
    assert d["metrics/recall(B)"] > 0.5, d
