"""GPU NMS parity: bit-exact keep indices / detections against (a) the golden vectors produced by torchvision's
CPU kernel and by the reference's ops.non_max_suppression, (b) the oracle on larger seeded inputs, and
(c) size-independent properties at BASELINE.json's stress size (B=256, A=8400, nc=80)."""
import ast

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from yololite import _C, _ops

    _C.init(0)
    return _ops


def synth_pred(B, A, nc, seed, mean, std, dup_frac=3):
    g = np.random.default_rng(seed)
    p = np.empty((B, 4 + nc, A), np.float32)
    p[:, 0:2] = g.uniform(0, 640, (B, 2, A))
    p[:, 2:4] = g.uniform(8, 256, (B, 2, A))
    p[:, 4:] = 1.0 / (1.0 + np.exp(-g.normal(mean, std, (B, nc, A))))
    if dup_frac:
        k = A // dup_frac
        src = g.integers(0, A, k)
        p[:, 0:4, :k] = p[:, 0:4, src] + g.normal(0, 1.5, (B, 4, k)).astype(np.float32)
    return p


def run_gpu(ops, pred, **kw):
    t = torch.from_numpy(pred).cuda()
    dets, counts = ops.nms_batched(t, kw.get("conf_thres", 0.25), kw.get("iou_thres", 0.45), kw.get("classes"),
                                   kw.get("agnostic", False), kw.get("multi_label", False), kw.get("max_det", 300),
                                   kw.get("max_nms", 30000), kw.get("max_wh", 7680))
    torch.cuda.synchronize()
    dets, counts = dets.cpu().numpy(), counts.cpu().numpy()
    return [dets[i, : counts[i]] for i in range(len(counts))]


def test_nms_boxes_matches_torchvision_golden(ops, golden):
    g = golden("nms_torchvision.npz")
    for i in range(int(g["n_cases"])):
        b, s = torch.from_numpy(g[f"c{i}.boxes"]).cuda(), torch.from_numpy(g[f"c{i}.scores"]).cuda()
        keep = ops.nms_boxes(b, s, float(g[f"c{i}.thr"]))
        np.testing.assert_array_equal(keep.cpu().numpy(), g[f"c{i}.keep"], err_msg=f"case {i}")


@pytest.mark.parametrize("tag", ["single_conf25", "single_conf001", "multi_conf001", "agnostic", "classes", "maxdet"])
def test_nms_batched_matches_reference_golden(ops, golden, tag):
    g = golden("nms_reference.npz")
    kw = ast.literal_eval(str(g[f"{tag}.kw"]))
    res = run_gpu(ops, g[f"{tag}.pred"], **kw)
    assert [len(r) for r in res] == g[f"{tag}.counts"].tolist()
    np.testing.assert_array_equal(np.concatenate(res, 0), g[f"{tag}.dets"])


CASES = [
    dict(tag="sparse_single", B=4, A=8400, mean=-12.0, std=1.5, kw=dict(conf_thres=0.001, iou_thres=0.7)),
    dict(tag="dense_single", B=3, A=8400, mean=-5.0, std=2.0, kw=dict(conf_thres=0.001, iou_thres=0.7)),
    dict(tag="dense_single_tight", B=2, A=8400, mean=-5.0, std=2.0, kw=dict(conf_thres=0.001, iou_thres=0.3), dup=1),
    dict(tag="sparse_multi", B=3, A=8400, mean=-12.0, std=1.5, kw=dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),
    dict(tag="mid_multi_gt_sortcap", B=2, A=8400, mean=-9.5, std=1.5, kw=dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),
    dict(tag="dense_multi_topk", B=2, A=8400, mean=-5.0, std=2.0, kw=dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),
    dict(tag="predict_defaults", B=8, A=8400, mean=-6.0, std=2.5, kw=dict(conf_thres=0.25, iou_thres=0.7)),
    dict(tag="odd_A_scalar_path", B=2, A=5041, mean=-6.0, std=2.5, kw=dict(conf_thres=0.05, iou_thres=0.5)),
    dict(tag="nothing_passes", B=2, A=8400, mean=-20.0, std=0.1, kw=dict(conf_thres=0.25, iou_thres=0.7)),
]


@pytest.mark.parametrize("case", CASES, ids=[c["tag"] for c in CASES])
def test_nms_batched_matches_oracle(ops, case):
    from oracle import nms_ref

    pred = synth_pred(case["B"], case["A"], 80, 1234, case["mean"], case["std"], case.get("dup", 3))
    res = run_gpu(ops, pred, **case["kw"])
    ref = nms_ref.non_max_suppression(pred, **case["kw"])
    assert [len(r) for r in res] == [len(r) for r in ref]
    for i, (a, b) in enumerate(zip(res, ref)):
        np.testing.assert_array_equal(a, b, err_msg=f"image {i}")


def test_nms_stress_properties_full_size(ops):
    """BASELINE config 5 at full size: B=256, A=8400, nc=80, conf=0.001, iou=0.7, max_det=300."""
    B = 256
    g = torch.Generator(device="cuda").manual_seed(2)
    pred = torch.empty((B, 84, 8400), device="cuda")
    pred[:, 0:2] = torch.rand((B, 2, 8400), device="cuda", generator=g) * 640
    pred[:, 2:4] = torch.rand((B, 2, 8400), device="cuda", generator=g) * 248 + 8
    pred[:, 4:] = torch.sigmoid(torch.randn((B, 80, 8400), device="cuda", generator=g) * 1.5 - 12.0)
    dets, counts = ops.nms_batched(pred, 0.001, 0.7)
    dets2, counts2 = ops.nms_batched(pred, 0.001, 0.7)
    torch.cuda.synchronize()
    assert torch.equal(dets, dets2) and torch.equal(counts, counts2)          # deterministic
    c = counts.cpu().numpy()
    d = dets.cpu().numpy()
    assert (c >= 0).all() and (c <= 300).all() and c.max() > 0
    amax = pred[:, 4:].amax(1)
    ncand = (amax > 0.001).sum(1).cpu().numpy()
    assert (c <= ncand).all()
    for i in range(0, B, 17):
        k = d[i, : c[i]]
        assert (np.diff(k[:, 4]) <= 0).all()                                    # descending confidence
        assert (k[:, 4] > 0.001).all()
        assert (k[:, 2] >= k[:, 0]).all() and (k[:, 3] >= k[:, 1]).all()
        # survivors of the same class never overlap above the threshold (float64 check with slack)
        for cls in np.unique(k[:, 5]):
            b = k[k[:, 5] == cls][:, :4].astype(np.float64)
            if len(b) < 2:
                continue
            x1 = np.maximum(b[:, None, 0], b[None, :, 0]); y1 = np.maximum(b[:, None, 1], b[None, :, 1])
            x2 = np.minimum(b[:, None, 2], b[None, :, 2]); y2 = np.minimum(b[:, None, 3], b[None, :, 3])
            inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
            area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
            iou = inter / (area[:, None] + area[None, :] - inter)
            np.fill_diagonal(iou, 0)
            assert iou.max() <= 0.7 + 1e-4
    # spot-check bit-exactness against the oracle on a few images of the big batch
    from oracle import nms_ref

    idx = [0, 101, 255]
    ref = nms_ref.non_max_suppression(pred[idx].cpu().numpy(), conf_thres=0.001, iou_thres=0.7)
    for j, i in enumerate(idx):
        np.testing.assert_array_equal(d[i, : c[i]], ref[j])


@pytest.mark.parametrize("multi_label", [False, True])
def test_nms_stress_dense_full_size_vs_oracle(ops, multi_label):
    """BASELINE config 5, DENSE regime at full size (B=256, A=8400, nc=80, conf=0.001): every anchor is a candidate
    (8400 per image; multi-label expands to > 30000 pairs per image, i.e. the max_nms sort-and-cut path, ops.py:254-255).
    32 images spread over the batch are compared bit-for-bit with the oracle."""
    from oracle import nms_ref

    B = 256
    g = torch.Generator(device="cuda").manual_seed(2)
    pred = torch.empty((B, 84, 8400), device="cuda")
    pred[:, 0:2] = torch.rand((B, 2, 8400), device="cuda", generator=g) * 640
    pred[:, 2:4] = torch.rand((B, 2, 8400), device="cuda", generator=g) * 248 + 8
    pred[:, 4:] = torch.sigmoid(torch.randn((B, 80, 8400), device="cuda", generator=g) * 2.0 - 5.0)
    k = 8400 // 3                               # clusters of near-duplicate boxes so that suppression happens
    src = torch.randint(0, 8400, (k,), device="cuda", generator=g)
    pred[:, 0:4, :k] = pred[:, 0:4, src] + torch.randn((B, 4, k), device="cuda", generator=g) * 1.5
    ncand = (pred[:, 4:].amax(1) > 0.001).sum(1)
    assert int(ncand.min()) >= 8300                                             # dense: (almost) every anchor passes
    if multi_label:
        assert int((pred[:, 4:] > 0.001).sum((1, 2)).min()) > 30000            # the max_nms path is taken
    dets, counts = ops.nms_batched(pred, 0.001, 0.7, multi_label=multi_label, max_det=300)
    torch.cuda.synchronize()
    c, d = counts.cpu().numpy(), dets.cpu().numpy()
    assert (c == 300).all()
    idx = list(range(0, B, 8))
    assert len(idx) == 32
    ref = nms_ref.non_max_suppression(pred[idx].cpu().numpy(), conf_thres=0.001, iou_thres=0.7, multi_label=multi_label,
                                      max_det=300)
    for j, i in enumerate(idx):
        np.testing.assert_array_equal(d[i, : c[i]], ref[j], err_msg=f"image {i}")


def test_nms_classwise_path_and_its_fallbacks_vs_oracle(ops):
    """The per-class select path (class offsets separate the classes) and every condition that sends an image back to
    the general path, one image each in ONE batch, bit-exact against the oracle:
      0  ordinary image (class-wise)            1  boxes whose x extent exceeds max_wh: classes DO overlap after the offset
      2  one class with > 256 candidates        3  empty image           4  a NaN box           5  small max_wh (extent test fails)
    plus the same batch with agnostic NMS, a class list and tiny max_det / max_nms."""
    from oracle import nms_ref

    g = np.random.default_rng(31)
    B, A, nc = 5, 1500, 80
    p = np.zeros((B, 4 + nc, A), np.float32)
    p[:, 0:2] = g.uniform(0, 640, (B, 2, A))
    p[:, 2:4] = g.uniform(8, 200, (B, 2, A))
    k = A // 2                                                      # near-duplicates so that suppression happens
    src = g.integers(0, A, k)
    p[:, 0:4, :k] = p[:, 0:4, src] + g.normal(0, 2.0, (B, 4, k)).astype(np.float32)
    logits = g.normal(-3.0, 2.0, (B, nc, A))
    p[:, 4:] = 1.0 / (1.0 + np.exp(-logits))
    # image 1: centres spread over 0..16000 px with huge boxes -> different classes overlap after `+ cls * 7680`
    p[1, 0] = g.uniform(0, 16000, A)
    p[1, 2] = g.uniform(2000, 9000, A)
    p[1, 3] = g.uniform(300, 640, A)
    p[1, 1] = g.uniform(200, 400, A)
    # image 2: everything is class 7 (one segment of ~1500 candidates)
    p[2, 4:] = 0.0
    p[2, 4 + 7] = g.uniform(0.3, 0.9, A)
    # image 3: nothing passes
    p[3, 4:] = 0.01
    # image 4: a NaN coordinate among ordinary boxes
    p[4, 0, 17] = np.nan
    for kw in (dict(conf_thres=0.25, iou_thres=0.6), dict(conf_thres=0.25, iou_thres=0.6, agnostic=True),
               dict(conf_thres=0.3, iou_thres=0.5, classes=[0, 7, 33]), dict(conf_thres=0.25, iou_thres=0.6, max_det=7),
               dict(conf_thres=0.25, iou_thres=0.6, max_nms=100), dict(conf_thres=0.25, iou_thres=0.6, max_wh=300),
               dict(conf_thres=0.05, iou_thres=0.7, multi_label=True)):
        res = run_gpu(ops, p, **kw)
        ref = nms_ref.non_max_suppression(p, **kw)
        assert [len(r) for r in res] == [len(r) for r in ref], kw
        for i, (a, r) in enumerate(zip(res, ref)):
            np.testing.assert_array_equal(a, r, err_msg=f"{kw} image {i}")
    assert len(ref[0]) > 20


def test_nms_argument_errors(ops):
    from yololite import _C

    pred = torch.zeros((1, 84, 64), device="cuda")
    with pytest.raises(_C.YLError):
        ops.nms_batched(pred, 1.5, 0.5)
    with pytest.raises(_C.YLError):
        ops.nms_batched(pred, 0.5, -0.1)
