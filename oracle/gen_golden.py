"""oracle/gen_golden.py — regenerate tests/golden/*.npz from the reference itself (build container only).

    python oracle/gen_golden.py --write [--only real scale_boxes ...]

Fixtures (all seeded; weights are the name-keyed values of oracle/weights.py, so they are not stored):
  model_yolo11{n,s,m}.npz   reference DetectionModel forward on a (2,3,64,64) batch: y, raw head maps
  modules.npz               reference nn modules (Conv, DWConv, Bottleneck, C3k, C3k2 x2, SPPF, Attention,
                            PSABlock, C2PSA, Detect, DFL) on small inputs
  nms_torchvision.npz       torchvision.ops.nms (CPU, the arithmetic behind ops.py:265) known-answer cases
  nms_reference.npz         reference ops.non_max_suppression (max_time_img=1e9) on synthetic predictions
  val_metrics.npz           reference box_iou + DetectionValidator.match_predictions (validator.py:195-233) per image
                            and utils.metrics.ap_per_class on seeded synthetic detections / labels
  ckpt_tiny.pt / ckpt_tiny_out.npz   a checkpoint PICKLED BY THE REFERENCE ({"model": DetectionModel.half(), ...},
                            engine/trainer.py:360-388 layout) of a narrow yolo11 (width 0.125, nc 16) + the reference's
                            fp32 CPU output for it: the checkpoint-ingest fixture
  letterbox.npz             reference LetterBox (data/augment.py:612-681) + predictor.preprocess arithmetic
                            (engine/predictor.py:67-85) on seeded uint8 images (only seeds + outputs stored)
  real_images.npz           BASELINE configs 1-2 on the reference's own fixtures with seeded weights (fp16-rounded, the
                            reference pickles `.half()`): boats.jpg through YOLOLite.predict (auto letterbox 384x640, conf .25)
                            and coco8 val through YOLOLite.val (rect batch 4x3x672x672, multi-label NMS conf .001):
                            JPEG bytes + labels in, the reference's preprocessed batch (crc + sample), top-score rows of y,
                            NMS output, scale_boxes'd detections, tp matrices and metrics out
  scale_boxes.npz           reference ops.scale_boxes / clip_boxes (utils/ops.py:66-98,276-295) on seeded boxes, incl. the
                            round(x - 0.1) pad and explicit ratio_pad
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.ref_harness import build_reference_model, import_reference  # noqa: E402
from oracle.weights import fill_state_dict_  # noqa: E402

GOLD = ROOT / "tests" / "golden"
REF = "/root/reference"


def seeded(shape, seed, lo=0.0, hi=1.0):
    g = np.random.default_rng(seed)
    return torch.from_numpy(g.uniform(lo, hi, shape).astype(np.float32))


def gen_models():
    for scale in "nsm":
        m = build_reference_model(scale)
        fill_state_dict_(m)
        x = seeded((2, 3, 64, 64), 100)
        with torch.no_grad():
            y, raw = m(x)
        shapes = {k: np.asarray(v.shape, dtype=np.int64) for k, v in m.state_dict().items()}
        np.savez_compressed(GOLD / f"model_yolo11{scale}.npz", x=x.numpy(), y=y.numpy(),
                            **{f"raw{i}": r.numpy() for i, r in enumerate(raw)},
                            keys=np.array(list(shapes)), shapes=np.array([list(s) + [0] * (4 - len(s)) for s in shapes.values()]),
                            ndims=np.array([len(s) for s in shapes.values()]))
        print("model", scale, tuple(y.shape), float(y[:, 4:].max()))


def gen_modules():
    import_reference()
    from yololite.nn.modules.block import C2PSA, C3k, C3k2, DFL, SPPF, Attention, Bottleneck, PSABlock
    from yololite.nn.modules.conv import Conv, DWConv
    from yololite.nn.modules.head import Detect

    out = {}

    def run(tag, mod, x):
        fill_state_dict_(mod.eval())
        with torch.no_grad():
            y = mod(x)
        out[f"{tag}.x"] = x.numpy()
        out[f"{tag}.y"] = y.numpy()
        print("module", tag, tuple(x.shape), "->", tuple(y.shape))

    run("conv_k3s2", Conv(16, 32, 3, 2), seeded((2, 16, 20, 24), 1, -1, 1))
    run("conv_k1", Conv(48, 64, 1, 1), seeded((2, 48, 12, 12), 2, -1, 1))
    run("conv_k3s1_noact", Conv(32, 16, 3, 1, act=False), seeded((1, 32, 10, 14), 3, -1, 1))
    run("conv_stem", Conv(3, 16, 3, 2), seeded((2, 3, 32, 32), 4, 0, 1))
    run("dwconv", DWConv(64, 64, 3), seeded((2, 64, 10, 10), 5, -1, 1))
    run("bottleneck", Bottleneck(32, 32, True), seeded((2, 32, 12, 12), 6, -1, 1))
    run("c3k", C3k(64, 64, 2), seeded((1, 64, 10, 10), 7, -1, 1))
    run("c3k2_plain", C3k2(64, 128, 1, False, 0.25), seeded((2, 64, 12, 12), 8, -1, 1))
    run("c3k2_c3k", C3k2(128, 128, 1, True), seeded((1, 128, 8, 8), 9, -1, 1))
    run("sppf", SPPF(128, 128, 5), seeded((2, 128, 9, 11), 10, -1, 1))
    run("attention", Attention(128, num_heads=2, attn_ratio=0.5), seeded((2, 128, 6, 7), 11, -1, 1))
    run("psablock", PSABlock(128, 0.5, 2), seeded((1, 128, 5, 5), 12, -1, 1))
    run("c2psa", C2PSA(256, 256, 1), seeded((1, 256, 6, 6), 13, -1, 1))
    # DFL: input (b, 64, a)
    dfl = DFL(16).eval()
    xd = seeded((2, 64, 50), 14, -3, 3)
    with torch.no_grad():
        out["dfl.x"], out["dfl.y"] = xd.numpy(), dfl(xd).numpy()
    # Detect on three levels (8x8, 4x4, 2x2), strides set like DetectionModel does
    det = Detect(80, (64, 128, 256)).eval()
    det.stride = torch.tensor([8.0, 16.0, 32.0])
    fill_state_dict_(det)
    feats = [seeded((2, c, s, s), 20 + i, -1, 1) for i, (c, s) in enumerate(((64, 8), (128, 4), (256, 2)))]
    with torch.no_grad():
        y, raw = det([f.clone() for f in feats])
    for i, f in enumerate(feats):
        out[f"detect.x{i}"] = f.numpy()
        out[f"detect.raw{i}"] = raw[i].numpy()
    out["detect.y"] = y.numpy()
    print("module detect", tuple(y.shape))
    np.savez_compressed(GOLD / "modules.npz", **out)


def nms_cases():
    g = np.random.default_rng(7)
    cases = []

    def rand_boxes(n, span=640.0, wh=(8, 256)):
        xy = g.uniform(0, span, (n, 2))
        s = g.uniform(wh[0], wh[1], (n, 2))
        return np.concatenate([xy - s / 2, xy + s / 2], 1).astype(np.float32)

    for n in (1, 2, 7, 64, 200, 400, 1000):
        for thr in (0.45, 0.5, 0.7):
            b = rand_boxes(n)
            # near-duplicate clusters
            k = max(1, n // 4)
            src = g.integers(0, n, k)
            b[:k] = b[src] + g.normal(0, 2.0, (k, 4)).astype(np.float32)
            s = g.uniform(0, 1, n).astype(np.float32)
            if n > 4:  # exact score ties
                s[g.integers(0, n, n // 3)] = s[0]
            cases.append((b, s, thr))
    # class-offset style boxes (fp32 cls * 7680)
    b = rand_boxes(300)
    cls = g.integers(0, 80, (300, 1)).astype(np.float32)
    cases.append(((b + cls * np.float32(7680)).astype(np.float32), g.uniform(0, 1, 300).astype(np.float32), 0.7))
    # known-answer probes
    a_ = [0, 0, 10, 10]
    cases.append((np.array([a_, [20, 20, 30, 30], a_], np.float32), np.array([0.5, 0.5, 0.5], np.float32), 0.5))
    cases.append((np.array([[0, 0, 10, 10], [0, 0, 10, 5]], np.float32), np.array([0.9, 0.8], np.float32), 0.5))  # IoU == thr
    cases.append((np.array([[0, 0, 3, 1], [0, 0, 1, 1]], np.float32), np.array([0.9, 0.8], np.float32), 1.0 / 3.0))
    cases.append((np.array([[5, 5, 5, 5], [5, 5, 5, 5]], np.float32), np.array([0.9, 0.8], np.float32), 0.5))  # NaN IoU
    cases.append((np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5))
    return cases


def gen_nms_torchvision():
    import torchvision

    out = {"n_cases": np.array(0)}
    cases = nms_cases()
    for i, (b, s, thr) in enumerate(cases):
        keep = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
        out[f"c{i}.boxes"], out[f"c{i}.scores"], out[f"c{i}.thr"], out[f"c{i}.keep"] = b, s, np.array(thr), keep
    out["n_cases"] = np.array(len(cases))
    out["torchvision_version"] = np.array(torchvision.__version__)
    np.savez_compressed(GOLD / "nms_torchvision.npz", **out)
    print("nms_torchvision", len(cases), "cases")


def synth_pred(B, A, nc, seed, mean, std):
    """(B, 4+nc, A) fp32: cx,cy~U(0,640), w,h~U(8,256), scores = sigmoid(N(mean, std))."""
    g = np.random.default_rng(seed)
    p = np.empty((B, 4 + nc, A), np.float32)
    p[:, 0:2] = g.uniform(0, 640, (B, 2, A))
    p[:, 2:4] = g.uniform(8, 256, (B, 2, A))
    p[:, 4:] = 1.0 / (1.0 + np.exp(-g.normal(mean, std, (B, nc, A))))
    # clusters of near-duplicate boxes so that suppression actually happens
    k = A // 3
    src = g.integers(0, A, k)
    p[:, 0:4, :k] = p[:, 0:4, src] + g.normal(0, 1.5, (B, 4, k)).astype(np.float32)
    return p


def gen_nms_reference():
    import_reference()
    from yololite.utils import ops

    out = {}
    specs = [
        dict(tag="single_conf25", mean=-3.0, std=2.0, kw=dict(conf_thres=0.25, iou_thres=0.7)),
        dict(tag="single_conf001", mean=-5.0, std=2.0, kw=dict(conf_thres=0.001, iou_thres=0.7)),
        dict(tag="multi_conf001", mean=-7.0, std=2.0, kw=dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),
        dict(tag="agnostic", mean=-3.0, std=2.0, kw=dict(conf_thres=0.25, iou_thres=0.45, agnostic=True)),
        dict(tag="classes", mean=-3.0, std=2.0, kw=dict(conf_thres=0.1, iou_thres=0.6, classes=[0, 3, 17, 79])),
        dict(tag="maxdet", mean=-2.0, std=2.0, kw=dict(conf_thres=0.05, iou_thres=0.9, max_det=20)),
    ]
    for i, sp in enumerate(specs):
        pred = synth_pred(2, 420, 80, 50 + i, sp["mean"], sp["std"])
        res = ops.non_max_suppression(torch.from_numpy(pred.copy()), max_time_img=1e9, **sp["kw"])
        out[f"{sp['tag']}.pred"] = pred
        out[f"{sp['tag']}.counts"] = np.array([len(r) for r in res])
        out[f"{sp['tag']}.dets"] = np.concatenate([r.numpy() for r in res], 0) if res else np.zeros((0, 6), np.float32)
        out[f"{sp['tag']}.kw"] = np.array(repr(sp["kw"]))
        print("nms_reference", sp["tag"], out[f"{sp['tag']}.counts"])
    np.savez_compressed(GOLD / "nms_reference.npz", **out)


def val_case(seed, B=6, max_det=40, nc=5):
    """Seeded synthetic validation batch: labels + detections (jittered labels and random boxes), conf-sorted.
    Coordinates are drawn on a 1/64 grid so IoUs are exact in fp32 but distinct (no ties)."""
    g = np.random.default_rng(seed)
    dets = np.zeros((B, max_det, 6), np.float32)
    counts = np.zeros((B,), np.int32)
    gtb, gtc, offs = [], [], [0]
    for b in range(B):
        L = int(g.integers(0, 12)) if b else 0            # image 0 has no labels
        xy = g.integers(0, 400 * 64, (L, 2)) / 64.0
        wh = g.integers(20 * 64, 200 * 64, (L, 2)) / 64.0
        lab = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        cl = g.integers(0, nc, (L,)).astype(np.float32)
        n = int(g.integers(0, max_det + 1)) if b != 1 else 0   # image 1 has no detections
        rows = []
        for i in range(n):
            if L and g.random() < 0.7:
                j = int(g.integers(0, L))
                jit = g.integers(-30 * 64, 30 * 64, (4,)) / 64.0
                box = lab[j] + jit.astype(np.float32)
                c = cl[j] if g.random() < 0.85 else float(g.integers(0, nc))
            else:
                p0 = g.integers(0, 400 * 64, (2,)) / 64.0
                box = np.concatenate([p0, p0 + g.integers(20 * 64, 200 * 64, (2,)) / 64.0]).astype(np.float32)
                c = float(g.integers(0, nc))
            rows.append([*box, 0.0, c])
        rows = np.asarray(rows, np.float32).reshape(-1, 6)
        rows[:, 4] = np.sort(g.random(n).astype(np.float32))[::-1]
        dets[b, :n] = rows
        counts[b] = n
        gtb.append(lab)
        gtc.append(cl)
        offs.append(offs[-1] + L)
    return dets, counts, np.concatenate(gtb).reshape(-1, 4), np.concatenate(gtc), np.asarray(offs, np.int32)


def gen_val_metrics():
    import_reference()
    from yololite.engine.validator import DetectionValidator  # noqa: F401  (the reference's class)
    from yololite.utils.metrics import ap_per_class, box_iou

    iouv = torch.linspace(0.5, 0.95, 10)
    fake = type("V", (), {"iouv": iouv})()                # match_predictions only reads self.iouv
    out = {"seeds": np.asarray([11, 12, 13])}
    for seed in (11, 12, 13):
        dets, counts, gtb, gtc, offs = val_case(seed)
        tps = np.zeros((dets.shape[0], dets.shape[1], 10), bool)
        for b in range(dets.shape[0]):
            n, s, e = counts[b], offs[b], offs[b + 1]
            if n == 0 or e == s:
                continue
            d = torch.from_numpy(dets[b, :n])
            iou = box_iou(torch.from_numpy(gtb[s:e]), d[:, :4])
            tps[b, :n] = DetectionValidator.match_predictions(fake, d[:, 5], torch.from_numpy(gtc[s:e]), iou).numpy()
        out[f"tp_{seed}"] = tps
        # ap_per_class on the flattened stats of this case
        sel = [(b, i) for b in range(dets.shape[0]) for i in range(counts[b])]
        tp = np.asarray([tps[b, i] for b, i in sel]).reshape(-1, 10)
        conf = np.asarray([dets[b, i, 4] for b, i in sel])
        pc = np.asarray([dets[b, i, 5] for b, i in sel])
        r = ap_per_class(tp, conf, pc, gtc)
        for k, name in zip(range(7), ("tpn", "fpn", "p", "r", "f1", "ap", "cls")):
            out[f"{name}_{seed}"] = np.asarray(r[k])
    np.savez_compressed(GOLD / "val_metrics.npz", **out)
    print("val_metrics.npz")


def gen_checkpoint():
    import_reference()
    from yololite.nn.tasks import DetectionModel, yaml_model_load

    d = yaml_model_load("/root/reference/yololite/cfg/yolo11n.yaml")
    d["scales"] = {"n": [0.5, 0.125, 1024]}              # narrow: 0.8 M parameters keep the fixture small
    d["nc"] = 16
    m = DetectionModel(d, verbose=False).eval()
    fill_state_dict_(m)
    ckpt = {"epoch": -1, "best_fitness": None, "model": m.half(), "ema": None, "updates": None, "optimizer": None,
            "train_args": {"imgsz": 64, "task": "detect"}, "date": "golden", "version": "reference"}
    torch.save(ckpt, GOLD / "ckpt_tiny.pt")
    ref = torch.load(GOLD / "ckpt_tiny.pt", map_location="cpu", weights_only=False)["model"].float().eval()
    x = seeded((2, 3, 64, 64), 321)
    with torch.no_grad():
        y, raw = ref(x)
    np.savez_compressed(GOLD / "ckpt_tiny_out.npz", y=y.numpy(), **{f"raw{i}": r.numpy() for i, r in enumerate(raw)},
                        n_params=np.array(sum(p.numel() for p in ref.parameters())))
    print("ckpt_tiny.pt", (GOLD / "ckpt_tiny.pt").stat().st_size, "bytes")


LETTERBOX_CASES = [
    # (src h, src w, new_shape, auto, scaleup, seed)
    (97, 131, (64, 64), False, True, 1),      # downscale, wide
    (131, 97, (64, 96), False, True, 2),      # downscale, tall, rectangular target
    (23, 31, (64, 64), False, True, 3),       # upscale
    (23, 31, (64, 64), False, False, 4),      # scaleup=False: no resize, centre pad
    (64, 64, (64, 64), False, True, 5),       # identity
    (120, 200, (128, 128), True, True, 6),    # auto=True: minimal-rectangle padding (mod stride)
    (200, 120, (128, 128), True, True, 7),
    (128, 256, (64, 64), False, True, 8),     # exact 2x / 4x downscale
    (1, 50, (32, 32), False, True, 9),        # degenerate one-row image
]


def letterbox_image(h, w, seed):
    return np.random.default_rng(1000 + seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def gen_letterbox():
    import_reference()
    from yololite.data.augment import LetterBox

    out = {"n_cases": np.array(len(LETTERBOX_CASES))}
    for i, (h, w, new_shape, auto, scaleup, seed) in enumerate(LETTERBOX_CASES):
        img = letterbox_image(h, w, seed)
        lb = LetterBox(new_shape, auto=auto, scaleup=scaleup, stride=32)(image=img)
        # predictor.preprocess (engine/predictor.py:76-84): stack, BGR->RGB, BHWC->BCHW, float, /255
        im = np.stack([lb])
        im = im[..., ::-1].transpose((0, 3, 1, 2))
        im = np.ascontiguousarray(im)
        t = torch.from_numpy(im).float()
        t /= 255
        out[f"lb_{i}"] = lb
        out[f"pre_{i}"] = t.numpy()
    np.savez_compressed(GOLD / "letterbox.npz", **out)
    print("letterbox.npz", len(LETTERBOX_CASES), "cases")


REAL_TOPK = 384     # rows of y (anchors with the highest class score) kept per image


def _capture_nms(ops_mod, store):
    """Wrap the reference's ops.non_max_suppression (harness-side wrapper, the function itself is untouched): record the
    pre-NMS tensor and the output, and lift the wall-clock bail-out (ops.py:207,269-271) that would truncate results."""
    orig = ops_mod.non_max_suppression

    def wrapped(prediction, *a, **kw):
        p = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
        store["y"] = p.detach().clone()
        kw["max_time_img"] = 1e9
        out = orig(prediction, *a, **kw)
        store["nms"] = [o.detach().clone() for o in out]
        return out

    ops_mod.non_max_suppression = wrapped
    return orig


def _top_rows(y, k=REAL_TOPK):
    """(B, 4+nc, A) -> indices (B, k) of the anchors with the highest max class score and their (B, 4+nc, k) columns."""
    sc = y[:, 4:].amax(1)
    idx = sc.argsort(dim=1, descending=True, stable=True)[:, :k]
    return idx.numpy().astype(np.int32), torch.gather(y, 2, idx[:, None, :].expand(-1, y.shape[1], -1)).numpy()


def gen_real_images():
    import shutil
    import tempfile
    import zlib

    import_reference()
    import yaml
    from yololite.data import dataset as ref_dataset
    from yololite.engine import validator as ref_validator
    from yololite.utils import ops as ref_ops

    tmp = Path(tempfile.mkdtemp(prefix="ylref_real_"))
    m = build_reference_model("n")
    fill_state_dict_(m)
    # the str-weights route is the one that works in the reference; it pickles / loads fp16 weights
    torch.save({"model": m.half(), "train_args": {}, "ema": None}, tmp / "seeded.pt")
    out = {"weights": np.array("oracle/weights.py fill_state_dict_(yolo11n), rounded to fp16 and back")}
    cap = {}
    orig_nms = _capture_nms(ref_ops, cap)

    # ---- config 1: main.py:15 — YOLOLite(weights)(["boats.jpg"]) -> predictor, auto letterbox 384x640
    boats = Path(REF) / "boats.jpg"
    out["boats.jpg"] = np.frombuffer(boats.read_bytes(), dtype=np.uint8)
    # YOLOLite.predict hands the in-memory module to AutoBackend, which crashes on the deleted `fuse`; the
    # predictor it would construct (engine/model.py:95-99: conf .25, batch 1, mode predict) is built here with the
    # weights as a path string, the route that works
    from yololite.engine.predictor import DetectionPredictor

    predictor = DetectionPredictor(overrides=dict(conf=0.25, batch=1, mode="predict", save=False, verbose=False,
                                                  device="cpu", iou=0.7, project=str(tmp / "runs")))
    res = predictor(source=[str(boats)], model=str(tmp / "seeded.pt"))
    r = res[0]
    y = cap["y"].float()
    idx, cols = _top_rows(y)
    out["boats.y_shape"] = np.array(y.shape)
    out["boats.top_idx"], out["boats.top_cols"] = idx, cols
    out["boats.nms"] = cap["nms"][0].numpy()            # letterboxed-image space (before scale_boxes)
    out["boats.boxes"] = r.boxes.data.numpy()           # original-image space (after scale_boxes + clip)
    out["boats.orig_shape"] = np.array(r.orig_shape)
    print("boats", tuple(y.shape), "dets", len(r.boxes.data), "top score", float(y[:, 4:].max()))
    # the predictor's preprocessed batch, recomputed with the reference's own LetterBox (what preprocess() ran)
    import cv2
    from yololite.data.augment import LetterBox

    im0 = cv2.imread(str(boats))
    lb = LetterBox((640, 640), auto=True, stride=32)(image=im0)
    im = np.ascontiguousarray(np.stack([lb])[..., ::-1].transpose((0, 3, 1, 2)))
    assert tuple(im.shape[2:]) == tuple(y.shape and (384, 640)), im.shape
    out["boats.im_crc"] = np.array(zlib.crc32(im.tobytes()), dtype=np.int64)
    out["boats.im_shape"] = np.array(im.shape)

    # ---- config 2: YOLOLite.val on a scratch copy of coco8 (rect batch of 4, multi-label NMS, conf 0.001)
    ds = tmp / "coco8"
    shutil.copytree(Path(REF) / "coco8", ds, ignore=shutil.ignore_patterns("*.cache"))
    cfg = yaml.safe_load((Path(REF) / "coco8" / "coco8.yaml").read_text())
    cfg["path"] = str(ds)
    (tmp / "coco8_abs.yaml").write_text(yaml.safe_dump(cfg))
    # harness shim: cache_labels never writes "version" but get_labels pops it
    orig_cache = ref_dataset.YOLODataset.cache_labels

    def cache_labels(self, *a, **kw):
        x = orig_cache(self, *a, **kw)
        x["version"] = "harness"
        return x

    ref_dataset.YOLODataset.cache_labels = cache_labels
    batches, stats_cap = [], {}
    orig_update = ref_validator.DetectionValidator.update_metrics

    def update_metrics(self, preds, batch):
        batches.append({k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})
        stats_cap["predn"] = [self._prepare_pred(p, self._prepare_batch(si, batch)) for si, p in enumerate(preds)]
        return orig_update(self, preds, batch)

    ref_validator.DetectionValidator.update_metrics = update_metrics
    orig_get_stats = ref_validator.DetectionValidator.get_stats

    def get_stats(self):
        stats_cap["stats"] = {k: torch.cat(v, 0).cpu().numpy() for k, v in self.stats.items()}
        return orig_get_stats(self)

    ref_validator.DetectionValidator.get_stats = get_stats
    # engine/model.py:101-107: YOLOLite.val builds DetectionValidator(args={rect: True, mode: val, ...})(model=...)
    validator = ref_validator.DetectionValidator(args=dict(rect=True, mode="val", data=str(tmp / "coco8_abs.yaml"),
                                                           device="cpu", batch=16, workers=0, plots=False, verbose=False,
                                                           project=str(tmp / "runs"), save_json=False))
    validator(model=str(tmp / "seeded.pt"))
    metrics = validator.metrics
    assert len(batches) == 1
    b = batches[0]
    img = b["img"]                                      # validator.preprocess output: float /255 on the device
    img_u8 = (img * 255).round().to(torch.uint8)
    assert torch.equal(img_u8.float() / 255, img)
    y = cap["y"].float()
    idx, cols = _top_rows(y)
    out["coco8.y_shape"] = np.array(y.shape)
    out["coco8.top_idx"], out["coco8.top_cols"] = idx, cols
    out["coco8.img_shape"] = np.array(img_u8.shape)
    out["coco8.img_crc"] = np.array(zlib.crc32(img_u8.numpy().tobytes()), dtype=np.int64)
    names = [Path(f).name for f in b["im_file"]]
    out["coco8.files"] = np.array(names)
    for i, f in enumerate(b["im_file"]):
        out[f"coco8.jpg{i}"] = np.frombuffer(Path(f).read_bytes(), dtype=np.uint8)
        lab = Path(f.replace("/images/", "/labels/", 1)).with_suffix(".txt")
        out[f"coco8.txt{i}"] = np.array([r.split() for r in lab.read_text().strip().splitlines()], dtype=np.float32)
    out["coco8.cls"] = b["cls"].numpy()
    out["coco8.bboxes"] = b["bboxes"].numpy()
    out["coco8.batch_idx"] = b["batch_idx"].numpy()
    out["coco8.ori_shape"] = np.array(b["ori_shape"])
    out["coco8.ratio"] = np.array([rp[0] for rp in b["ratio_pad"]], dtype=np.float64)
    out["coco8.pad"] = np.array([rp[1] for rp in b["ratio_pad"]], dtype=np.float64)
    nms = cap["nms"]
    out["coco8.nms_counts"] = np.array([len(o) for o in nms])
    out["coco8.nms"] = torch.cat(nms, 0).numpy()
    out["coco8.predn"] = torch.cat(stats_cap["predn"], 0).numpy()
    st = stats_cap["stats"]
    out["coco8.tp"], out["coco8.conf"], out["coco8.pred_cls"] = st["tp"], st["conf"], st["pred_cls"]
    out["coco8.target_cls"] = st["target_cls"]
    rd = metrics.results_dict
    out["coco8.metric_keys"] = np.array(list(rd.keys()))
    out["coco8.metric_vals"] = np.array([float(v) for v in rd.values()], dtype=np.float64)
    print("coco8", tuple(img_u8.shape), "nms counts", out["coco8.nms_counts"], "tp any", bool(st["tp"].any()), rd)
    ref_ops.non_max_suppression = orig_nms
    ref_dataset.YOLODataset.cache_labels = orig_cache
    ref_validator.DetectionValidator.update_metrics = orig_update
    ref_validator.DetectionValidator.get_stats = orig_get_stats
    np.savez_compressed(GOLD / "real_images.npz", **out)
    shutil.rmtree(tmp, ignore_errors=True)
    print("real_images.npz", (GOLD / "real_images.npz").stat().st_size, "bytes")


def gen_scale_boxes():
    import_reference()
    from yololite.utils import ops

    g = np.random.default_rng(77)
    cases = [  # (img1 (letterboxed) h,w ; img0 (original) h,w ; explicit ratio_pad or None)
        ((384, 640), (1080, 1920), None), ((640, 640), (480, 640), None), ((640, 640), (427, 640), None),
        ((672, 672), (428, 640), None), ((640, 480), (1080, 810), None), ((64, 64), (97, 131), None),
        ((64, 96), (131, 97), None), ((640, 640), (333, 500), None), ((640, 640), (500, 333), None),
        ((672, 672), (480, 640), ((1.05, 1.05), (0.0, 84.0))), ((672, 672), (640, 427), ((1.05, 1.05), (112.0, 0.0))),
        ((640, 640), (1281, 1920), None),
    ]
    out = {"n_cases": np.array(len(cases))}
    for i, (s1, s0, rp) in enumerate(cases):
        n = 40
        xy = g.uniform(-20, max(s1) + 20, (n, 2))
        wh = g.uniform(1, 300, (n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        got = ops.scale_boxes(s1, torch.from_numpy(boxes.copy()), s0, ratio_pad=rp).numpy()
        out[f"c{i}.img1"], out[f"c{i}.img0"] = np.array(s1), np.array(s0)
        out[f"c{i}.ratio_pad"] = np.array([rp[0][0], rp[1][0], rp[1][1]] if rp else [0.0, 0.0, 0.0], dtype=np.float64)
        out[f"c{i}.has_rp"] = np.array(rp is not None)
        out[f"c{i}.boxes"], out[f"c{i}.out"] = boxes, got
    np.savez_compressed(GOLD / "scale_boxes.npz", **out)
    print("scale_boxes.npz", len(cases), "cases")


GENERATORS = {
    "val": lambda: gen_val_metrics(), "ckpt": lambda: gen_checkpoint(), "letterbox": lambda: gen_letterbox(),
    "nms_torchvision": lambda: gen_nms_torchvision(), "nms_reference": lambda: gen_nms_reference(),
    "modules": lambda: gen_modules(), "models": lambda: gen_models(), "real": lambda: gen_real_images(),
    "scale_boxes": lambda: gen_scale_boxes(),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Regenerate tests/golden/ from the unmodified reference in /root/reference")
    ap.add_argument("--write", action="store_true", help="actually (over)write tests/golden/; without it nothing is touched")
    ap.add_argument("--only", nargs="*", choices=sorted(GENERATORS), help="fixture groups to regenerate (default: all)")
    args = ap.parse_args()
    if not args.write:
        ap.error("refusing to overwrite tests/golden/ without --write")
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    for name in (args.only or list(GENERATORS)):
        GENERATORS[name]()
