"""Image preprocess: LetterBox + BGR->RGB + CHW + float/255.

CPU: the oracle (oracle/letterbox_ref.py) against the goldens written by the unmodified reference
(tests/golden/letterbox.npz, oracle/gen_golden.py) and against the installed cv2 on random shapes.
GPU: yl_letterbox_u8 through the C-ABI, bit-exact against goldens and oracle, and through the predictor."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]

from oracle import letterbox_ref  # noqa: E402
from oracle.gen_golden import LETTERBOX_CASES, letterbox_image  # noqa: E402

GOLD = np.load(ROOT / "tests" / "golden" / "letterbox.npz")


@pytest.mark.parametrize("i", range(len(LETTERBOX_CASES)))
def test_oracle_matches_reference_golden(i):
    h, w, new_shape, auto, scaleup, seed = LETTERBOX_CASES[i]
    img = letterbox_image(h, w, seed)
    lb = letterbox_ref.letterbox(img, new_shape, auto=auto, scaleup=scaleup, stride=32)
    assert np.array_equal(lb, GOLD[f"lb_{i}"])
    pre = np.ascontiguousarray(lb[None][..., ::-1].transpose(0, 3, 1, 2)).astype(np.float32) / np.float32(255)
    assert np.array_equal(pre, GOLD[f"pre_{i}"])


def test_oracle_resize_matches_cv2_bit_exact():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for it in range(60):
        sh, sw, dh, dw = (int(v) for v in rng.integers(3, 300, 4))
        if it % 6 == 0:
            dh, dw = max(sh // 2, 1), max(sw // 2, 1)
        if it % 7 == 0:
            dh, dw = sh * 2, sw * 2
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(letterbox_ref.resize_linear_u8(src, (dw, dh)), want), (sh, sw, dh, dw)


def test_host_letterbox_class_matches_oracle():
    from yololite.data import LetterBox

    for h, w, new_shape, auto, scaleup, seed in LETTERBOX_CASES:
        img = letterbox_image(h, w, seed)
        got = LetterBox(new_shape, auto=auto, scaleup=scaleup, stride=32)(image=img)
        assert np.array_equal(got, letterbox_ref.letterbox(img, new_shape, auto=auto, scaleup=scaleup, stride=32))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(LETTERBOX_CASES)))
def test_cuda_letterbox_matches_golden(i):
    from yololite.data import letterbox_batch_cuda

    h, w, new_shape, auto, scaleup, seed = LETTERBOX_CASES[i]
    img = letterbox_image(h, w, seed)
    got = letterbox_batch_cuda([img, img], new_shape, auto=auto, stride=32, scaleup=scaleup)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.shape[0] == 2 and np.array_equal(got[0], got[1])
    assert np.array_equal(got[:1], GOLD[f"pre_{i}"]), np.abs(got[:1] - GOLD[f"pre_{i}"]).max()


@pytest.mark.gpu
def test_cuda_letterbox_random_shapes_vs_oracle():
    from yololite.data import letterbox_batch_cuda

    rng = np.random.default_rng(3)
    for it in range(12):
        h, w = (int(v) for v in rng.integers(8, 700, 2))
        imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(3)]
        target = (int(rng.integers(2, 12)) * 32, int(rng.integers(2, 12)) * 32)
        auto = bool(it % 2)
        want = letterbox_ref.preprocess(imgs, target, auto=auto, stride=32)
        got = letterbox_batch_cuda(imgs, target, auto=auto, stride=32).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), (h, w, target, auto)


@pytest.mark.gpu
def test_cuda_letterbox_full_size_640():
    """BASELINE size: 1080x810 -> 640 canvas; checked against the oracle bit for bit."""
    from yololite.data import letterbox_batch_cuda

    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (1080, 810, 3), dtype=np.uint8) for _ in range(2)]
    want = letterbox_ref.preprocess(imgs, (640, 640), auto=True, stride=32)
    got = letterbox_batch_cuda(imgs, (640, 640), auto=True, stride=32).cpu().numpy()
    assert got.shape == want.shape == (2, 3, 640, 480) and np.array_equal(got, want)


@pytest.mark.gpu
def test_predictor_numpy_source_uses_cuda_letterbox():
    """A list of BGR uint8 images through YOLOLite.predict: same detections as feeding the oracle-preprocessed
    tensor (the preprocess is bit-exact, so the two runs see identical network inputs)."""
    from oracle.weights import fill_state_dict_
    from yololite import YOLOLite

    yl = YOLOLite("yolo11n.yaml")
    fill_state_dict_(yl.model)
    rng = np.random.default_rng(9)
    imgs = [rng.integers(0, 256, (120, 200, 3), dtype=np.uint8) for _ in range(2)]
    r1 = yl.predict(imgs, imgsz=128, conf=0.001, verbose=False, device=0)
    x = torch.from_numpy(letterbox_ref.preprocess(imgs, (128, 128), auto=True, stride=32))
    r2 = yl.predict(x, imgsz=tuple(x.shape[2:]), conf=0.001, verbose=False, device=0)
    assert len(r1) == len(r2) == 2
    for a, b in zip(r1, r2):
        da, db = a.boxes.data.cpu().numpy(), b.boxes.data.cpu().numpy()
        assert da.shape == db.shape
        # identical network input -> identical raw detections; only the box rescale back to the original image
        # differs (r2's "original" is the letterboxed tensor itself)
        assert np.array_equal(da[:, 4:], db[:, 4:])
