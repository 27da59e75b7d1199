"""Checkpoint ingest: a `.pt` pickled BY THE REFERENCE (tests/golden/ckpt_tiny.pt, written by
oracle/gen_golden.py::gen_checkpoint from the unmodified reference classes) must unpickle into this package's classes
(identical module paths), keep every state-dict key, and — on the GPU — reproduce the reference's fp32 output."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
CKPT = ROOT / "tests" / "golden" / "ckpt_tiny.pt"
OUT = np.load(ROOT / "tests" / "golden" / "ckpt_tiny_out.npz")


def test_reference_checkpoint_unpickles_into_our_classes():
    import yololite
    from yololite.nn.modules import C2PSA, C3k2, SPPF, Conv, Detect
    from yololite.nn.tasks import DetectionModel, attempt_load_one_weight

    assert str(ROOT / "yolo-lite_b200") in yololite.__file__
    model, ckpt = attempt_load_one_weight(str(CKPT))
    assert type(model) is DetectionModel and not model.training
    assert ckpt["train_args"]["imgsz"] == 64 and model.args["imgsz"] == 64
    kinds = [type(m) for m in model.model]
    assert kinds[0] is Conv and kinds[2] is C3k2 and kinds[9] is SPPF and kinds[10] is C2PSA and kinds[-1] is Detect
    assert sum(p.numel() for p in model.parameters()) == int(OUT["n_params"])
    assert all(p.dtype == torch.float32 for p in model.parameters())       # `.float()` after the fp16 checkpoint
    det = model.model[-1]
    assert det.nc == 16 and det.no == 16 + 64 and det.stride.tolist() == [8.0, 16.0, 32.0]
    # same keys as a freshly built model of the same yaml
    fresh = DetectionModel(model.yaml, verbose=False)
    assert set(fresh.state_dict()) == set(model.state_dict())


def test_yololite_front_end_loads_pt():
    from yololite import YOLOLite

    yl = YOLOLite(str(CKPT))
    assert yl.task == "detect" and yl.ckpt_path == str(CKPT)
    assert yl.overrides["model"] == str(CKPT)


@pytest.mark.gpu
def test_reference_checkpoint_reproduces_reference_output():
    from yololite.nn.tasks import attempt_load_one_weight

    model, _ = attempt_load_one_weight(str(CKPT), device="cuda:0")
    x = torch.from_numpy(np.random.default_rng(321).uniform(0, 1, (2, 3, 64, 64)).astype(np.float32)).cuda()
    y, raws = model(x)
    y = y.float().cpu().numpy()
    ref = OUT["y"]
    assert y.shape == ref.shape
    box_err = np.abs(y[:, :4] - ref[:, :4]).max()
    cls_err = np.abs(y[:, 4:] - ref[:, 4:]).max()
    assert box_err <= 0.5 and cls_err <= 1e-2, (box_err, cls_err)          # north-star head tolerances
    for i, r in enumerate(raws):
        want = OUT[f"raw{i}"]
        d = np.abs(r.float().cpu().numpy() - want)
        assert (d <= 0.15 + 3e-2 * np.abs(want)).all(), (i, d.max())
