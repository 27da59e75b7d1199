"""pytest config: registers the `gpu` marker and puts the product package (import name `yololite`, living in
yolo-lite_b200/) and the repo root (for `oracle`) on sys.path."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "yolo-lite_b200", ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _gpu_ready() -> str | None:
    """None when the `gpu` tests can run here, else the reason they cannot (no silent CPU fallback exists)."""
    try:
        import torch

        from yololite import _C

        if not _C.lib_path().exists():
            return f"{_C.lib_path()} is not built (python yolo-lite_b200/csrc/build.py)"
        if not torch.cuda.is_available():
            return "no CUDA device"
        if torch.cuda.get_device_capability(0)[0] != 10:
            return f"device is sm_{torch.cuda.get_device_capability(0)[0]}x, libyl11 is sm_100a only"
    except Exception as e:  # pragma: no cover
        return f"{type(e).__name__}: {e}"
    return None


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a box without a B200 skips the gpu-marked tests (with the reason) instead of failing;
    `-m gpu` on the GPU box runs them all."""
    why = _gpu_ready()
    if why is None:
        return
    skip = pytest.mark.skip(reason=f"needs a B200: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(GOLDEN / name, allow_pickle=False)

    return load


def model_state_dict(npz):
    """Rebuild the name-keyed deterministic state_dict whose shapes are recorded in a model_*.npz fixture."""
    from oracle.weights import tensor_for

    sd = {}
    for k, s, nd in zip(npz["keys"], npz["shapes"], npz["ndims"]):
        sd[str(k)] = tensor_for(str(k), tuple(int(v) for v in s[: int(nd)]))
    return sd
