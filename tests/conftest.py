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
