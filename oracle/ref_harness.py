"""oracle/ref_harness.py — imports the UNMODIFIED reference from /root/reference (build container only).

The reference needs three shims to import and run offline: a matplotlib stub, offline env
vars and a scratch config dir.  Nothing here is used at test time on the GPU box: it only serves
oracle/gen_golden.py, which writes tests/golden/.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

REF_ROOT = "/root/reference"


def import_reference():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"{REF_ROOT} not present: the reference can only be imported in the build container")
    cur = sys.modules.get("yololite")
    if cur is not None and str(getattr(cur, "__file__", "")).startswith(REF_ROOT):
        return cur                       # already the reference: keep module identities stable across calls
    os.environ.setdefault("YOLO_OFFLINE", "true")
    os.environ.setdefault("YOLO_AUTOINSTALL", "false")
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="ylref_cfg_"))
    os.environ.setdefault("YOLO_VERBOSE", "false")
    sys.dont_write_bytecode = True
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    # the product package is also called `yololite`; make sure the reference wins in this process
    sys.path[:] = [p for p in sys.path if "yolo-lite_b200" not in p]
    for k in [k for k in sys.modules if k == "yololite" or k.startswith("yololite.")]:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    import yololite  # noqa: F401

    assert yololite.__file__.startswith(REF_ROOT), yololite.__file__
    return yololite


def build_reference_model(scale: str = "n"):
    import_reference()
    from yololite.nn.tasks import DetectionModel, yaml_model_load

    cfg = yaml_model_load(f"{REF_ROOT}/yololite/cfg/yolo11{scale}.yaml")
    return DetectionModel(cfg, verbose=False).eval()
