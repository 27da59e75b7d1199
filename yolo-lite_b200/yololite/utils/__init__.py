"""Small runtime utilities (reference yololite/utils/__init__.py is 1000 lines of env probing, settings files and
auto-install logic, none of which is on the inference path and none of which is reproduced)."""
from __future__ import annotations

import logging
from pathlib import Path

import yaml

LOGGER = logging.getLogger("yololite")
if not LOGGER.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(message)s"))
    LOGGER.addHandler(_h)
    LOGGER.setLevel(logging.INFO)
    LOGGER.propagate = False

ROOT = Path(__file__).resolve().parents[1]


def yaml_load(file) -> dict:
    with open(file, errors="ignore", encoding="utf-8") as f:
        return yaml.safe_load(f) or {}


def colorstr(*a):
    return str(a[-1])
