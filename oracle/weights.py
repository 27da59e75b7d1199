"""oracle/weights.py — TEST INFRASTRUCTURE.  Deterministic, name-keyed parameters.

Every tensor of a state_dict is regenerated from crc32(name) alone, so the reference (when generating
tests/golden/), the oracle restatement and the CUDA implementation all see identical weights without storing
them.  BatchNorm statistics are randomised (default init gamma=1, beta=0, mean=0, var=1 would leave the BN
fold untested); the Detect class bias is spread so scores cover (0, 1).
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _rng(name: str, salt: int = 0) -> np.random.Generator:
    return np.random.default_rng((zlib.crc32(name.encode()) + 7919 * salt) & 0xFFFFFFFF)


def tensor_for(name: str, shape, salt: int = 0) -> torch.Tensor:
    g = _rng(name, salt)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if ".bn." in name or name.startswith("bn."):
        if leaf == "weight":
            v = g.uniform(0.5, 1.5, shape)
        elif leaf == "running_var":
            v = g.uniform(0.5, 1.5, shape)
        else:  # bias, running_mean
            v = g.normal(0.0, 0.1, shape)
    elif ".dfl." in name or name.startswith("dfl."):
        v = np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape)  # frozen 0..15 (block.py:61-63)
    elif leaf == "bias":
        v = g.normal(-4.0, 1.5, shape) if ".cv3." in name else g.normal(1.0, 0.5, shape)
    else:  # conv weight (co, ci/g, k, k)
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        b = 1.7 / np.sqrt(max(fan_in, 1))
        v = g.uniform(-b, b, shape)
    return torch.from_numpy(np.asarray(v, dtype=np.float32))


def fill_state_dict_(module: torch.nn.Module, salt: int = 0) -> torch.nn.Module:
    """Overwrite every parameter/buffer of `module` in place with its name-keyed deterministic value."""
    sd = module.state_dict()
    new = {k: tensor_for(k, v.shape, salt).to(v.dtype) for k, v in sd.items()}
    module.load_state_dict(new, strict=True)
    return module


def state_dict_like(shapes: dict, salt: int = 0) -> dict:
    return {k: tensor_for(k, s, salt) for k, s in shapes.items()}
