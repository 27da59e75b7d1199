"""Device selection / inference-mode helpers (reference yololite/utils/torch_utils.py:37-47, 92-172)."""
from __future__ import annotations

import torch


def select_device(device=None, batch=0, verbose=False) -> torch.device:
    """'' / None / 'cuda' / 'cuda:1' / 0 / '0' -> torch.device.  'cpu' is refused: there is no CPU path."""
    if isinstance(device, torch.device):
        dev = device
    else:
        s = str(device if device is not None else "").lower().replace("cuda:", "").replace("cuda", "").strip()
        if s == "cpu":
            raise RuntimeError("device='cpu' requested, but yololite runs on sm_100 GPUs only (no CPU fallback)")
        dev = torch.device("cuda", int(s.split(",")[0]) if s else (torch.cuda.current_device() if torch.cuda.is_available() else 0))
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available: yololite runs on sm_100 GPUs only (no CPU fallback)")
    return dev


def smart_inference_mode():
    def decorate(fn):
        return torch.inference_mode()(fn)

    return decorate
