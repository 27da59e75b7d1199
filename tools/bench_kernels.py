"""Micro-benchmarks of single kernels through the C-ABI on the GPU box (CUDA events, L2-cold rotation).

    python tools/bench_kernels.py dwconv      # Detect-head depthwise shapes x strip widths
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from yololite import _C, _ops  # noqa: E402

_C.init(0)


def timeit(fn, reps=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def dwconv():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (n, c, h, w) in [(64, 64, 80, 80), (64, 80, 80, 80), (64, 128, 40, 40), (64, 256, 20, 20), (64, 128, 20, 20)]:
        wt = torch.randn(c, 1, 3, 3) * 0.3
        pc = _ops.pack_conv(wt, bn=(torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c), 1e-3))
        x = _ops.new_buffer(n, h, w, c)
        x.buf.normal_()
        y = _ops.new_buffer(n, h, w, c)
        gb = n * h * w * c * 2 * 2 / 1e9
        row = f"dwconv {c:>3}ch {h}x{w} bs{n}: "
        for strip in (1, 2, 4):
            os.environ["YL_DW_STRIP"] = str(strip)
            us = timeit(lambda: _ops.conv(x, y, pc, 1, True), flush=flush)
            row += f" P={strip}: {us:7.1f} us {gb / us * 1e6:6.0f} GB/s |"
        print(row)


if __name__ == "__main__":
    {"dwconv": dwconv}[sys.argv[1] if len(sys.argv) > 1 else "dwconv"]()
