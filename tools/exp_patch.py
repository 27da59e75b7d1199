"""Experiment: which UMMA descriptor variant makes the halo-patch conv mode exact?  (GPU box only)

    python tools/exp_patch.py            # spawns one subprocess per variant
    python tools/exp_patch.py <pitch> <bo> <k64>
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]

SHAPES = [  # ci, co, n, h, w
    (16, 8, 1, 24, 24), (8, 16, 1, 24, 24), (16, 32, 2, 48, 40), (32, 16, 1, 32, 32), (32, 32, 2, 40, 40),
    (64, 32, 1, 40, 40), (32, 64, 1, 40, 40), (64, 64, 1, 32, 16),
]


def run_variant(pitch, bo, k64):
    import torch
    import torch.nn.functional as F

    from yololite import _C, _ops

    os.environ.update(YL_PATCH="1", YL_PATCH_PITCH=str(pitch), YL_PATCH_BO=str(bo), YL_PATCH_K64=str(k64))
    _C.init(0)
    for ci, co, n, h, w in SHAPES:
        g = torch.Generator().manual_seed(ci * 1000 + co)
        x = torch.randn(n, ci, h, w, generator=g)
        wt = torch.randn(co, ci, 3, 3, generator=g) * (1.5 / (ci * 9) ** 0.5)
        bias = torch.randn(co, generator=g) * 0.1
        pc = _ops.pack_conv(wt, None, bias)
        ref = F.conv2d(x.bfloat16().float(), wt.bfloat16().float(), bias, 1, 1)
        xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
        yb = torch.zeros((n, h, w, co), dtype=torch.bfloat16, device="cuda")
        _ops.conv(_ops.View(xb, 0, ci), _ops.View(yb, 0, co), pc, 1, False, None, False, impl=_C.IMPL_TCGEN05)
        torch.cuda.synchronize()
        got = yb.float().cpu().permute(0, 3, 1, 2)
        err = (got - ref).abs().max().item()
        print(f"pitch={pitch} bo={bo} k64={k64}  {ci:>3}->{co:<3} {h}x{w} n={n}: max|err|={err:.4f} "
              f"{'OK' if err < 3e-2 else 'BAD'}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 4:
        run_variant(*map(int, sys.argv[1:]))
    else:
        for pitch in (10, 16):
            for bo in (1, 0):
                for k64 in (0, 1):
                    r = subprocess.run([sys.executable, __file__, str(pitch), str(bo), str(k64)], capture_output=True,
                                       text=True, timeout=300)
                    print(r.stdout, end="")
                    if r.returncode != 0:
                        print(f"pitch={pitch} bo={bo} k64={k64}: rc={r.returncode} {r.stderr[-300:]}")
