"""Per-layer timeline of the conv chains of one yolo11n plan (yl_conv_chain_debug): where a chain layer's time goes.

    python tools/chain_timeline.py [batch]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from bench import randomise_model_  # noqa: E402
from yololite import _C  # noqa: E402
from yololite.nn.tasks import DetectionModel  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
m.use_cuda_graph = False
x = torch.rand(batch, 3, 640, 640, device="cuda")
for _ in range(3):
    m.infer(x)
torch.cuda.synchronize()
plan = m._get_plan(x.shape, x.device)[0]
lib = _C.load()
s = _C.stream_ptr()
for i, ((fn, args, _), md) in enumerate(zip(plan.calls, plan.meta)):
    if md["kind"] != "conv_chain":
        continue
    n = len(md["members"])
    buf = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    lib.yl_conv_chain_debug(buf.data_ptr())
    for _ in range(2):                       # second pass: warm L2 / descriptors
        _C.check(fn(*args, s), "chain")
    torch.cuda.synchronize()
    lib.yl_conv_chain_debug(None)
    t = buf.cpu().numpy().astype("float64")
    t0 = t[0, 0]
    print(f"chain at launch {i}: {n} layers, {(t[-1, 5] - t0) / 1e3:.1f} us (CTA 0)")
    print(f"{'layer':<34}{'total':>7}{'->operands':>11}{'->mma done':>11}{'->epi out':>10}{'->barrier':>10}   (us; epilogue entered at)")
    for L, mm in enumerate(md["members"]):
        a = t[L]
        print(f"{mm['desc']:<34}{(a[5] - a[0]) / 1e3:>7.2f}{(a[1] - a[0]) / 1e3:>11.2f}{(a[2] - a[0]) / 1e3:>11.2f}"
              f"{(a[4] - a[0]) / 1e3:>10.2f}{(a[5] - a[0]) / 1e3:>10.2f}   {(a[3] - a[0]) / 1e3:.2f}")
