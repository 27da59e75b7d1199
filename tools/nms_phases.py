"""Where does nms_select_kernel spend its cycles on the bench workload?  python tools/nms_phases.py [batch]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from bench import randomise_model_  # noqa: E402
from yololite import _C, _ops  # noqa: E402
from yololite.nn.tasks import DetectionModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _C.init(0)
torch.manual_seed(0)
m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
x = torch.rand(B, 3, 640, 640, device="cuda")
y, _ = m.infer(x)
y = y.clone()
buf = torch.zeros((B, 8), dtype=torch.int64, device="cuda")
for _ in range(3):
    _ops.nms_batched(y, 0.25, 0.7)
lib.yl_debug_nms_phases(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_ops.nms_batched(y, 0.25, 0.7)
e1.record()
torch.cuda.synchronize()
lib.yl_debug_nms_phases(None)
t = buf.cpu().numpy().astype(float)
names = ["setup", "stage+sort", "vs kept + block matrix", "(unused)", "(unused)", "greedy+append"]
tot = t[:, :6].sum(1)
print(f"B={B}: filter+select {e0.elapsed_time(e1) * 1e3:.1f} us; per-image select cycles: mean {tot.mean():.0f} max {tot.max():.0f}"
      f" ({tot.max() / 1.965e3:.1f} us @1.965 GHz)")
for k, nm in enumerate(names):
    print(f"  {nm:<12} {t[:, k].mean():>10.0f} cyc  {100 * t[:, k].mean() / tot.mean():5.1f} %")
print("  kept", t[:, 6].mean(), "candidates consumed", t[:, 7].mean())
