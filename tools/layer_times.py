"""Per-launch table of one model plan on the GPU box: kind, shape, device time, achieved GB/s and TFLOP/s.

    python tools/layer_times.py [n|s|m] [batch]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from bench import randomise_model_  # noqa: E402
from yololite.nn.tasks import DetectionModel  # noqa: E402

scale = sys.argv[1] if len(sys.argv) > 1 else "n"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = randomise_model_(DetectionModel(f"yolo11{scale}.yaml", verbose=False)).eval().cuda()
x = torch.rand(batch, 3, 640, 640, device="cuda")
for _ in range(3):
    m.infer(x)
torch.cuda.synchronize()
plan = m._get_plan(x.shape, x.device)[0]
lt = plan.time_launches(reps=5)
tot = sum(lt)
print(f"yolo11{scale} bs={batch}: {len(lt)} launches, sum of per-launch times {tot:.3f} ms, activations {plan.bytes / 1e9:.2f} GB")
print(f"{'#':>3} {'kind':<20}{'shape':<34}{'us':>9}{'%':>6}{'GB/s':>9}{'TF/s':>8}")
for i, (md, t) in enumerate(zip(plan.meta, lt)):
    print(f"{i:>3} {md['kind']:<20}{md['desc']:<34}{t * 1e3:>9.1f}{100 * t / tot:>6.1f}"
          f"{md['bytes'] / 1e9 / (t / 1e3):>9.0f}{md['flops'] / 1e12 / (t / 1e3):>8.1f}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    m.infer(x)
e1.record()
torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 20:.3f} ms/forward -> {batch / (e0.elapsed_time(e1) / 20) * 1e3:.0f} img/s (model only)")
