"""Where does the end-to-end (host tensor -> Results) time go?   python tools/e2e_probe.py [batch]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from bench import randomise_model_  # noqa: E402
from yololite import YOLOLite  # noqa: E402
from yololite.engine.predictor import DetectionPredictor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
host = [torch.rand(B, 3, 640, 640).pin_memory() for _ in range(3)]
d = torch.empty_like(host[0], device=dev)
torch.cuda.synchronize()
for _ in range(2):
    d.copy_(host[0], non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    d.copy_(host[i % 3], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D pinned {host[0].numel() * 4 / 1e6:.0f} MB: {dt * 1e3:.2f} ms -> {host[0].numel() * 4 / dt / 1e9:.1f} GB/s "
      f"-> PCIe ceiling {B / dt:.0f} img/s")

yl = YOLOLite("yolo11n.yaml")
randomise_model_(yl.model)
yl.model.to(dev)
kw = dict(conf=0.25, iou=0.7, max_det=300, verbose=False, device=dev, batch=B)


def run(x, n=12):
    for i in range(3):
        yl.predict(x[i % len(x)], **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prev = None
    for i in range(n):
        r = yl.predict(x[i % len(x)], **kw)
        if prev is not None and LAG:
            len(prev[0].boxes.data.cpu())
        elif not LAG:
            len(r[0].boxes.data.cpu())
        prev = r
    if LAG:
        len(prev[0].boxes.data.cpu())
    torch.cuda.synchronize()
    return B * n / (time.perf_counter() - t0)


LAG = False
print(f"device-resident input through predict(): {run([h.to(dev) for h in host]):.0f} img/s (python + GPU, no PCIe)")
for LAG in (False, True):
  for fly in (2,):
    for chunks in (1, 2, 4, 8):
        DetectionPredictor.in_flight = fly
        DetectionPredictor.pipeline_chunks = chunks
        yl.predictor = None
        print(f"read-one-behind={LAG} in_flight={fly} chunks={chunks}: {run(host):.0f} img/s")
