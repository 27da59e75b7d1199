"""Debug aid: where does the conv-chain plan first differ from the layer-by-layer plan?  Builds yolo11n twice (same
weights), runs both plans eagerly on the same input and compares every plan buffer in allocation order.

    python tools/dbg_chain.py [batch] [graph]
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from bench import randomise_model_  # noqa: E402
from yololite.nn.tasks import DetectionModel  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
os.environ["YL_CHAIN_MIN_BATCH"] = "1"
x = torch.rand(batch, 3, 640, 640, generator=torch.Generator().manual_seed(9)).cuda()
plans = {}
for mode in ("0", "1"):
    os.environ["YL_CHAIN"] = mode
    torch.manual_seed(4)
    m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval().cuda()
    m.use_cuda_graph = graph
    for rep in range(3):
        y, _ = m.infer(x)
        torch.cuda.synchronize()
        if rep == 0:
            y0 = y.clone()
        else:
            print(f"mode {mode} rep {rep}: deterministic = {torch.equal(y, y0)}")
    plans[mode] = (m, m._get_plan(x.shape, x.device)[0], y.clone())
p0, p1 = plans["0"][1], plans["1"][1]
print("y equal:", torch.equal(plans["0"][2], plans["1"][2]), float((plans["0"][2] - plans["1"][2]).abs().max()))
print("launches", len(p0.calls), len(p1.calls))
for md in p1.meta:
    if md["kind"] == "conv_chain":
        print("chain:", len(md["members"]), [mm["desc"] for mm in md["members"]])
b0 = [b for b in p0.buffers if b.dtype in (torch.bfloat16, torch.float32)]
b1 = [b for b in p1.buffers if b.dtype in (torch.bfloat16, torch.float32)]
print("buffers", len(b0), len(b1))
bad = 0
for i, (a, b) in enumerate(zip(b0, b1)):
    if a.shape != b.shape:
        print(i, "shape mismatch", a.shape, b.shape)
        continue
    if not torch.equal(a, b):
        d = (a.float() - b.float()).abs()
        nz = (d > 0)
        # which images / channels differ
        if a.dim() == 4:
            per_img = nz.flatten(1).sum(1).tolist()
            per_ch = nz.sum((0, 1, 2))
            chs = per_ch.nonzero().flatten().tolist()
            rows = nz.sum((0, 2, 3)).nonzero().flatten().tolist()
            print(f"buffer {i} {tuple(a.shape)}: {int(nz.sum())} elems differ, max {float(d.max()):.4f}, per image {per_img}, "
                  f"channels {chs[:4]}..{chs[-4:]} ({len(chs)}), rows {rows[:6]}..{rows[-3:]}")
        else:
            print(f"buffer {i} {tuple(a.shape)}: {int(nz.sum())} elems differ, max {float(d.max()):.4f}")
        bad += 1
        if bad >= 6:
            break
print("done, differing buffers:", bad)
