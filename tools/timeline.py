"""In-kernel timeline of every conv_tc launch of one model plan (CUDA-graph replay): where do the microseconds
of a launch-latency-bound layer go?   python tools/timeline.py [n|s|m] [batch]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from bench import randomise_model_  # noqa: E402
from yololite import _C  # noqa: E402
from yololite.nn.tasks import DetectionModel  # noqa: E402

scale = sys.argv[1] if len(sys.argv) > 1 else "n"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = _C.init(0)
buf = torch.zeros((256, 16), dtype=torch.int64, device="cuda")
lib.yl_debug_timeline(buf.data_ptr(), 256)
m = randomise_model_(DetectionModel(f"yolo11{scale}.yaml", verbose=False)).eval().cuda()
m.model[-1].parallel_branches = False       # one lane: launches are strictly sequential in the graph
x = torch.rand(batch, 3, 640, 640, device="cuda")
m.infer(x)                                   # build: the eager warm-up pass + capture both take slots
lib.yl_debug_timeline(None, 0)
for _ in range(5):
    m.infer(x)
torch.cuda.synchronize()
plan = m._get_plan(x.shape, x.device)[0]
t = buf.cpu().numpy()
convs = [md for md in plan.meta if md["kind"] == "conv_tc"]
n = len(convs)
# slots: first n = eager warm-up pass, next n = capture pass (those parameters are the ones the graph replays)
rows = t[n:2 * n]
names = ["prologue", "dep wait", "operands", "mma", "epilogue", "drain", "exit"]
print(f"yolo11{scale} bs={batch}: {n} conv launches; per-phase ns of CTA 0 (graph replay)")
print(f"{'#':>3} {'shape':<34}" + "".join(f"{k:>10}" for k in names) + f"{'total':>10}{'gap->next':>10}")
tot = [0] * 9
for i, (md, r) in enumerate(zip(convs, rows)):
    d = [int(r[k + 1] - r[k]) for k in range(7)]
    total = int(r[7] - r[0])
    gap = int(rows[i + 1][2] - r[7]) if i + 1 < n else 0   # our exit -> the next kernel's dependency wait returns
    print(f"{i:>3} {md['desc']:<34}" + "".join(f"{v:>10}" for v in d) + f"{total:>10}{gap:>10}")
    for k in range(7):
        tot[k] += d[k]
    tot[7] += total
    tot[8] += gap
    if "--epi" in sys.argv:
        # epilogue detail of the first tile (relative to 'accumulator ready'): ld0 math0 bar0 sts0 | ld1 math1 | ld2 bar2
        e = [int(r[k] - r[4]) if r[k] else -1 for k in (8, 9, 10, 11, 12, 13, 14, 15)]
        print(f"      epi: ld0 {e[0]} math0 {e[1]} bar0 {e[2]} sts0 {e[3]} | ld1 {e[4]} math1 {e[5]} | ld2 {e[6]} bar2 {e[7]}")
print(f"{'':>3} {'mean':<34}" + "".join(f"{v / n:>10.0f}" for v in tot))
