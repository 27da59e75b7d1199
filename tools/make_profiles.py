"""Turn one `tools/gpu_profile_round.sh <tag>` result (gpurun_out/<tag>/) into the tracked summaries under profiles/.

    python tools/make_profiles.py r01n

Writes profiles/<tag>_bench*.json, <tag>_layers_*.txt, <tag>_launches_yolo11n_bs64.md, <tag>_ncu_conv_tc.md,
<tag>_ncu_other_kernels.md and refreshes profiles/traffic.json (read by bench.py for roofline.traffic)."""
import csv
import io
import json
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
src = ROOT / "gpurun_out" / tag
dst = ROOT / "profiles"

for a, b in (("bench.json", "bench.json"), ("bench_ref.json", "bench_reference_arm.json"),
             ("bench_m256.json", "bench_yolo11m_bs256.json"), ("layers_n64.txt", "layers_yolo11n_bs64.txt"),
             ("layers_m64.txt", "layers_yolo11m_bs64.txt"), ("layers_n1.txt", "layers_yolo11n_bs1.txt")):
    if (src / a).exists():
        shutil.copy(src / a, dst / f"{tag}_{b}")

bench = json.loads((src / "bench.json").read_text().splitlines()[-1])
per_step = bench["launches_per_step"]

# launch list + traffic of the LAST FULL STEP: the launches from the last image-ingest kernel (stem_*) on
def last_step_csv(path: Path, n: int) -> Path:
    lines = path.read_text().splitlines()
    head = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.reader(lines[head + 1:]))
    hdr = next(csv.reader([lines[head]]))
    iid, ik = hdr.index("ID"), hdr.index("Kernel Name")
    starts = sorted({int(r[iid]) for r in rows if len(r) > ik and "stem_" in r[ik]})
    first = starts[-1]
    keep = [l for l, r in zip(lines[head + 1:], rows) if len(r) > iid and first <= int(r[iid]) < first + n]
    out = path.with_name("launches_last_step.csv")
    out.write_text("\n".join([lines[head]] + keep) + "\n")
    return out


step_csv = last_step_csv(src / "launches.csv", per_step)
out = subprocess.run([sys.executable, str(ROOT / "tools" / "summarize_launches.py"), str(step_csv)], capture_output=True, text=True)
(dst / f"{tag}_launches_yolo11n_bs64.md").write_text(out.stdout.replace(str(ROOT) + "/", ""))
subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_traffic.py"), str(step_csv), "yolo11n_bs64"], check=False)

METRICS = [("gpu__time_duration.sum", "us", 1e-3 if False else None),
           ("dram__bytes_read.sum", "DRAM rd MB", None), ("dram__bytes_write.sum", "DRAM wr MB", None),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", None),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", None),
           ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe %", None),
           ("lts__t_sector_hit_rate.pct", "L2 hit %", None),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", None),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", None),
           ("launch__registers_per_thread", "regs", None), ("launch__grid_size", "grid", None)]


def raw_table(rep: Path):
    r = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return [], []
    return rows[0], rows[1:]


def fmt(v, unit):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    if unit in ("ns", "nsecond"):
        f /= 1e3
    if unit in ("byte",):
        f /= 1e6
    if unit in ("Kbyte",):
        f /= 1e3
    if unit in ("Gbyte",):
        f *= 1e3
    if unit in ("ms", "msecond"):
        f *= 1e3
    return f"{f:.1f}" if abs(f) < 1e5 else f"{f:.0f}"


def write_md(rep: Path, path: Path, title: str, cmd: str, names=None):
    h, rows = raw_table(rep)
    if not h:
        path.write_text(f"# {title}\n\n(no data: {rep.name} missing or unreadable)\n")
        return
    units, data = rows[0], rows[1:]
    ik = h.index("Kernel Name")
    cols = [(h.index(m), lab) for m, lab, _ in METRICS if m in h]
    lines = [f"# {title}", "", f"command: `{cmd}` (report: gpurun_out/{tag}/{rep.name}, not committed)", "",
             "| # | kernel | " + " | ".join(lab for _, lab in cols) + " |", "|---|---|" + "---:|" * len(cols)]
    for n, r in enumerate(data):
        k = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("yl::", "")
        if names and n < len(names):
            k = f"{k} — {names[n]}"
        lines.append(f"| {n} | {k} | " + " | ".join(fmt(r[i], units[i]) for i, _ in cols) + " |")
    lines += ["", "Per-launch times under ncu are cold-cache, serialised and at ncu's clocks; the in-plan per-layer times are in "
              f"{tag}_layers_yolo11n_bs64.txt."]
    path.write_text("\n".join(lines) + "\n")


# names of the conv_tc launches captured (-s 240 -c 12 over the conv_tc launches of the run): launches of one pass
layers = [l for l in (src / "layers_n64.txt").read_text().splitlines() if " conv_tc " in l]
convs = [" ".join(l.split()[2:5]) for l in layers]
n_conv = len(convs)
names = [convs[(240 + i) % n_conv] for i in range(12)] if n_conv else None
write_md(src / "conv_tc.ncu-rep", dst / f"{tag}_ncu_conv_tc.md",
         f"ncu --set full, conv_tc_kernel, yolo11n bs=64, 12 consecutive launches of the plan ({tag})",
         "ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 240 -c 12 python bench.py --steps 1 "
         "--warmup 3 --inflight 1 --no-cpu-baseline --no-e2e --no-latency", names)
write_md(src / "others.ncu-rep", dst / f"{tag}_ncu_other_kernels.md",
         f"ncu --set full, the other kernels of one yolo11n bs=64 step ({tag})",
         "ncu --set full --clock-control none --import-source on -k regex:\"c3k2_tail|dwconv3x3_mma|stem_|letterbox|nms_|psa_att|sppf\" "
         "-s 12 -c 16 python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline --no-latency")
print("profiles written for", tag)
