"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

    python tools/summarize_launches.py gpurun_out/<tag>/launches.csv [--last-step N | --range A:B] > profiles/<name>.md

ncu's per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_breakdown, not
absolutes.  With --last-step N only the final N launches (one step of the plan + NMS) are counted."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    last = int(sys.argv[sys.argv.index("--last-step") + 1]) if "--last-step" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        ours = name.startswith("yl::") or "yl_" in name
        rows.append((name, float(r["Metric Value"]) / 1e3, ours, r["Grid Size"], r["Block Size"]))
    if last:
        rows = rows[-last:]
    if "--range" in sys.argv:
        a, b = sys.argv[sys.argv.index("--range") + 1].split(":")
        rows = rows[int(a):int(b)]
    tot = sum(t for _, t, _, _, _ in rows)
    agg = OrderedDict()
    for name, t, ours, _, _ in rows:
        a = agg.setdefault(name, [0, 0.0, ours])
        a[0] += 1
        a[1] += t
    print(f"# launch list summary: {path}\n")
    print(f"{len(rows)} launches, {tot:.1f} us total device time (ncu-serialised, cold cache)\n")
    print("| kernel | ours | launches | total us | share | avg us |")
    print("|---|---|---:|---:|---:|---:|")
    for name, (n, t, ours) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {'yes' if ours else 'torch'} | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.2f} |")


if __name__ == "__main__":
    main()
