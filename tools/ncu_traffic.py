"""Per-kernel DRAM traffic from an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
launch log of one bench step -> profiles/traffic.json (read by bench.py for `roofline.traffic`).

    python tools/ncu_traffic.py gpurun_out/<tag>/traffic.csv yolo11n_bs64 [--last N]

`--last N`: only the final N launches (= one step: plan launches + NMS) are counted.
"""
import csv
import json
import re
import sys
from pathlib import Path

KIND = {"conv_tc_kernel": "conv_tc", "stem_conv_kernel": "stem_conv", "dwconv3x3_kernel": "dwconv3x3",
        "dwconv3x3_mma_kernel": "dwconv3x3", "c3k2_tail_kernel": "c3k2_tail", "stem_fused_kernel": "stem_fused", "letterbox_u8_kernel": "letterbox_u8",
        "psa_attention_kernel": "psa_attention", "sppf_pool_kernel": "sppf_pool", "nms_select_kernel": "nms_select",
        "nms_filter_kernel": "nms_filter", "detect_decode_kernel": "detect_decode"}


def main():
    path, key = sys.argv[1], sys.argv[2]
    last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per_id = {}
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        e = per_id.setdefault(i, {"name": r["Kernel Name"], "rd": 0.0, "wr": 0.0, "us": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3}.get(unit, 1)
        if r["Metric Name"] == "dram__bytes_read.sum":
            e["rd"] = v * scale
        elif r["Metric Name"] == "dram__bytes_write.sum":
            e["wr"] = v * scale
        elif r["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = v * scale
    ids = sorted(per_id)
    if last:
        ids = ids[-last:]
    out = {}
    for i in ids:
        e = per_id[i]
        m = re.search(r"(\w+_kernel)", e["name"])
        kind = KIND.get(m.group(1) if m else "", None)
        if kind is None:
            continue
        o = out.setdefault(kind, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "us": 0.0})
        o["launches"] += 1
        o["dram_read"] += e["rd"]
        o["dram_write"] += e["wr"]
        o["us"] += e["us"]
    for o in out.values():
        o["dram_bytes_per_launch"] = (o["dram_read"] + o["dram_write"]) / o["launches"]
    dst = Path(__file__).resolve().parents[1] / "profiles" / "traffic.json"
    d = json.loads(dst.read_text()) if dst.exists() else {}
    d[key] = out
    d.setdefault("_how", "ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                 "gpu__time_duration.sum over one bench step (cold-cache, serialised launches)")
    dst.write_text(json.dumps(d, indent=1, sort_keys=True))
    for k, o in out.items():
        print(f"{k:<16}{o['launches']:>4} launches  rd {o['dram_read'] / 1e6:9.1f} MB  wr {o['dram_write'] / 1e6:9.1f} MB  {o['us']:9.1f} us")


if __name__ == "__main__":
    main()
