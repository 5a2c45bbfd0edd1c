"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`): launches, mean duration, mean DRAM bytes, DRAM GB/s and its fraction of the measured
copy peak (MEASURED_PEAKS.json).  Usage: python profiles/summarize_launch_list.py <csv> [peak_GBps] > <md>"""
import collections
import csv
import json
import os
import sys


def main():
    path = sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if peak is None:
        try:
            pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
            peak = float(pk.get("hbm_copy_GBps_burst") or pk.get("hbm_GBps") or 6457.1)
        except Exception:
            peak = 6457.1
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols = rows[hdr]
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(cols):
            continue
        d = dict(zip(cols, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        key = (d["Kernel Name"].split("(")[0].replace("void ", ""), d["Grid Size"], d["Block Size"])
        per.setdefault(key, collections.defaultdict(dict))[d["ID"]][d["Metric Name"]] = v
    print(f"Source: `{os.path.basename(path)}`; peak = {peak:.1f} GB/s (measured copy bandwidth).  Launches under ncu are")
    print("serialised and cold-cache: the durations are upper bounds of what the kernels cost inside an iteration.\n")
    print("| kernel | grid | block | launches | mean µs | DRAM read MB | DRAM write MB | DRAM GB/s | of peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    for (name, grid, block), launches in per.items():
        full = [m for m in launches.values() if "gpu__time_duration.sum" in m]
        if not full:
            continue
        t = sum(m["gpu__time_duration.sum"] for m in full) / len(full) * 1e-9
        rd = sum(m.get("dram__bytes_read.sum", 0.0) for m in full) / len(full)
        wr = sum(m.get("dram__bytes_write.sum", 0.0) for m in full) / len(full)
        gbps = (rd + wr) / t * 1e-9
        print(f"| `{name}` | {grid} | {block} | {len(full)} | {t * 1e6:.1f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | "
              f"{gbps:.0f} | {gbps / peak:.3f} |")


if __name__ == "__main__":
    main()
