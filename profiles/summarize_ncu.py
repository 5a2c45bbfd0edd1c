#!/usr/bin/env python
"""Turns one `ncu --set full --import-source on` capture of k_sweep into the markdown summary kept under profiles/.

  python profiles/summarize_ncu.py gpurun_out/sweep_prof.ncu-rep <steps in the profiled launch> <algorithmic MB/step> > profiles/rNN_ncu_k_sweep.md

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source cuda,sass` (ncu is in the
image; no GPU needed)."""
import collections
import csv
import io
import subprocess
import sys


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, steps, mb_step = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    raw = ncu_csv(rep, "--page", "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    val = {h: (row[i], units[i]) for i, h in enumerate(hdr)}

    def f(name):
        return float(val[name][0].replace(",", ""))

    unit_scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = f("dram__bytes_read.sum") * unit_scale[val["dram__bytes_read.sum"][1]]
    wr = f("dram__bytes_write.sum") * unit_scale[val["dram__bytes_write.sum"][1]]
    dur_ms = f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3}[val["gpu__time_duration.sum"][1]]
    print(f"# ncu --set full of `{val['Kernel Name'][0].strip()}`\n")
    print(f"Report: `{rep}` ({steps} time steps in the profiled launch; a number measured under ncu is never a bench value).\n")
    print("| metric | value |\n|---|---|")
    print(f"| duration | {dur_ms:.3f} ms = {dur_ms / steps * 1e3:.1f} us per time step |")
    print(f"| grid x block, registers | {val['launch__grid_size'][0]} x {val['launch__block_size'][0]}, {val['launch__registers_per_thread'][0]} regs/thread |")
    print(f"| DRAM read + write | {rd / 1e9:.3f} + {wr / 1e9:.3f} GB = {(rd + wr) / steps / 1e6:.1f} MB per step |")
    print(f"| algorithmic bytes (tree, SURVEY 8d) | {mb_step:.1f} MB per step -> traffic / algorithmic = {(rd + wr) / steps / 1e6 / mb_step:.3f} |")
    print(f"| DRAM throughput under ncu | {(rd + wr) / 1e9 / (dur_ms * 1e-3):.0f} GB/s |")
    for name, label in [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
                        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (of 64/SM)"),
                        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per executed instruction"),
                        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
                        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
                        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
                        ("smsp__inst_executed.sum", "warp instructions")]:
        if name in val:
            print(f"| {label} | {val[name][0]} {val[name][1]} |")
    # stall samples and instructions by source line
    src = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    h = next(r for r in src if r and r[0] == "Line No")
    i_s, i_e = h.index("# Samples"), h.index("Instructions Executed")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    per = collections.OrderedDict()
    cur, fname = None, "?"
    tot_s = tot_e = 0
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) < len(h) or r[0] == "Line No":
            continue
        if r[0] != "":
            try:
                cur = (fname, int(r[0]), r[1].strip())
            except ValueError:
                pass
            continue
        try:
            ns, ne = int(r[i_s]), int(r[i_e])
        except ValueError:
            continue
        d = per.setdefault(cur, [0, 0, collections.Counter()])
        d[0] += ns
        d[1] += ne
        tot_s += ns
        tot_e += ne
        for c in stalls:
            try:
                d[2][c] += int(r[h.index(c)])
            except ValueError:
                pass
    print(f"\n## Hottest source lines (warp-state samples: {tot_s}, warp instructions: {tot_e})\n")
    print("| file:line | samples | instr | top stall reasons | source |\n|---|---|---|---|---|")
    for (fn, ln, text), (ns, ne, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:25]:
        why = ", ".join(f"{k.replace('stall_', '')} {100 * v // max(1, ns)}%" for k, v in st.most_common(2))
        print(f"| {fn}:{ln} | {100 * ns / tot_s:.1f}% | {100 * ne / tot_e:.1f}% | {why} | `{text[:90].replace('|', '/')}` |")


if __name__ == "__main__":
    main()
