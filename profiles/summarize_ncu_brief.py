#!/usr/bin/env python
"""Brief markdown summary of every kernel result in `ncu --set full` reports (duration, DRAM bytes, FP64 tensor pipe,
issue slots, occupancy, shared-memory conflicts, local-memory traffic).

  python profiles/summarize_ncu_brief.py gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep > profiles/rNN_ncu_kernels.md
"""
import csv
import io
import subprocess
import sys

SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
TIME = {"ms": 1e3, "us": 1.0, "s": 1e6, "ns": 1e-3}
ROWS = [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of 64/SM"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "DMMA warp instructions"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe active % (of SM-active cycles)"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 (non-tensor) pipe active %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
        ("l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "local-memory load wavefronts"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("launch__registers_per_thread", "registers per thread"),
        ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), CTAs/SM"),
        ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM")]


def main():
    print("# ncu --set full, brief per-kernel summaries (numbers under ncu are never bench values)\n")
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "raw"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            val = {h: (row[i], units[i]) for i, h in enumerate(hdr)}

            def f(name):
                return float(val[name][0].replace(",", "")) if name in val and val[name][0] not in ("", "n/a") else float("nan")
            dur = f("gpu__time_duration.sum") * TIME[val["gpu__time_duration.sum"][1]]
            rd = f("dram__bytes_read.sum") * SCALE[val["dram__bytes_read.sum"][1]]
            wr = f("dram__bytes_write.sum") * SCALE[val["dram__bytes_write.sum"][1]]
            dmma = f("sm__inst_executed_pipe_tensor_subpipe_dmma.sum")
            print(f"## `{val['Kernel Name'][0].strip()[:110]}`\n")
            print(f"Report `{rep}`, result {val['ID'][0]}: grid {val['launch__grid_size'][0]} x {val['launch__block_size'][0]} threads.\n")
            print("| metric | value |\n|---|---|")
            print(f"| duration | {dur:.1f} us |")
            print(f"| DRAM read + write | {rd / 1e6:.2f} + {wr / 1e6:.2f} MB -> {(rd + wr) / 1e3 / max(dur, 1e-9):.0f} GB/s |")
            if dmma == dmma and dmma > 0:
                print(f"| DMMA m8n8k4 flops (512 per warp instruction) | {dmma * 512 / 1e9:.3f} GFLOP -> {dmma * 512 / 1e6 / dur:.2f} TFLOP/s |")
            for name, label in ROWS:
                if name in val and val[name][0] not in ("", "n/a"):
                    print(f"| {label} | {val[name][0]} |")
            print()


if __name__ == "__main__":
    main()
