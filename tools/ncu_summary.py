#!/usr/bin/env python
"""Text summary of an `ncu --set full` report: per launch the headline utilisation counters, the warp-stall breakdown
(smsp__pcsamp_warps_issue_stalled_*) and the pipe mix.  Usage: python tools/ncu_summary.py report.ncu-rep > summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.max"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    pcs = [h for h in hdr if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    print(f"# {path}: {len(rows) - 2} launches (ncu --set full --clock-control none; times are cold-cache, serialised)")
    for r in rows[2:]:
        grid = r[idx["Grid Size"]] if "Grid Size" in idx else ""
        block = r[idx["Block Size"]] if "Block Size" in idx else ""
        print(f"\n== {r[idx['Kernel Name']].split('(')[0]}  grid {grid} block {block}")
        for w in WANT:
            if w in idx:
                print(f"   {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}")
        vals = []
        for h in pcs:
            try:
                vals.append((float(r[idx[h]].replace(",", "") or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        tot = sum(v for v, _ in vals) or 1.0
        print("   warp-stall samples (share of all samples): " +
              ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(vals, reverse=True)[:9]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
