#!/usr/bin/env python3
"""Condense an Nsight Compute report into a small, committable summary (profiles/*.md).

usage: tools/ncu_summary.py <report.ncu-rep> <out.md> [title]
Reads `ncu -i <rep> --page raw --csv` here on the CPU box (ncu needs no GPU to read a report)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_elapsed.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none), read with `ncu -i ... --page raw --csv`", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines += [f"## kernel `{d.get('Kernel Name', '?')}` grid {d.get('Grid Size','?')} block {d.get('Block Size','?')}", "",
                  "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u[k]} |")
        lines += ["", "stall reasons (warps per issue-active cycle):", "", "| reason | ratio |", "|---|---|"]
        st = [(k[len(STALL):].replace("_per_issue_active.ratio", ""), float(d[k] or 0)) for k in hdr
              if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")]
        for name, v in sorted(st, key=lambda x: -x[1]):
            if v >= 0.01:
                lines.append(f"| {name} | {v:.3f} |")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main()
