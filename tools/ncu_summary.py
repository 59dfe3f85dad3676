#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: ncu_summary.py <report.ncu-rep> <out.txt> [note]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_op_write.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.sum",
    "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full summary of %s\n# %s\n" % (rep, note))
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            f.write("\n== kernel: %s\n" % name)
            for i, k in enumerate(h):
                if k in KEYS:
                    f.write("%-70s %-14s %s\n" % (k, u[i], r[i]))
            f.write("-- warp stall reasons (per issue active) --\n")
            st = [(float(r[i]), k) for i, k in enumerate(h)
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and r[i]]
            for v, k in sorted(st, reverse=True)[:8]:
                f.write("%-70s %.3f\n" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    print("wrote", out)


if __name__ == "__main__":
    main()
