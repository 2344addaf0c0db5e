#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into a small text file for profiles/:  python tools/ncu_summary.py rep out.txt"""
import csv
import io
import re
import subprocess
import sys

PAT = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                 r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum(\.pct_of_peak_sustained_elapsed)?|launch__registers_per_thread|"
                 r"launch__(grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_.*)|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                 r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                 r"smsp__inst_executed\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sector(_op_read)?_hit_rate\.pct|"
                 r"sm__cycles_elapsed\.max|sm__pipe_tensor.*cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_(alu|fma|lsu|uniform|tensor.*)\.sum)$")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"# kernel: {name}\n# source: {rep} (ncu --set full --clock-control none)\n")
            for i, h in enumerate(hdr):
                if PAT.match(h):
                    f.write(f"{h:95s} {units[i]:16s} {r[i]}\n")
            f.write("\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
