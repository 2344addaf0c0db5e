#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures of THIS tree: per-launch DRAM bytes and the busiest pipes of the dominant
kernels, keyed by the hash of densephrases_b200/csrc so that bench.py only quotes a number measured on the sources it runs.

    python tools/make_traffic.py KEY=report.ncu-rep [KEY=report.ncu-rep ...]
KEY = "<kernel>|<workload>|nprobe<P>|n<world>" (what bench.py:lookup_traffic asks for), e.g. "scan_quad_kernel|C2|nprobe256|n1"."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

WANT = {"dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write", "gpu__time_duration.sum": "duration",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "lts__t_sector_op_read_hit_rate.pct": "l2_read_hit_pct"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def parse(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    out = {"kernel_name": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
    for i, h in enumerate(hdr):
        name = h.split(".", 2)[-1] if h.startswith(("SM_", "TPC.")) else h
        if h in WANT:
            out[WANT[h]] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        elif name.startswith("sm__inst_executed_pipe_alu") and "pct_of_peak" in name:
            out["alu_pipe_pct"] = float(r[i].replace(",", ""))
    return out


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    sha = bench.csrc_sha()
    tj = {"csrc_sha": sha, "entries": {}}
    if os.path.exists(path):
        old = json.load(open(path))
        if old.get("csrc_sha") == sha:
            tj = old
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        m = parse(rep)
        pipes = {k: m[k] for k in ("alu_pipe_pct", "lsu_data_pipe_pct", "issue_active_pct", "dram_pct", "tensor_pipe_pct") if k in m}
        top = max(pipes, key=pipes.get) if pipes else None
        tj["entries"][key] = {"dram_bytes_per_launch": m.get("dram_bytes_read", 0.0) + m.get("dram_bytes_write", 0.0), "kernel_seconds_under_ncu": m.get("duration"),
                              "report": os.path.basename(rep), "kernel": m["kernel_name"], "pipes_pct_of_peak": pipes,
                              "limiter": {"pipe": top, "pct_of_peak": pipes.get(top)} if top else None, "l2_read_hit_pct": m.get("l2_read_hit_pct")}
        print(key, json.dumps(tj["entries"][key]))
    json.dump(tj, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
