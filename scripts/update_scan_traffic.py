#!/usr/bin/env python
"""Reads the ncu --set full captures of the scan kernels (scripts/gpu_r2_profile.sh -> gpurun_out/r2_*.ncu-rep) with `ncu -i ... --page raw
--csv`, writes profiles/scan_traffic.json (DRAM bytes per launch, stamped with the commit and the SHA-256 of the kernel sources, so
bench.py can refuse a stale figure) and profiles/r02_scan_ncu.md (the table a reader checks the roofline claims against)."""
import csv
import datetime
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["spiral_b200/csrc/spiral_kernels.cu", "spiral_b200/csrc/scan_tma.cu", "spiral_b200/csrc/pack_kernels.cu"]
CAPS = {  # capture -> (json key, description, algorithmic bytes)
    "scan_cfg1": ("dram_bytes_per_launch", "k_scan_spiral_tma, cfg1 ./spiral 8 7 (2 GiB)", 2 << 30),
    "scan_cfg5": ("dram_bytes_per_launch_cfg5_1gpu", "k_scan_spiral_tma, cfg5 ./spiral 9 8 (8 GiB)", 8 << 30),
    "scan_shard64_9_5": ("dram_bytes_per_launch_cfg5_8gpu", "k_scan_spiral_tma<64>, ./spiral 9 5 = cfg5's 1 GiB shard of an 8-GPU run (64 columns)", 1 << 30),
    "scan_pack_cfg4": ("dram_bytes_per_launch_cfg4_1gpu", "k_scan_pack_narrow, cfg4 ./spiral 11 3 (25 planes x 8 columns, 6.25 GiB)", 25 * 8 * 2048 * (1 << 14)),
    "scan_pack_cfg3": ("dram_bytes_per_launch_cfg3_1gpu", "k_scan_pack_wide, cfg3 ./spiral 10 8 (16 planes x 256 columns, 64 GiB)", 64 << 30),
}
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "lts__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def read_rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {}
        for h, u, v in zip(head, units, r):
            d[h] = (v, u)
        res.append(d)
    return res


def main():
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    tj = {"commit": commit, "captured": datetime.date.today().isoformat(),
          "kernel_source_sha256": {s: hashlib.sha256(open(os.path.join(ROOT, s), "rb").read()).hexdigest() for s in SOURCES},
          "source": "ncu --set full --clock-control none, scripts/gpu_r2_profile.sh; table in profiles/r02_scan_ncu.md"}
    md = ["# Scan kernels under `ncu --set full --clock-control none` (round 2, scripts/gpu_r2_profile.sh; one launch each)", "",
          f"Captured at commit {commit}.  `traffic` = dram__bytes_read + dram__bytes_write of that launch; `algorithmic` = 8 B x NTT coefficients of the shard (SURVEY 8d).", ""]
    for cap, (key, desc, algo) in CAPS.items():
        path = os.path.join(ROOT, "gpurun_out", f"r2_{cap}.ncu-rep")
        if not os.path.exists(path):
            md += [f"## {desc}", "", "capture missing", ""]
            continue
        d = read_rep(path)[0]
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        tj[key] = rd + wr
        dur_v, dur_u = d["gpu__time_duration.sum"]
        dur_us = float(dur_v.replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(dur_u, 1)
        md += [f"## {desc}", "", f"kernel `{d.get('Kernel Name', ('?', ''))[0]}`", "", "| metric | value |", "|---|---:|",
               f"| duration | {dur_us:.1f} us |", f"| algorithmic bytes | {algo / 1e9:.3f} GB |", f"| DRAM traffic (read + write) | {(rd + wr) / 1e9:.3f} GB ({(rd + wr) / algo:.3f} x algorithmic) |",
               f"| GB/s under the profiler (algorithmic / duration) | {algo / dur_us / 1e3:.0f} |"]
        for m in METRICS[3:]:
            if m in d:
                md.append(f"| {m} | {d[m][0]} {d[m][1]} |")
        md.append("")
    json.dump(tj, open(os.path.join(ROOT, "profiles", "scan_traffic.json"), "w"), indent=1)
    open(os.path.join(ROOT, "profiles", "r02_scan_ncu.md"), "w").write("\n".join(md) + "\n")
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    sys.exit(main())
