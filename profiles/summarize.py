#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the short per-launch summary committed under profiles/.

  python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/r01_<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def main(path):
  raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                       text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  name_col = hdr.index("Kernel Name")
  print(f"# {path}: {len(rows) - 2} launch(es); ncu --set full --clock-control none")
  for r in rows[2:]:
    print(f"\n## {r[name_col][:110]}")
    for k in KEYS:
      if k in hdr:
        i = hdr.index(k)
        print(f"{k:85s} {r[i]:>16s} {units[i]}")
    if "dram__bytes_read.sum" in hdr:
      conv = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
      tot = sum(float(r[hdr.index(k)]) * conv.get(units[hdr.index(k)], 1.0)
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
      dur = float(r[hdr.index("gpu__time_duration.sum")])
      du = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
      print(f"{'traffic (dram read + write), bytes':85s} {tot:16.4g}")
      print(f"{'dram GB/s under ncu (cold, serialised: not a bench number)':85s} {tot / (dur * du) / 1e9:16.1f}")


if __name__ == "__main__":
  main(sys.argv[1])
