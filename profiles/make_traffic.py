#!/usr/bin/env python
"""profiles/r02_traffic.json from this round's `ncu --set full` captures.

  python profiles/make_traffic.py gpurun_out/r02_z_*.ncu-rep

One entry per (workload, kernel): DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum), the launch grid / block and the batch of the captured run (taken from the
file name: r02_z_<workload>_b<batch>.ncu-rep).  bench.py copies `bytes` into roofline.traffic only
when kernel name, batch AND grid of its own run match the capture - otherwise the field is null.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def kernels_of(path):
  raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
                                   "launch__block_size", "gpu__time_duration.sum")}
  for r in rows[2:]:
    name = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
    name = re.sub(r"^pgx::", "", name).split("<")[0].split("(")[0]
    total = sum(float(r[col[k]]) * UNIT.get(units[col[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    yield name, total, int(float(r[col["launch__grid_size"]])), int(float(r[col["launch__block_size"]]))


def main(paths):
  out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures of this round "
                     "(profiles/make_traffic.py); bench.py uses an entry only if kernel, batch and grid match its own run"}
  for path in paths:
    m = re.search(r"r02_z_([a-z0-9_]+?)_b(\d+)\.ncu-rep$", os.path.basename(path))
    if not m:
      print("skipping (name is not r02_z_<workload>_b<batch>.ncu-rep):", path)
      continue
    workload, batch = m.group(1), int(m.group(2))
    best = {}
    for name, total, grid, block in kernels_of(path):
      if name not in best or total > best[name][0]:
        best[name] = (total, grid, block)
    for name, (total, grid, block) in best.items():
      out[f"{workload}:{name}"] = {"bytes": int(total), "batch": batch, "grid": grid, "block": block,
                                   "source": "profiles/" + os.path.basename(path).replace(".ncu-rep", ".txt")}
      print(f"{workload}:{name}: {total / 1e6:.1f} MB per launch, grid {grid} x {block}")
  with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_traffic.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
  main(sys.argv[1:])
