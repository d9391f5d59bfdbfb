#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) from an .ncu-rep captured with
--import-source on:  python profiles/source_hot.py gpurun_out/x.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
  raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True,
                       check=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr = None
  cur_file = ""
  lines = []
  total = 0
  for r in rows:
    if len(r) == 2 and r[0] == "File Path":
      cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
      hdr = r
    elif hdr and len(r) == len(hdr) and r[0]:
      inst = int(r[hdr.index("Instructions Executed")] or 0)
      samples = int(r[hdr.index("# Samples")] or 0)
      lines.append((inst, samples, cur_file, r[0], r[1].strip()[:110]))
      total += inst
  print(f"# {path}: {total} warp instructions attributed to source lines")
  tot_s = sum(l[1] for l in lines) or 1
  for inst, samples, f, no, src in sorted(lines, reverse=True)[:top]:
    print(f"{100.0 * inst / total:5.1f}% inst {100.0 * samples / tot_s:5.1f}% stall  {f}:{no}  {src}")


if __name__ == "__main__":
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
