#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report (needs -lineinfo and
--import-source on).  usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = ""
rows = []
hdr = None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].strip().isdigit():
        try:
            inst = int(r[hdr.index("Instructions Executed")])
            samples = int(r[hdr.index("# Samples")])
        except Exception:
            continue
        rows.append((inst, samples, cur_file, int(r[0]), r[1].strip()[:110]))
tot_i = sum(r[0] for r in rows) or 1
tot_s = sum(r[1] for r in rows) or 1
print(f"total instructions {tot_i:,}  samples {tot_s:,}")
for inst, samples, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{100 * inst / tot_i:5.1f}% inst {100 * samples / tot_s:5.1f}% smp  {f}:{ln:<4} {src}")
