#!/usr/bin/env python
"""gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum over `bench.py --steps 1 --warmup 1`) -> profiles/<tag>_launches.md:
the kernel launches of ONE classify step (between the second and the third extract_kernel of the run; the synthetic-index
generator's own torch sorts at the start of the run are left out)."""
import csv
import io
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
txt = open(os.path.join(ROOT, "gpurun_out", "launches.csv"), errors="replace").read()
rows = list(csv.DictReader(io.StringIO(txt[txt.find('"ID"'):])))
seq = []
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"<.*", "", k).split("::")[-1][:44]
    v = float(r["Metric Value"].replace(",", ""))
    seq.append((k, v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r.get("Metric Unit", "ns"), 1) / 1e6))
idx = [i for i, (k, _) in enumerate(seq) if k == "extract_kernel"]
a, b = (idx[1] - 2, idx[2] - 2) if len(idx) > 2 else (idx[-1] - 2, len(seq))
step = seq[a:b]
agg = {}
for k, ms in step:
    x = agg.setdefault(k, [0, 0.0])
    x[0] += 1
    x[1] += ms
tot = sum(v[1] for v in agg.values())
lines = [f"# {tag}: kernel launches of ONE classify step (10 M reads vs 8.21 GiB index)", "",
         "ncu `--metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 1 --warmup 1`; launches are serialised and",
         "cold-cache under ncu — compare shares, not absolute times.  Rows = the launches between the second and the third",
         "`extract_kernel` of the run (the synthetic-index generator's own torch sorts at the start of the run are left out).", "",
         "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k in sorted(agg, key=lambda k: -agg[k][1]):
    lines.append(f"| {k} | {agg[k][0]} | {agg[k][1]:.3f} | {100 * agg[k][1] / tot:.1f}% |")
lines.append(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.3f} | |")
out = []
for k, ms in step:
    if out and out[-1][0] == k:
        out[-1][1] += ms
        out[-1][2] += 1
    else:
        out.append([k, ms, 1])
lines += ["", "Sequence (consecutive launches of the same kernel merged, > 0.3 ms; the score_* groups are the 2^20-read chunks):", ""]
for k, ms, n in out:
    if ms > 0.3:
        lines.append(f"* {k} x{n}: {ms:.2f} ms")
open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[6:22]))
