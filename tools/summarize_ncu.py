#!/usr/bin/env python
"""Turn the ncu artefacts a GPU visit left in gpurun_out/ into the small, tracked summaries under profiles/.

  launches.csv        (ncu --metrics gpu__time_duration.sum ...)  -> profiles/<tag>_launches.md  (share per kernel)
  merge_prof.ncu-rep  (ncu --set full -k regex:merge_kernel)       -> profiles/<tag>_merge_ncu.md + merge_ncu_summary.json
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1][:60]


launch_csv = os.path.join(G, "launches.csv")
if os.path.exists(launch_csv):
    txt = open(launch_csv, errors="replace").read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:]))) if start >= 0 else []
    agg, order = {}, []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        if k not in agg:
            agg[k] = [0, 0.0]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values()) or 1
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: kernel launch list (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised — compare shares)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k in sorted(agg, key=lambda k: -agg[k][1]):
            f.write(f"| {k} | {agg[k][0]} | {agg[k][1] / 1e6:.3f} | {100 * agg[k][1] / total:.1f}% |\n")
    print("wrote", f"{tag}_launches.md", len(rows), "rows")

for kname in ["merge"] + sys.argv[2:]:
  rep = os.path.join(G, f"{kname}_prof.ncu-rep")
  if os.path.exists(rep):
      raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
      rows = list(csv.reader(io.StringIO(raw)))
      hdr, units, data = rows[0], rows[1], rows[2:]
      want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
              "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg",
              "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
              "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
              "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
              "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
              "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]
      idx = {h: i for i, h in enumerate(hdr)}
      with open(os.path.join(out_dir, f"{tag}_{kname}_ncu.md"), "w") as f:
          f.write(f"# {tag}: ncu --set full on {kname} kernel (per launch)\n\n")
          dram = []
          for li, d in enumerate(data):
              f.write(f"## launch {li}\n\n| metric | value | unit |\n|---|---|---|\n")
              for w in want:
                  if w in idx:
                      f.write(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n")
              f.write("\n")
              try:
                  def to_bytes(name):
                      v = float(d[idx[name]].replace(",", ""))
                      u = units[idx[name]].lower()
                      return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                  dram.append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
              except Exception:
                  pass
      if dram and kname == "merge":
          json.dump({"dram_bytes_per_launch": sum(dram) / len(dram), "launches": len(dram), "source": f"profiles/{tag}_merge_ncu.md",
                     "note": "ncu workload = bench.py default workload (10M reads vs ~8 GiB index), one merge launch"},
                    open(os.path.join(out_dir, "merge_ncu_summary.json"), "w"))
      print("wrote", f"{tag}_{kname}_ncu.md", len(data), "launches")
