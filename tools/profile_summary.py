#!/usr/bin/env python
"""Turn ncu artefacts of a GPU visit (gpurun_out/, scratch) into the small tracked summaries under profiles/.

  tools/profile_summary.py launches <launches.csv> <tag>            -> profiles/<tag>_launches.md  (one classify step)
  tools/profile_summary.py kernel <report.ncu-rep> <tag> <name> [--lines N] [--json]
        -> profiles/<tag>_<name>_ncu.md (raw metrics per launch + top source lines); --json also refreshes
           profiles/merge_ncu_summary.json (dram bytes per launch, what bench.py reports as roofline.traffic)
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1][:48]


def launches(path, tag):
    txt = open(path, errors="replace").read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.find('"ID"'):])))
    seq = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        seq.append((short(r["Kernel Name"]), v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r.get("Metric Unit", "ns"), 1) / 1e6))
    idx = [i for i, (k, _) in enumerate(seq) if k == "extract_kernel"]
    a, b = (idx[1] - 2, idx[2] - 2) if len(idx) > 2 else (max(0, idx[-1] - 2), len(seq))
    step = seq[a:b]
    agg = {}
    for k, ms in step:
        x = agg.setdefault(k, [0, 0.0])
        x[0] += 1
        x[1] += ms
    total = sum(v[1] for v in agg.values()) or 1
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: kernel launches of ONE classify step (10 M reads vs 8.21 GiB index)\n\n"
                "ncu `--metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 1 --warmup 1`; launches are serialised and\n"
                "cold-cache under ncu — compare shares, not absolute times.  Rows = the launches of the last complete step of the run.\n\n"
                "| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k in sorted(agg, key=lambda k: -agg[k][1]):
            f.write(f"| {k} | {agg[k][0]} | {agg[k][1]:.3f} | {100 * agg[k][1] / total:.1f}% |\n")
        f.write(f"| **total** | {len(step)} | {total:.3f} | |\n")
    print("wrote", f"{tag}_launches.md", len(step), "launches", round(total, 1), "ms")


def source_lines(rep, top):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr, rows = "", None, []
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[0].strip().isdigit():
            try:
                rows.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), cur, int(r[0]), r[1].strip()[:100]))
            except Exception:
                continue
    ti = sum(r[0] for r in rows) or 1
    ts = sum(r[1] for r in rows) or 1
    by_file = {}
    for i, s, f, _, _ in rows:
        x = by_file.setdefault(f, [0, 0])
        x[0] += i
        x[1] += s
    return ti, ts, by_file, sorted(rows, reverse=True)[:top]


def kernel(rep, tag, name, top, want_json):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    dram = []
    with open(os.path.join(OUT, f"{tag}_{name}_ncu.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none --import-source on, {name} (per launch; times under ncu are not bench values)\n\n")
        for li, d in enumerate(data):
            f.write(f"## launch {li}: {short(d[idx['Kernel Name']])}\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT[1:]:
                if w in idx:
                    f.write(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n")
            f.write("\n")
            try:
                def to_bytes(n):
                    return float(d[idx[n]].replace(",", "")) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[idx[n]].lower(), 1)
                dram.append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
            except Exception:
                pass
        if top:
            ti, ts, by_file, lines = source_lines(rep, top)
            f.write(f"## source view (all launches of the report): {ti:,} warp instructions, {ts:,} stall samples\n\n| file | instructions | samples |\n|---|---|---|\n")
            for fn, (i, s) in sorted(by_file.items(), key=lambda x: -x[1][0])[:8]:
                f.write(f"| {fn} | {100 * i / ti:.1f}% | {100 * s / ts:.1f}% |\n")
            f.write("\n| instructions | samples | line | source |\n|---|---|---|---|\n")
            for i, s, fn, ln, src in lines:
                f.write(f"| {100 * i / ti:.1f}% | {100 * s / ts:.1f}% | {fn}:{ln} | `{src.replace('|', '/')}` |\n")
    if want_json and dram:
        json.dump({"dram_bytes_per_launch": sum(dram) / len(dram), "launches": len(dram), "source": f"profiles/{tag}_{name}_ncu.md",
                   "note": "ncu workload = bench.py default workload (10M reads vs ~8 GiB index), one merge launch"},
                  open(os.path.join(OUT, "merge_ncu_summary.json"), "w"))
    print("wrote", f"{tag}_{name}_ncu.md", len(data), "launches")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        top = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 40
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], top, "--json" in sys.argv)
