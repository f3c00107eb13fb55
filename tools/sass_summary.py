#!/usr/bin/env python
"""SASS evidence per kernel of libmetabuli_b200.so -> profiles/<tag>_sass_summary.md: registers / shared memory (cuobjdump
--dump-resource-usage) and counts of the instructions that show what the kernel is built from (B200_PROFILING.md): UBLKCP =
cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier operations, LDGSTS = cp.async, ATOMS / ATOMG / RED = shared / global
atomics, SHFL / VOTE = warp shuffles and ballots, BAR = block barriers, LDS / STS / LDG / STG = memory instructions.
usage: tools/sass_summary.py [tag]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "metabuli_b200", "_lib", "libmetabuli_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
OPS = ["UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "BAR", "LDS", "STS", "LDG", "STG", "POPC", "IMAD", "LOP3"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for ln in res.split("\n"):
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in ln:
        d = dict(re.findall(r"(\w+):(\d+)", ln))
        usage[cur] = d
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, total = {}, {}
cur = None
arch = set()
for ln in sass.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = {o: 0 for o in OPS}
        total[cur] = 0
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if not m:
        continue
    op = m.group(1)
    total[cur] += 1
    for o in OPS:
        if op == o or op.startswith(o + "."):
            counts[cur][o] += 1
own = [k for k in counts if "mbl" in k and "cub" not in k]
own.sort(key=lambda k: -total[k])
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md"), "w") as f:
    f.write(f"# {tag}: SASS summary of the repo's own kernels in metabuli_b200/_lib/libmetabuli_b200.so (arch: {', '.join(sorted(arch))})\n\n"
            "`cuobjdump -sass` / `--dump-resource-usage`, static instruction counts (tools/sass_summary.py).  UBLKCP = TMA bulk copy "
            "(cp.async.bulk), SYNCS = mbarrier, LDGSTS = cp.async; CUB's radix-sort / scan / select kernels are left out.\n\n")
    f.write("| kernel | regs | smem (static) | SASS instr | " + " | ".join(OPS[:9]) + " | LDS | STS | LDG | STG |\n|---|---|---|---|" + "---|" * 13 + "\n")
    for k in own:
        u = usage.get(k, {})
        name = re.sub(r"\(.*", "", demangle(k).replace("(anonymous namespace)::", "")).replace("mbl::", "")[:60]
        c = counts[k]
        f.write(f"| `{name}` | {u.get('REG', '?')} | {u.get('SHARED', '?')} | {total[k]} | " + " | ".join(str(c[o]) for o in OPS[:9]) +
                f" | {c['LDS']} | {c['STS']} | {c['LDG']} | {c['STG']} |\n")
print("wrote", f"profiles/{tag}_sass_summary.md", len(own), "kernels")
