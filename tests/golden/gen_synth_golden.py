"""Runs the reference binary on the seeded synthetic cases (tests/synth_cases.py) and stores its TSVs.
Called by gen_golden.sh; container-only."""
import gzip
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import synth_cases  # noqa: E402

binary, work = sys.argv[1], sys.argv[2]
out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth")
os.makedirs(out_dir, exist_ok=True)
only = set(sys.argv[3:])                                   # optional: case names to (re)generate
for name in list(synth_cases.CASES) + list(synth_cases.CPU_CASES) + list(synth_cases.EDGE_CASES):
    if only and name not in only:
        continue
    sdb, reads, seq_mode = synth_cases.build(name)
    db_dir = os.path.join(work, "db_" + name)
    sdb.write(db_dir)
    q1 = os.path.join(work, name + "_1.fna")
    synth_cases.write_fasta(q1, reads[0], reads[1])
    args = [binary, "classify"]
    if seq_mode == 2:
        q2 = os.path.join(work, name + "_2.fna")
        synth_cases.write_fasta(q2, reads[2], reads[3])
        args += [q1, q2]
    else:
        args += ["--seq-mode", str(seq_mode), q1]
    args += [db_dir, work, name, "--threads", "4", "--max-ram", "8"]
    for k, v in synth_cases.FLAGS.get(name, {}).items():
        args += [k, str(v)]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    tsv = open(os.path.join(work, name + "_classifications.tsv"), "rb").read()
    with open(os.path.join(out_dir, name + ".tsv.gz"), "wb") as f:
        f.write(gzip.compress(tsv, 9, mtime=0))
    rep = open(os.path.join(work, name + "_report.tsv"), "rb").read()          # Reporter::writeReportFile
    with open(os.path.join(out_dir, name + ".report.gz"), "wb") as f:
        f.write(gzip.compress(rep, 9, mtime=0))
    with open(os.path.join(out_dir, name + ".md5"), "w") as f:
        f.write(synth_cases.fingerprint(sdb, reads) + "\n")
    stats = [l for l in r.stdout.split("\n") if "match count" in l or "k-mer number" in l]
    ranks = {}
    for ln in tsv.decode().split("\n")[1:]:
        if ln:
            c = ln.split("\t")
            ranks[c[5]] = ranks.get(c[5], 0) + 1
    print(name, stats[:2], ranks)
