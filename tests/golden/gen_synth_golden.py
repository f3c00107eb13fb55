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
if not only or "mask_misc" in only:
    # the masking alone on odd inputs, at three thresholds: the reference's objects behind oracle/ref_mask_main.cpp
    b, o = synth_cases.mask_misc_reads()
    lines = b"".join(bytes(b[int(o[i]):int(o[i + 1])]) + b"\n" for i in range(o.size - 1))
    blob = b""
    for prob in synth_cases.MASK_MISC_PROBS:
        m = subprocess.run([os.path.join(os.path.dirname(binary), "ref_mask"), str(prob)], input=lines, capture_output=True)
        assert m.returncode == 0 and m.stdout.count(b"\n") == o.size - 1, m.stderr[-500:]
        blob += m.stdout
        print("mask_misc", prob, "masked letters", m.stdout.count(b"N") - lines.count(b"N"))
    with open(os.path.join(out_dir, "mask_misc.masked.gz"), "wb") as f:
        f.write(gzip.compress(blob, 9, mtime=0))
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
    # --mask 1 runs on one thread: with more, the reference's extraction tasks all write the one maskedSeq buffer of the thread
    # that spawned them (KmerExtractor.cpp:118-166, the pointer is firstprivate in the task) and the output is not deterministic
    args += [db_dir, work, name, "--threads", "1" if synth_cases.mask_flags(name)[0] else "4", "--max-ram", "8"]
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
    if synth_cases.mask_flags(name)[0] and seq_mode != 2:
        # per-letter golden of the masking itself: the reference's matrix and tantan objects (oracle/ref_mask_main.cpp)
        b, o = reads[0], reads[1]
        lines = b"".join(bytes(b[int(o[i]):int(o[i + 1])]) + b"\n" for i in range(o.size - 1))
        m = subprocess.run([os.path.join(os.path.dirname(binary), "ref_mask"), str(synth_cases.mask_flags(name)[1])], input=lines, capture_output=True)
        assert m.returncode == 0 and m.stdout.count(b"\n") == o.size - 1
        with open(os.path.join(out_dir, name + ".masked.gz"), "wb") as f:
            f.write(gzip.compress(m.stdout, 9, mtime=0))
        print(name, "masked letters", m.stdout.count(b"N") - lines.count(b"N"))
    with open(os.path.join(out_dir, name + ".md5"), "w") as f:
        f.write(synth_cases.fingerprint(sdb, reads) + "\n")
    stats = [l for l in r.stdout.split("\n") if "match count" in l or "k-mer number" in l]
    ranks = {}
    for ln in tsv.decode().split("\n")[1:]:
        if ln:
            c = ln.split("\t")
            ranks[c[5]] = ranks.get(c[5], 0) + 1
    print(name, stats[:2], ranks)
