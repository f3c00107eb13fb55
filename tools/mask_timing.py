"""K0 (device masking, --mask 1) against the host masker on the same reads: time and equality.  Usage: mask_timing.py [n_reads] [len]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_cases  # noqa: E402
from metabuli_b200 import Classifier, ClassifyOptions  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
length = int(sys.argv[2]) if len(sys.argv) > 2 else 150
sdb, reads, _ = synth_cases.build("mask_se")
rng = np.random.default_rng(5)
# the case's 3000 reads (tandem repeats, homopolymer stretches) tiled to n reads, every tile with fresh substitutions
base = reads[0][: 3000 * 150].reshape(3000, 150)[:, :length] if length <= 150 else None
if base is not None:
    bases = np.tile(base, (n // 3000 + 1, 1))[:n].copy()
    mut = rng.random(bases.shape) < 0.01
    bases[mut] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(mut.sum()))]
    bases = bases.reshape(-1)
else:
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n * length)].copy()
off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
out = {"n_reads": n, "length": length}
clf = Classifier(None, ClassifyOptions(seq_mode=1, mask=1), database=sdb.database)
try:
    for rep in range(2):
        t = time.time()
        clf.classify_batch(bases, off)
        out[f"classify_s_run{rep}"] = round(time.time() - t, 3)
        st = clf.stats()
        out[f"ms_mask_run{rep}"] = round(st["ms_mask"], 2)
    out["stages_ms"] = {k: round(v, 1) for k, v in st.items() if k.startswith("ms_")}
    import ctypes as C
    got = np.zeros(bases.size, dtype=np.uint8)
    clf._check(clf.lib.mbl_download_reads(clf.ctx, 1, got.ctypes.data_as(C.c_void_p), got.size))
    m = min(n, 400_000)
    t = time.time()
    want = clf.mask_reads(bases[: m * length], off[: m + 1])
    out["host_mask_s"] = round(time.time() - t, 3)
    out["host_reads"] = m
    out["host_threads"] = os.cpu_count()
    out["equal_letters"] = bool(np.array_equal(got[: m * length], want))
    out["masked_letters_device"] = int((got == ord("N")).sum())
    out["device_reads_per_s"] = round(n / (out["ms_mask_run1"] / 1e3))
    out["host_reads_per_s"] = round(m / out["host_mask_s"])
finally:
    clf.close()
print(json.dumps(out))
