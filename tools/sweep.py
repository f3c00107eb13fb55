#!/usr/bin/env python
"""BASELINE configs[4] (read-length sweep 100 bp - 50 kbp at a fixed base count, 1 x B200) and configs[3] in single-GPU form
(ONT-like log-normal reads, mean 8 kbp): reads/s, bases/s, stage split and the merge kernel's algorithmic GB/s per length, the
database loaded once.  Writes one JSON document (default gpurun_out/r02_sweep.json); tools/sweep_md.py turns it into
profiles/r02_sweep.md.   usage: tools/sweep.py [--db-gib 8] [--gbp 1.5] [--lens 100,150,...] [--ont-reads 1000000] [--out PATH]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def run_case(clf, reads, steps, warmup):
    lib = clf.lib
    bases, offs = np.ascontiguousarray(reads[0]), np.ascontiguousarray(reads[1])
    for a in (bases, offs):
        lib.mbl_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes)
    batch, keep = clf.make_batch(bases, offs)
    rc = lib.mbl_upload_batch(clf.ctx, C.byref(batch))
    assert rc == 0, lib.mbl_last_error(clf.ctx)
    import torch
    for _ in range(warmup):
        assert lib.mbl_classify_resident(clf.ctx) == 0, lib.mbl_last_error(clf.ctx)
    torch.cuda.synchronize()
    acc, t0 = {}, time.perf_counter()
    for _ in range(steps):
        assert lib.mbl_classify_resident(clf.ctx) == 0, lib.mbl_last_error(clf.ctx)
        for k, v in clf.stats().items():
            acc[k] = acc.get(k, 0) + v
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    clf._n_resident = offs.size - 1
    res, pairs = clf.download_results()
    for a in (bases, offs):
        lib.mbl_host_unregister(a.ctypes.data_as(C.c_void_p))
    n = offs.size - 1
    peak, _ = bench.measured_peak_gbs()
    gbs = (acc["merge_bytes"] / 1e9) / (acc["ms_merge_kernel"] / 1e3) if acc["ms_merge_kernel"] > 0 else 0.0
    return {"reads": int(n), "bases": int(offs[-1]), "ms_per_step": 1000 * dt, "reads_per_s": n / dt, "gbp_per_s": float(offs[-1]) / dt / 1e9,
            "classified": int(res["is_classified"].sum()), "sub_batches": int(acc["sub_batches"] / steps), "merge_launches_per_step": acc["merge_launches"] / steps,
            "matches_per_step": acc["n_matches"] / steps, "merge_queries_per_step": acc["n_merge_queries"] / steps,
            "merge_ms_per_step": acc["ms_merge_kernel"] / steps, "merge_gbs": gbs, "merge_frac_of_peak": gbs / peak,
            "stages_ms": {k[3:]: round(v / steps, 1) for k, v in acc.items() if k.startswith("ms_") and k not in ("ms_h2d", "ms_d2h")}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db-gib", type=float, default=8.0)
    ap.add_argument("--gbp", type=float, default=1.5)
    ap.add_argument("--lens", default="100,150,250,500,1000,2000,5000,10000,25000,50000")
    ap.add_argument("--ont-reads", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_sweep.json"))
    args = ap.parse_args()
    import torch
    from metabuli_b200 import Classifier, ClassifyOptions, synth
    args.reads = 1000
    sdb, _, winfo = bench.build_workload(args, "cuda:0", seed_reads=4)
    doc = {"index": winfo, "gbp_per_case": args.gbp, "sweep": [], "ont": None}
    clfs = {}

    def clf_for(mode):
        if mode not in clfs:
            for c in clfs.values():
                c.close()
            clfs.clear()
            clfs[mode] = Classifier(None, ClassifyOptions(seq_mode=mode, device=0), database=sdb.database)
        return clfs[mode]

    # every read set is generated first (the generator works in multi-GB torch temporaries); the library's workspace then has
    # the device to itself
    cases = []
    for L in [int(x) for x in args.lens.split(",") if x]:
        n = max(1, int(args.gbp * 1e9 / L))
        cases.append((L, synth.make_reads(sdb, n, L, seed=500 + L, random_frac=0.3, sub_rate=0.01 if L <= 500 else 0.05)))
        torch.cuda.empty_cache()
    ont = None
    if args.ont_reads > 0:
        t0 = time.time()
        ont = synth.make_long_reads(sdb, args.ont_reads, seed=900)
        ont_gen_s = round(time.time() - t0, 1)
    sdb.genomes = None
    torch.cuda.empty_cache()
    for L, reads in cases:
        r = run_case(clf_for(1 if L <= 500 else 3), reads, args.steps, args.warmup)
        r["read_len"] = L
        doc["sweep"].append(r)
        print(json.dumps(r), file=sys.stderr, flush=True)
    del cases
    if ont is not None:
        r = run_case(clf_for(3), ont, max(1, args.steps - 1), args.warmup)
        r["read_len"] = "log-normal, mean %.0f bp (sigma 0.6, 200 bp - 30 kbp), 8 %% substitutions" % (r["bases"] / r["reads"])
        r["reads_gen_s"] = ont_gen_s
        doc["ont"] = r
        print(json.dumps(r), file=sys.stderr, flush=True)
    for c in clfs.values():
        c.close()
    json.dump(doc, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
