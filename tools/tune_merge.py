#!/usr/bin/env python
"""Tuning harness (GPU box): builds one synthetic workload, then times the device pipeline for several tile
geometries (MBL_TILE_CELLS is read at mbl_create).  Prints one JSON line per setting and the best one last."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from metabuli_b200 import Classifier, ClassifyOptions  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--db-gib", type=float, default=1.0)
ap.add_argument("--reads", type=int, default=2_000_000)
ap.add_argument("--cells", default="1,2,4")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--env", default="", help="extra settings to sweep, e.g. MBL_DYN_CHUNKS=0,1 (applied at the first --cells value)")
args = ap.parse_args()
args.read_len = 150
sdb, reads, winfo = bench.build_workload(args, "cuda:0", seed_reads=4)
bases, offs = np.ascontiguousarray(reads[0]), np.ascontiguousarray(reads[1])
best = None
settings = [(int(x), None) for x in args.cells.split(",")]
if args.env:
    name, vals = args.env.split("=")
    settings = [(settings[0][0], (name, v)) for v in vals.split(",")] + settings[1:]
for cells, extra in settings:
    os.environ["MBL_TILE_CELLS"] = str(cells)
    if extra:
        os.environ[extra[0]] = extra[1]
    clf = Classifier(None, ClassifyOptions(seq_mode=1), database=sdb.database)
    batch, keep = clf.make_batch(bases, offs)
    assert clf.lib.mbl_upload_batch(clf.ctx, C.byref(batch)) == 0
    for _ in range(2):
        assert clf.lib.mbl_classify_resident(clf.ctx) == 0, clf.lib.mbl_last_error(clf.ctx)
    acc = {}
    for _ in range(args.steps):
        assert clf.lib.mbl_classify_resident(clf.ctx) == 0
        for k, v in clf.stats().items():
            acc[k] = acc.get(k, 0) + v
    st = {k: v / args.steps for k, v in acc.items()}
    gbs = st["merge_bytes"] / 1e9 / (st["ms_merge_kernel"] / 1e3)
    line = {"tile_cells": cells, "env": extra, "merge_kernel_ms": round(st["ms_merge_kernel"], 3), "merge_gbs": round(gbs, 1), "tiles": clf.db_info()["n_tiles"],
            "jumbo": clf.db_info()["n_jumbo"], "matches": st["n_matches"],
            "stages_ms": {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")}}
    print(json.dumps(line), flush=True)
    if best is None or line["merge_kernel_ms"] < best["merge_kernel_ms"]:
        best = line
    clf.close()
print(json.dumps({"best_tile_cells": best["tile_cells"], **winfo}))
