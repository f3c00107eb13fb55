#!/usr/bin/env python
import os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from metabuli_b200 import Classifier, ClassifyOptions
from metabuli_b200.fastx import read_fastx
fx = os.path.join(ROOT, "tests", "golden", "fixtures")
db_dir = os.path.join(fx, "db_in")
names, b1, o1 = read_fastx(os.path.join(fx, "reads", "ERR9594652_5000_1.fna.gz"))
odb = oracle.OracleDb(db_dir)
ov, oq, c1, c2 = oracle.extract(b1, o1, None, None, kmer_format=odb.kmer_format)
osv, osq = oracle.sort_kmers(ov, oq)
om = odb.match(osv, osq)
nb = osv != np.uint64(0xFFFFFFFFFFFFFFFF)
key = {int(q): i for i, q in enumerate(osq[nb])}
for cells in (1, 2, 4, 8):
    os.environ["MBL_TILE_CELLS"] = str(cells)
    clf = Classifier(db_dir, ClassifyOptions(seq_mode=1))
    gm = clf.match(osv, osq)
    pos = np.array([key[int(q)] for q in gm["qinfo"]])
    h = sorted(collections.Counter((pos // 32768).tolist()).items())
    print("cells", cells, "gpu", gm.size, "of", om.size, "info", clf.db_info() if hasattr(clf, "db_info") else "", "\n   bins", h)
    clf.close()
