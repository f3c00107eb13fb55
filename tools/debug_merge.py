#!/usr/bin/env python
"""Debug aid: run the merge stage on the in-se fixture and show which matches differ from the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from metabuli_b200 import Classifier, ClassifyOptions
from metabuli_b200.fastx import read_fastx
fx = os.path.join(ROOT, "tests", "golden", "fixtures")
db_dir = os.path.join(fx, "db_in")
import glob
reads = sorted(glob.glob(os.path.join(fx, "reads", "*")))
print(reads)
names, b1, o1 = read_fastx(reads[0])
clf = Classifier(db_dir, ClassifyOptions(seq_mode=1))
odb = oracle.OracleDb(db_dir)
gv, gq = clf.extract(b1, o1, None, None)
sv, sq = clf.sort_kmers(gv, gq)
ov, oq, c1, c2 = oracle.extract(b1, o1, None, None, kmer_format=odb.kmer_format)
osv, osq = oracle.sort_kmers(ov, oq)
om = odb.match(osv, osq)
gm = clf.match(sv, sq)
print("oracle", om.size, "gpu", gm.size, "stats", clf.stats())
# position of each match's query in the sorted query array
nb = sv != np.uint64(0xFFFFFFFFFFFFFFFF)
key = {int(q): i for i, q in enumerate(sq[nb])}
oq_pos = np.array([key[int(q)] for q in om["qinfo"]])
gq_pos = np.array([key[int(q)] for q in gm["qinfo"]])
print("n_query", int(nb.sum()))
import collections
def hist(a, f): return sorted(collections.Counter(f(a).tolist()).items())
print("oracle by chunk32%8 ", hist(oq_pos, lambda a: (a // 32) % 8))
print("gpu    by chunk32%8 ", hist(gq_pos, lambda a: (a // 32) % 8))
print("oracle by item(32768)", hist(oq_pos, lambda a: a // 32768))
print("gpu    by item(32768)", hist(gq_pos, lambda a: a // 32768))
print("oracle by round(256) first 8", hist(oq_pos[oq_pos < 32768], lambda a: a // 4096))
print("gpu    by round(256) first 8", hist(gq_pos[gq_pos < 32768], lambda a: a // 4096))
og = set(zip(om["qinfo"].tolist(), om["target_id"].tolist() if "target_id" in om.dtype.names else om[om.dtype.names[1]].tolist(), om["dna_encoding"].tolist()))
gg = set(zip(gm["qinfo"].tolist(), gm[gm.dtype.names[1]].tolist(), gm["dna_encoding"].tolist()))
print("in both", len(og & gg), "only oracle", len(og - gg), "only gpu", len(gg - og))
