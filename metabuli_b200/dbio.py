"""Readers of the on-disk database the reference writes and reads (formats unchanged):

    db.parameters   src/commons/common.cpp:88-133 (loadDbParameters), IndexCreator.cpp:1251-1272
    diffIdx         u16 stream, IndexCreator.cpp:874-892 / KmerMatcher.h:282-297
    info            int32 taxid per k-mer, KmerMatcher.cpp:381
    split           n x {u64 ADkmer, u64 diffIdxOffset, u64 infoIdxOffset}, Kmer.h:111-119
    taxonomyDB      TaxonomyWrapper.cpp:289-421 (serialize / unserialize)
    taxID_list      KmerMatcher.cpp:56-120 (loadTaxIdList)

Host-side plumbing only: everything here feeds mbl_load_db(); none of it is on the hot path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# NcbiTaxonomy.h:52-80
NCBI_RANKS = {"forma": 1, "varietas": 2, "subspecies": 3, "species": 4, "species subgroup": 5, "species group": 6,
              "subgenus": 7, "genus": 8, "subtribe": 9, "tribe": 10, "subfamily": 11, "family": 12, "superfamily": 13,
              "parvorder": 14, "infraorder": 15, "suborder": 16, "order": 17, "superorder": 18, "infraclass": 19,
              "subclass": 20, "class": 21, "superclass": 22, "subphylum": 23, "phylum": 24, "superphylum": 25,
              "subkingdom": 26, "kingdom": 27, "superkingdom": 28, "domain": 28}


@dataclass
class DbParameters:
    kmer_format: int = 1          # classify.cpp:13: default when db.parameters has no Kmer_format
    reduced_aa: int = 0
    skip_redundancy: int = 0
    syncmer: int = 0
    smer_len: int = 5
    accession_level_db: int = -1
    raw: dict = field(default_factory=dict)


def load_db_parameters(db_dir: str) -> DbParameters:
    p = DbParameters()
    path = os.path.join(db_dir, "db.parameters")
    if not os.path.exists(path):
        return p
    with open(path) as f:
        for line in f:
            tok = line.rstrip("\n").split("\t")
            if len(tok) < 2:
                continue
            k, v = tok[0], tok[1]
            p.raw[k] = v
            if k == "Reduced_alphabet":
                p.reduced_aa = int(v)
            elif k == "Skip_redundancy" and v == "1":
                p.skip_redundancy = 1
            elif k == "Syncmer" and v == "1":
                p.syncmer = 1
            elif k == "S-mer_len":
                p.smer_len = int(v)
            elif k == "Kmer_format":
                p.kmer_format = int(v)
            elif k == "Accession_level":
                p.accession_level_db = int(v)
    return p


class TaxonomyDB:
    """taxonomyDB parsed into the arrays the C-ABI takes (kept exactly as stored, quirk Q5)."""

    def __init__(self, path: str):
        blob = np.fromfile(path, dtype=np.uint8)
        self.blob = blob
        off = 0
        version = int(blob[off:off + 4].view("<i4")[0]); off += 4
        if version != 2:
            raise ValueError(f"unsupported taxonomyDB version {version}")
        flag = int(blob[off:off + 8].view("<u8")[0])
        self.internal_ids = flag == 1                      # detected by value, TaxonomyWrapper.cpp:374-382
        if self.internal_ids:
            off += 8
        self.max_nodes = int(blob[off:off + 8].view("<u8")[0]); off += 8
        self.max_taxid = int(blob[off:off + 4].view("<i4")[0]); off += 4
        n, t = self.max_nodes, self.max_taxid + 1
        node_dt = np.dtype([("id", "<i4"), ("taxId", "<i4"), ("parentTaxId", "<i4"), ("pad", "<i4"), ("rankIdx", "<u8"), ("nameIdx", "<u8")])
        nodes = blob[off:off + 32 * n].view(node_dt); off += 32 * n
        self.node_taxid = np.ascontiguousarray(nodes["taxId"])
        self.node_parent = np.ascontiguousarray(nodes["parentTaxId"])
        self.node_rank_idx = np.ascontiguousarray(nodes["rankIdx"])
        self.node_name_idx = np.ascontiguousarray(nodes["nameIdx"])

        def take(count):
            nonlocal off
            a = np.ascontiguousarray(blob[off:off + 4 * count].view("<i4"))
            off += 4 * count
            return a
        self.D = take(t)
        self.internal2org = take(t) if self.internal_ids else None
        self.E = take(2 * n)
        self.L = take(2 * n)
        self.H = take(n)
        self.M_k = max(1, (2 * n).bit_length())            # (int)flog2(2*maxNodes) + 1
        self.M = take(2 * n * self.M_k)
        byte_cap = int(blob[off:off + 8].view("<u8")[0]); off += 8
        entry_cap = int(blob[off:off + 4].view("<u4")[0]); off += 4
        self.str_count = int(blob[off:off + 4].view("<u4")[0]); off += 4
        self.str_data = bytes(blob[off:off + byte_cap]); off += byte_cap
        self.str_offsets = blob[off:off + 4 * entry_cap].view("<u4").copy(); off += 4 * entry_cap
        if off > blob.size:
            raise ValueError("taxonomyDB truncated")
        ranks = [self.string(int(i)) for i in self.node_rank_idx]
        self.node_rank = np.array([NCBI_RANKS.get(r, -1) for r in ranks], dtype=np.int8)
        self.node_prune = np.array([1 if r in ("", "accession") else 0 for r in ranks], dtype=np.uint8)
        self.node_rank_name = ranks
        self.eukaryota = 0                                  # TaxonomyWrapper.h setEukaryoteTaxID
        for i in range(n):
            if self.node_name_idx[i] != 0 and self.string(int(self.node_name_idx[i])) == "Eukaryota":
                self.eukaryota = int(self.node_taxid[i])
                break

    def string(self, idx: int) -> str:
        o = int(self.str_offsets[idx])
        e = self.str_data.index(b"\0", o)
        return self.str_data[o:e].decode("utf-8", "replace")

    def node_exists(self, taxid: int) -> bool:
        return taxid <= self.max_taxid and self.D[taxid] != -1

    def original(self, internal: int) -> int:              # TaxonomyWrapper::getOriginalTaxID
        return int(self.internal2org[internal]) if self.internal_ids else int(internal)

    def rank_of(self, taxid: int) -> str:
        return self.node_rank_name[int(self.D[taxid])]

    # ExtendedShortRanks (TaxonomyWrapper.h:9-26)
    SHORT_RANKS = {"subspecies": "ss", "species": "s", "subgenus": "sg", "genus": "g", "subfamily": "sf", "family": "f", "suborder": "so",
                   "order": "o", "subclass": "sc", "class": "c", "subphylum": "sp", "phylum": "p", "subkingdom": "sk", "kingdom": "k",
                   "superkingdom": "d", "domain": "d", "realm": "r"}

    def lineage(self, taxid: int) -> str:
        """TaxonomyWrapper::taxLineage2 (TaxonomyWrapper.cpp:431-454): short rank '_' name of every node from below the root
        down to the taxon, ';'-separated; the walk stops at the node that is its own parent."""
        chain = []
        node = int(self.D[taxid])
        while True:
            chain.append(node)
            node = int(self.D[self.node_parent[node]])
            if self.node_parent[node] == self.node_taxid[node]:
                break
        return ";".join(self.SHORT_RANKS.get(self.node_rank_name[n], "-") + "_" + self.string(int(self.node_name_idx[n])) for n in reversed(chain))

    def taxid_at_rank(self, taxid: int, rank: str) -> int:  # TaxonomyWrapper.cpp:479-498
        if taxid == 0 or not self.node_exists(taxid) or taxid == 1:
            return 0
        want = NCBI_RANKS.get(rank, -1)
        node = int(self.D[taxid])
        cnt = 0
        while cnt < 30 and self.node_rank[node] < want:
            node = int(self.D[self.node_parent[node]])
            cnt += 1
        return taxid if cnt == 30 else int(self.node_taxid[node])

    def build_taxid2species(self, taxid_list) -> np.ndarray:   # KmerMatcher.cpp:96-119
        out = np.zeros(self.max_taxid + 1, dtype=np.int32)
        for taxid in taxid_list:
            species = self.taxid_at_rank(int(taxid), "species")
            node = int(self.D[taxid])
            if taxid != self.node_taxid[node]:
                out[taxid] = species
            while self.node_taxid[node] != species:
                out[self.node_taxid[node]] = species
                node = int(self.D[self.node_parent[node]])
            out[species] = species
        return out


@dataclass
class Database:
    params: DbParameters
    diff_idx: np.ndarray
    info: np.ndarray
    split: np.ndarray
    tax: TaxonomyDB
    taxid2species: np.ndarray


def load_database(db_dir: str) -> Database:
    params = load_db_parameters(db_dir)
    diff_idx = np.fromfile(os.path.join(db_dir, "diffIdx"), dtype="<u2")
    info = np.fromfile(os.path.join(db_dir, "info"), dtype="<i4")
    split = np.fromfile(os.path.join(db_dir, "split"), dtype="<u8")
    tax = TaxonomyDB(os.path.join(db_dir, "taxonomyDB"))
    with open(os.path.join(db_dir, "taxID_list")) as f:
        ids = [int(x) for x in f.read().split() if x.strip()]
    return Database(params, diff_idx, info, split, tax, tax.build_taxid2species(ids))
