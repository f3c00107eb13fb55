"""Seeded synthetic databases and reads in the reference's on-disk format.

Test and benchmark tooling, not part of the hot path.  Everything is generated with torch tensor ops so
the same code makes a few-thousand-k-mer DB on the CPU for the parity tests and a multi-GiB index on the
GPU for bench.py (torch is plumbing here: RNG, sort, cumsum).

Model: a 5-level taxonomy root > superkingdom > genus > species > strain ("subspecies").  Every genus
has a random protein-coding ancestor of `codons` codons; species are copies with `species_div` of the
codons redrawn, strains copies of their species with `strain_div` redrawn.  The index holds every
in-frame metamer of every strain, de-duplicated per (value, species) with taxid = strain or, when several
strains of the species share it, the species (IndexCreator.h:475-629 semantics).  Files written:
diffIdx / info / split (IndexCreator.cpp:817-892), taxonomyDB (TaxonomyWrapper.cpp:289-360), taxID_list,
db.parameters (IndexCreator.cpp:1251-1272).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np
import torch

from .dbio import Database, DbParameters, TaxonomyDB

# ---- genetic code in the reference's encoding (GeneticCode.h:32-194): base code A0 C1 T2 G3 ------------------
_AA_ORDER = "ARNDCQEGHILKMFPSTWYVX"
_STD = {  # codon -> amino acid (X = stop)
    "GCA": "A", "GCC": "A", "GCT": "A", "GCG": "A", "CGA": "R", "CGC": "R", "CGT": "R", "CGG": "R", "AGG": "R", "AGA": "R",
    "AAC": "N", "AAT": "N", "GAC": "D", "GAT": "D", "TGC": "C", "TGT": "C", "CAA": "Q", "CAG": "Q", "GAA": "E", "GAG": "E",
    "GGA": "G", "GGC": "G", "GGT": "G", "GGG": "G", "CAC": "H", "CAT": "H", "ATA": "I", "ATC": "I", "ATT": "I",
    "CTA": "L", "CTC": "L", "CTT": "L", "CTG": "L", "TTG": "L", "TTA": "L", "AAA": "K", "AAG": "K", "ATG": "M",
    "TTC": "F", "TTT": "F", "CCA": "P", "CCC": "P", "CCT": "P", "CCG": "P", "TCA": "S", "TCC": "S", "TCT": "S", "TCG": "S",
    "AGT": "S", "AGC": "S", "ACA": "T", "ACC": "T", "ACT": "T", "ACG": "T", "TGG": "W", "TAC": "Y", "TAT": "Y",
    "GTA": "V", "GTC": "V", "GTT": "V", "GTG": "V", "TAA": "X", "TAG": "X", "TGA": "X"}
_BASES = "ACTG"


def _codon_tables():
    aa = np.zeros(64, dtype=np.int64)
    cid = np.zeros(64, dtype=np.int64)
    asc = np.zeros((64, 3), dtype=np.uint8)
    for a in range(4):
        for b in range(4):
            for c in range(4):
                cod = _BASES[a] + _BASES[b] + _BASES[c]
                i = a * 16 + b * 4 + c
                aa[i] = _AA_ORDER.index(_STD[cod])
                k = c
                if cod == "AGG": k = 4
                if cod == "AGA": k = 5
                if cod == "TTG": k = 4
                if cod == "TTA": k = 5
                if cod == "AGT": k = 6
                if cod == "AGC": k = 7
                if cod == "TGA": k = 5
                cid[i] = k
                asc[i] = [ord(x) for x in cod]
    return aa, cid, asc


_AA, _CID, _ASC = _codon_tables()


# ---- taxonomy ------------------------------------------------------------------------------------------------
@dataclass
class SynthTaxonomy:
    parent: np.ndarray       # per taxid (index = internal taxid), parent taxid; taxid 0 unused
    rank: list               # rank string per taxid
    name: list
    strain_ids: np.ndarray   # [n_strains]
    species_of_strain: np.ndarray
    species_ids: np.ndarray
    genus_of_species: np.ndarray


def make_taxonomy(genera: int, species_per_genus: int, strains_per_species: int, eukaryote_genera: int = 0,
                  accession_leaves: bool = False) -> SynthTaxonomy:
    """accession_leaves: every strain gets one child of rank "accession" that carries the k-mers (databases built with
    --accession-level 1, IndexCreator / Taxonomer.cpp:256-267)."""
    # internal taxid 1 must not carry k-mers (getTaxIdAtRank returns 0 for it, TaxonomyWrapper.cpp:480)
    parent, rank, name = [0, 1], ["", "no rank"], ["", "root"]

    def add(p, r, n):
        parent.append(p); rank.append(r); name.append(n)
        return len(parent) - 1
    bact = add(1, "superkingdom", "Bacteria")
    euk = add(1, "superkingdom", "Eukaryota") if eukaryote_genera else 0
    species_ids, genus_of_species, strain_ids, species_of_strain = [], [], [], []
    for g in range(genera):
        top = euk if g < eukaryote_genera else bact
        gid = add(top, "genus", f"Genus{g}")
        for s in range(species_per_genus):
            sid = add(gid, "species", f"Genus{g} species{s}")
            species_ids.append(sid); genus_of_species.append(gid)
            for t in range(strains_per_species):
                tid = add(sid, "subspecies", f"Genus{g} species{s} strain{t}")
                if accession_leaves:
                    tid = add(tid, "accession", f"ACC_{g}_{s}_{t}.1")
                strain_ids.append(tid); species_of_strain.append(sid)
    return SynthTaxonomy(np.array(parent, dtype=np.int32), rank, name, np.array(strain_ids, dtype=np.int32),
                         np.array(species_of_strain, dtype=np.int32), np.array(species_ids, dtype=np.int32),
                         np.array(genus_of_species, dtype=np.int32))


def taxonomy_db_bytes(tx: SynthTaxonomy) -> bytes:
    """Serialise in the layout TaxonomyWrapper::unserialize reads (internal ids == original ids)."""
    n_tax = len(tx.parent)                     # taxids 1..n_tax-1
    n = n_tax - 1                              # nodes; node id = taxid - 1 (root = node 0)
    max_taxid = n_tax - 1
    children = [[] for _ in range(n_tax)]
    for t in range(2, n_tax):
        children[tx.parent[t]].append(t)
    # Euler tour (iterative), E holds node ids, L levels, H first occurrence
    E, L = [], []
    H = np.zeros(n, dtype=np.int32)
    seen = np.zeros(n_tax, dtype=bool)
    stack = [(1, 0, 0)]
    while stack:
        t, lvl, ci = stack.pop()
        if not seen[t]:
            seen[t] = True
            H[t - 1] = len(E)
        E.append(t - 1); L.append(lvl)
        if ci < len(children[t]):
            stack.append((t, lvl, ci + 1))
            stack.append((children[t][ci], lvl + 1, 0))
    m = 2 * n
    E = np.array(E + [0] * (m - len(E)), dtype=np.int32)
    L = np.array(L + [1 << 30] * (m - len(L)), dtype=np.int32)
    K = max(1, m.bit_length())
    M = np.zeros((m, K), dtype=np.int32)
    M[:, 0] = np.arange(m, dtype=np.int32)
    for k in range(1, K):
        half = 1 << (k - 1)
        a = M[:, k - 1]
        b = np.concatenate([M[half:, k - 1], np.full(min(half, m), m - 1, dtype=np.int32)])[:m]
        M[:, k] = np.where(L[a] < L[b], a, b)
    # strings
    strings, index = [""], {"": 0}

    def sidx(s):
        if s not in index:
            index[s] = len(strings); strings.append(s)
        return index[s]
    nodes = np.zeros(n, dtype=np.dtype([("id", "<i4"), ("taxId", "<i4"), ("parentTaxId", "<i4"), ("pad", "<i4"), ("rankIdx", "<u8"), ("nameIdx", "<u8")]))
    for t in range(1, n_tax):
        nodes[t - 1] = (t - 1, t, tx.parent[t] if t > 1 else 1, 0, sidx(tx.rank[t]), sidx(tx.name[t]))
    D = np.full(max_taxid + 1, -1, dtype=np.int32)
    D[1:] = np.arange(n, dtype=np.int32)
    i2o = np.arange(max_taxid + 1, dtype=np.int32)
    data = b"".join(s.encode() + b"\0" for s in strings)
    offs = np.zeros(len(strings), dtype=np.uint32)
    pos = 0
    for i, s in enumerate(strings):
        offs[i] = pos; pos += len(s.encode()) + 1
    out = [struct.pack("<i", 2), struct.pack("<Q", 1), struct.pack("<Q", n), struct.pack("<i", max_taxid), nodes.tobytes(),
           D.tobytes(), i2o.tobytes(), E.tobytes(), L.tobytes(), H.tobytes(), M.tobytes(),
           struct.pack("<QII", len(data), len(strings), len(strings)), data, offs.tobytes()]
    return b"".join(out)


# ---- genomes and the index -------------------------------------------------------------------------------------
def _mutate(codons: torch.Tensor, rate: float, gen: torch.Generator) -> torch.Tensor:
    if rate <= 0:
        return codons.clone()
    mask = torch.rand(codons.shape, generator=gen, device=codons.device) < rate
    rnd = torch.randint(0, 64, codons.shape, generator=gen, device=codons.device, dtype=torch.uint8)
    return torch.where(mask, rnd, codons)


def make_genomes(tx: SynthTaxonomy, codons: int, species_div: float, strain_div: float, seed: int, device="cpu") -> torch.Tensor:
    """-> uint8 [n_strains, codons] of codon indices (a*16 + b*4 + c over A0 C1 T2 G3)."""
    gen = torch.Generator(device=device); gen.manual_seed(seed)
    genus_ids, genus_inv = np.unique(tx.genus_of_species, return_inverse=True)
    anc = torch.randint(0, 64, (len(genus_ids), codons), generator=gen, device=device, dtype=torch.uint8)
    sp = _mutate(anc[torch.as_tensor(genus_inv, device=device)], species_div, gen)
    sp_index = {int(s): i for i, s in enumerate(tx.species_ids)}
    idx = torch.as_tensor([sp_index[int(s)] for s in tx.species_of_strain], device=device)
    return _mutate(sp[idx], strain_div, gen)


def _metamers(codons: torch.Tensor, kmer_format: int = 2) -> torch.Tensor:
    """All in-frame metamers of each row: int64 bit patterns of the uint64 values [rows, codons-7].
    format 2 (KmerScanner.h:74-117): 8 x 5-bit amino acids, first codon most significant.
    format 1 (KmerScanner.h:137-181): codons are read from the far end, the amino-acid part is a base-21 number
    whose most significant digit is the LAST codon of the window; codon ids follow the same order."""
    dev = codons.device
    aa = torch.as_tensor(_AA, device=dev)[codons.long()]
    cid = torch.as_tensor(_CID, device=dev)[codons.long()]
    w = codons.shape[1] - 7
    val = torch.zeros((codons.shape[0], w), dtype=torch.int64, device=dev)
    if kmer_format == 2:
        for k in range(8):
            val |= (aa[:, k:k + w] << (24 + 5 * (7 - k))) | (cid[:, k:k + w] << (3 * (7 - k)))
    else:
        aap = torch.zeros_like(val)
        for k in range(7, -1, -1):
            aap = aap * 21 + aa[:, k:k + w]
            val |= cid[:, k:k + w] << (3 * k)
        val |= aap << 24
    return val


_TOP = -(1 << 63)  # int64 with only the sign bit set: x ^ _TOP maps unsigned order onto signed order


@dataclass
class SynthDb:
    database: Database
    tax: SynthTaxonomy
    genomes: torch.Tensor         # uint8 [n_strains, codons]
    taxonomy_blob: bytes
    taxid_list: np.ndarray
    shard_first_value: int = 0    # build_db_parts(part=k): lower bound of the value range this stream covers
    shard_is_tail: bool = True    #                        ... and whether it ends with the numerically last k-mer of the index
    range_cuts: tuple = ()

    def write(self, path: str):
        os.makedirs(path, exist_ok=True)
        self.database.diff_idx.tofile(os.path.join(path, "diffIdx"))
        self.database.info.tofile(os.path.join(path, "info"))
        self.database.split.tofile(os.path.join(path, "split"))
        with open(os.path.join(path, "taxonomyDB"), "wb") as f:
            f.write(self.taxonomy_blob)
        with open(os.path.join(path, "taxID_list"), "w") as f:
            f.write("".join(f"{int(t)}\n" for t in self.taxid_list))
        p = self.database.params
        with open(os.path.join(path, "db.parameters"), "w") as f:
            f.write("DB_name\tsynthetic\nCreation_date\t2026-1-1\nReduced_alphabet\t0\nAccession_level\t%d\nMask_mode\t0\n" % (1 if p.accession_level_db == 1 else 0) +
                    "Mask_prob\t0.900000\nSkip_redundancy\t%d\nSyncmer\t%d\n%sKmer_format\t%d\n"
                    % (p.skip_redundancy, p.syncmer, ("S-mer_len\t%d\n" % p.smer_len) if p.syncmer else "", p.kmer_format))    # IndexCreator.cpp:1258-1270


def encode_index(values_i64: torch.Tensor, split_num: int = 4096, prev: int = 0):
    """values: sorted uint64 bit patterns (int64 tensor); prev: the value the first delta is relative to (0 at the start of a
    stream).  -> (diffIdx u16 numpy, split u64 numpy[split_num*3])"""
    dev = values_i64.device
    n = values_i64.numel()
    prev0 = torch.tensor([prev if prev < (1 << 63) else prev - (1 << 64)], dtype=torch.int64, device=dev)
    prev = torch.cat([prev0, values_i64[:-1]])
    d = values_i64 - prev                      # wraps; only d[0] can exceed 2^63 (as unsigned)
    # number of 15-bit fragments; treat d as unsigned
    neg = d < 0
    nfrag = torch.ones(n, dtype=torch.int64, device=dev)
    for j in range(1, 5):
        nfrag += ((d >> (15 * j)) != 0) & ~neg if j < 5 else 0
    nfrag = torch.where(neg, torch.full_like(nfrag, 5), nfrag)
    end = torch.cumsum(nfrag, 0)               # exclusive end offset of each k-mer in the u16 stream
    total = int(end[-1].item()) if n else 0
    out = torch.zeros(total, dtype=torch.int16, device=dev)
    for j in range(5):
        m = nfrag > j
        frag = (d[m] >> (15 * j)) & 0x7FFF
        if j == 4:
            frag = (d[m] >> 60) & 0xF          # arithmetic shift would smear the sign; keep the 4 real bits
        if j == 0:
            frag = frag | 0x8000
        frag = torch.where(frag >= 0x8000, frag - 0x10000, frag)
        out[end[m] - 1 - j] = frag.to(torch.int16)
    diff = out.cpu().numpy().view(np.uint16)
    # split checkpoints (IndexCreator.cpp:817-872)
    split = np.zeros(split_num * 3, dtype=np.uint64)
    size = n // (split_num - 1) if split_num > 1 else 0
    if size > 0:
        aa = values_i64 >> 24                  # arithmetic shift keeps equality semantics of the AA part
        starts = torch.nonzero(torch.cat([torch.ones(1, dtype=torch.bool, device=dev), aa[1:] != aa[:-1]])).flatten()
        recs = []
        for os_ in range(1, split_num):
            t = os_ * size - 1                 # k-mer after which the checkpoint is armed
            if t >= n:
                break
            pos = int(torch.searchsorted(starts, torch.tensor([t], device=dev), right=True).item())
            if pos >= starts.numel():
                break
            recs.append(int(starts[pos].item()))
        recs = sorted(set(recs))
        for i, j in enumerate(recs, start=1):
            v = int(values_i64[j].item()) & 0xFFFFFFFFFFFFFFFF
            split[3 * i:3 * i + 3] = (v, int(end[j].item()), j + 1)
    return diff, split


def syncmer_mask(values_i64: torch.Tensor, smer_len: int) -> torch.Tensor:
    """Closed-syncmer predicate of format-2 metamers (SyncmerScanner.h:36-74): the smallest of the 8 - s + 1 s-mers of the
    8-residue window (leftmost on ties) sits at the first or at the last s-mer position."""
    aa = (values_i64 >> 24) & ((1 << 40) - 1)
    n = 8 - smer_len + 1
    mask = (1 << (5 * smer_len)) - 1
    sm = [(aa >> (5 * (n - 1 - j))) & mask for j in range(n)]
    first = torch.ones_like(aa, dtype=torch.bool)
    for j in range(1, n):
        first &= sm[0] <= sm[j]
    last = torch.ones_like(aa, dtype=torch.bool)
    for j in range(n - 1):
        last &= sm[n - 1] < sm[j]
    return first | last


def collect_range(tx: SynthTaxonomy, genomes: torch.Tensor, lo: int | None = None, hi: int | None = None, chunk_rows: int = 4096,
                  kmer_format: int = 2, syncmer: int = 0, smer_len: int = 5):
    """The index entries whose value lies in [lo, hi) (unsigned; None = open end), sorted by value: -> (values int64 bit patterns,
    info int32).  One entry per (value, species) with taxid = the strain, or the species when several strains share the k-mer."""
    dev = genomes.device
    strain_t = torch.as_tensor(tx.strain_ids.astype(np.int64), device=dev)
    species_t = torch.as_tensor(tx.species_of_strain.astype(np.int64), device=dev)
    lo_s = None if lo is None else torch.tensor(lo if lo < (1 << 63) else lo - (1 << 64), dtype=torch.int64, device=dev) ^ _TOP
    hi_s = None if hi is None else torch.tensor(hi if hi < (1 << 63) else hi - (1 << 64), dtype=torch.int64, device=dev) ^ _TOP
    vals, tids, sps = [], [], []
    chunk_rows = max(8, min(chunk_rows, (1 << 28) // max(1, genomes.shape[1])))      # <= 2 GiB of int64 metamers per step
    for r0 in range(0, genomes.shape[0], chunk_rows):
        g = genomes[r0:r0 + chunk_rows]
        v = _metamers(g, kmer_format)
        t_ = strain_t[r0:r0 + chunk_rows, None].expand_as(v)
        s_ = species_t[r0:r0 + chunk_rows, None].expand_as(v)
        keep = None
        if syncmer:                                  # IndexCreator.cpp:940 / 1052: the index holds syncmers only
            keep = syncmer_mask(v, smer_len)
        if lo_s is not None or hi_s is not None:
            sv = v ^ _TOP                            # unsigned order as signed order
            rng = torch.ones_like(v, dtype=torch.bool)
            if lo_s is not None:
                rng &= sv >= lo_s
            if hi_s is not None:
                rng &= sv < hi_s
            keep = rng if keep is None else (keep & rng)
        if keep is not None:
            vals.append(v[keep]); tids.append(t_[keep]); sps.append(s_[keep])
            continue
        vals.append(v.flatten())
        tids.append(t_.flatten())
        sps.append(s_.flatten())
    val = torch.cat(vals); tid = torch.cat(tids); sp = torch.cat(sps)
    del vals, tids, sps
    if val.numel() == 0:
        return val, torch.zeros(0, dtype=torch.int32, device=dev)
    # order by (value unsigned, species, taxid): three stable sorts, least significant key first
    o = torch.sort(tid, stable=True).indices
    val, tid, sp = val[o], tid[o], sp[o]
    o = torch.sort(sp, stable=True).indices
    val, tid, sp = val[o], tid[o], sp[o]
    o = torch.sort(val ^ _TOP, stable=True).indices
    val, tid, sp = val[o], tid[o], sp[o]
    del o
    # one entry per (value, species); taxid = the strain, or the species when several strains share it
    first = torch.ones(val.numel(), dtype=torch.bool, device=dev)
    first[1:] = (val[1:] != val[:-1]) | (sp[1:] != sp[:-1])
    grp = torch.cumsum(first.long(), 0) - 1
    ng = int(grp[-1].item()) + 1
    tmin = torch.full((ng,), 1 << 40, dtype=torch.int64, device=dev).scatter_reduce(0, grp, tid, "amin")
    tmax = torch.zeros(ng, dtype=torch.int64, device=dev).scatter_reduce(0, grp, tid, "amax")
    uval = val[first]
    usp = sp[first]
    info = torch.where(tmin == tmax, tmin, usp).to(torch.int32)
    return uval, info


def range_cuts(genomes: torch.Tensor, n_parts: int, kmer_format: int = 2, sample_rows: int = 64) -> list:
    """n_parts - 1 ascending cut values (unsigned ints with the DNA bits cleared, i.e. amino-acid-group aligned) that split the
    metamers of these genomes into ranges of about equal size — quantiles of a sample of rows."""
    if n_parts <= 1:
        return []
    rows = torch.linspace(0, genomes.shape[0] - 1, min(sample_rows, genomes.shape[0])).long().to(genomes.device)
    v = torch.sort(_metamers(genomes[rows], kmer_format).flatten() ^ _TOP).values ^ _TOP
    cuts = []
    for k in range(1, n_parts):
        x = int(v[(v.numel() * k) // n_parts].item()) & 0xFFFFFFFFFFFFFFFF
        x &= ~0xFFFFFF
        if cuts and x <= cuts[-1]:
            continue
        cuts.append(x)
    return cuts


def _finish_db(tx, genomes, diff, info_np, split, kmer_format, syncmer, smer_len) -> SynthDb:
    blob = taxonomy_db_bytes(tx)
    tmp = "/tmp/_mbl_synth_tax_%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(blob)
    taxdb = TaxonomyDB(tmp)
    os.remove(tmp)
    taxid_list = np.concatenate([tx.strain_ids, tx.species_ids]).astype(np.int32)
    params = DbParameters(kmer_format=kmer_format, skip_redundancy=1, syncmer=1 if syncmer else 0, smer_len=smer_len)
    db = Database(params, diff, info_np, split, taxdb, taxdb.build_taxid2species(taxid_list))
    return SynthDb(db, tx, genomes, blob, taxid_list)


def build_db(tx: SynthTaxonomy, genomes: torch.Tensor, split_num: int = 4096, chunk_rows: int = 4096, kmer_format: int = 2,
             syncmer: int = 0, smer_len: int = 5) -> SynthDb:
    uval, info = collect_range(tx, genomes, None, None, chunk_rows, kmer_format, syncmer, smer_len)
    diff, split = encode_index(uval, split_num)
    return _finish_db(tx, genomes, diff, info.cpu().numpy(), split, kmer_format, syncmer, smer_len)


def build_db_parts(tx: SynthTaxonomy, genomes: torch.Tensor, n_parts: int, part: int | None = None, kmer_format: int = 2,
                   chunk_rows: int = 4096) -> SynthDb:
    """The same index as build_db, generated one value range at a time so that the generator's working set (three int64 per
    metamer before de-duplication, plus sort scratch) stays at 1 / n_parts of a whole-index build — the 40 GiB configurations.
    part = None: all ranges, concatenated on the host (split checkpoints are left empty: the shard planner then scans the stream).
    part = k or (k0, k1): ONLY those ranges, encoded as a stream of its own whose first delta is relative to 0 — the shard of one
    rank in the index-sharded benchmark (SynthDb.shard_first_value / shard_is_tail tell mbl_load_db_shard where it sits;
    SynthDb.range_cuts are the cut values of all ranges, identical on every rank)."""
    cuts = range_cuts(genomes, n_parts, kmer_format)
    bounds = [None] + cuts + [None]
    n_ranges = len(bounds) - 1
    diffs, infos = [], []
    prev = 0
    first_value = 0
    if part is None:
        p_lo, p_hi = 0, n_ranges
    elif isinstance(part, tuple):
        p_lo, p_hi = min(part[0], n_ranges), min(part[1], n_ranges)
    else:
        p_lo, p_hi = min(part, n_ranges), min(part + 1, n_ranges)
    for k in range(p_lo, p_hi):
        uval, info = collect_range(tx, genomes, bounds[k], bounds[k + 1], chunk_rows, kmer_format)
        if uval.numel() == 0:
            continue
        d, _ = encode_index(uval, 1, prev=prev)
        if not diffs:
            first_value = int(uval[0].item()) & 0xFFFFFFFFFFFFFFFF
        prev = int(uval[-1].item()) & 0xFFFFFFFFFFFFFFFF
        diffs.append(d); infos.append(info.cpu().numpy())
        del uval, info
        if genomes.device.type == "cuda":
            torch.cuda.empty_cache()
    diff = np.concatenate(diffs) if diffs else np.zeros(0, dtype=np.uint16)
    info_np = np.concatenate(infos) if infos else np.zeros(0, dtype=np.int32)
    sdb = _finish_db(tx, genomes, diff, info_np, np.zeros(3, dtype=np.uint64), kmer_format, 0, 5)
    sdb.shard_first_value = (bounds[p_lo] or 0) if p_lo < n_ranges else 0xFFFFFFFFFFFFFFFF
    sdb.shard_is_tail = p_hi >= n_ranges
    sdb.range_cuts = tuple(cuts)
    return sdb


def make_db(genera=4, species_per_genus=3, strains_per_species=2, codons=2000, species_div=0.12, strain_div=0.01, seed=3,
            eukaryote_genera=0, device="cpu", split_num=4096, kmer_format=2, syncmer=0, smer_len=5, accession_leaves=False,
            parts=1, part=None) -> SynthDb:
    """parts > 1: build the index range by range (build_db_parts); part = k: only range k (one rank's shard)."""
    tx = make_taxonomy(genera, species_per_genus, strains_per_species, eukaryote_genera, accession_leaves)
    genomes = make_genomes(tx, codons, species_div, strain_div, seed, device)
    if parts > 1:
        return build_db_parts(tx, genomes, parts, part, kmer_format=kmer_format)
    sdb = build_db(tx, genomes, split_num, kmer_format=kmer_format, syncmer=syncmer, smer_len=smer_len)
    if accession_leaves:
        sdb.database.params.accession_level_db = 1
    return sdb


def make_long_reads(sdb: SynthDb, n_reads: int, mean_len: float = 8000.0, sigma: float = 0.6, min_len: int = 200, max_len: int = 30000,
                    seed: int = 4, random_frac: float = 0.1, sub_rate: float = 0.08, chunk: int = 8192):
    """ONT-like reads (BASELINE configs[3]): log-normal lengths with the given mean, clipped to [min_len, max_len], substitutions
    only.  Generated in chunks (a chunk is a [reads, max_len] matrix on the device) -> (bases u8, offsets u64) numpy SoA."""
    rng = np.random.default_rng(seed)
    mu = np.log(mean_len) - 0.5 * sigma * sigma
    lens_all = np.clip(np.exp(rng.normal(mu, sigma, n_reads)), min_len, max_len).astype(np.int64)
    bases, offs, total = [], [np.zeros(1, dtype=np.uint64)], 0
    for c0 in range(0, n_reads, chunk):
        lens = lens_all[c0:c0 + chunk]
        L = int(lens.max())
        b, o = make_reads(sdb, len(lens), L, seed=seed + 1 + c0, random_frac=random_frac, sub_rate=sub_rate)
        mat = b.reshape(len(lens), L)
        keep = np.arange(L)[None, :] < lens[:, None]
        bases.append(mat[keep])
        offs.append((np.cumsum(lens) + total).astype(np.uint64))
        total += int(lens.sum())
    return np.concatenate(bases), np.concatenate(offs)


# ---- reads ---------------------------------------------------------------------------------------------------------
def make_reads(sdb: SynthDb, n_reads: int, length: int = 150, seed: int = 4, random_frac: float = 0.3, sub_rate: float = 0.01,
               n_rate: float = 0.0, paired: bool = False, insert: int = 350, length_jitter: int = 0, mate2_jitter: int = 0):
    """Reads drawn from the strain genomes (either strand, any phase) with substitutions, plus random reads.
    -> (bases1 u8, offsets1 u64[, bases2, offsets2]) as numpy arrays (SoA of mbl_batch)."""
    dev = sdb.genomes.device
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    asc = torch.as_tensor(_ASC, device=dev)
    n_strain, codons = sdb.genomes.shape
    glen = codons * 3
    frag = max(length, insert) if paired else length
    strain = torch.randint(0, n_strain, (n_reads,), generator=gen, device=dev)
    start = torch.randint(0, max(1, glen - frag), (n_reads,), generator=gen, device=dev)
    pos = start[:, None] + torch.arange(frag, device=dev)[None, :]
    pos = pos.clamp_(max=glen - 1)
    cod = sdb.genomes[strain[:, None], pos // 3].long()
    seq = asc[cod, pos % 3]                                         # uint8 ASCII [n, frag]
    rnd = torch.randint(0, 4, (n_reads, frag), generator=gen, device=dev)
    letters = torch.as_tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    is_rand = torch.rand(n_reads, generator=gen, device=dev) < random_frac
    sub = torch.rand((n_reads, frag), generator=gen, device=dev) < sub_rate
    seq = torch.where(sub | is_rand[:, None], letters[rnd], seq)
    if n_rate > 0:
        seq = torch.where(torch.rand((n_reads, frag), generator=gen, device=dev) < n_rate, torch.full_like(seq, 78), seq)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    rc = comp[seq.flip(1).long()]
    flip = torch.rand(n_reads, generator=gen, device=dev) < 0.5
    frag_seq = torch.where(flip[:, None], rc, seq)

    def pack(mat, lens):
        lens_np = lens.cpu().numpy().astype(np.uint64)
        off = np.zeros(n_reads + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens_np)
        if int(lens.min()) == mat.shape[1]:
            return mat.contiguous().flatten().cpu().numpy(), off
        keep = torch.arange(mat.shape[1], device=dev)[None, :] < lens[:, None]
        return mat[keep].cpu().numpy(), off

    if length_jitter > 0:
        lens = (length - torch.randint(0, length_jitter + 1, (n_reads,), generator=gen, device=dev)).clamp_(min=1)
    else:
        lens = torch.full((n_reads,), length, device=dev, dtype=torch.int64)
    b1, o1 = pack(frag_seq[:, :length], lens)
    if not paired:
        return b1, o1
    mate2 = comp[frag_seq.flip(1).long()][:, :length]
    lens2 = lens
    if mate2_jitter > 0:                 # mates of different lengths (drawn last, so every other output is unchanged)
        lens2 = (lens - torch.randint(0, mate2_jitter + 1, (n_reads,), generator=gen, device=dev)).clamp_(min=1)
    b2, o2 = pack(mate2, lens2)
    return b1, o1, b2, o2
