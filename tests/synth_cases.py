"""Seeded synthetic parity cases shared by the golden generator, the CPU tests and the GPU tests."""
import hashlib
import os

import numpy as np

CASES = {
    # name: (db kwargs, read kwargs, seq_mode)
    "multi_se": (dict(genera=6, species_per_genus=4, strains_per_species=2, codons=3000, seed=3, eukaryote_genera=1),
                 dict(n_reads=4000, length=150, seed=4, n_rate=0.002), 1),
    "multi_pe": (dict(genera=6, species_per_genus=4, strains_per_species=2, codons=3000, seed=3, eukaryote_genera=1),
                 dict(n_reads=3000, length=150, seed=5, n_rate=0.002, paired=True), 2),
    # close species => many score ties within tie-ratio => LCA classifications above species
    "ties_se": (dict(genera=3, species_per_genus=5, strains_per_species=3, codons=2500, seed=7, species_div=0.02, strain_div=0.004),
                dict(n_reads=4000, length=151, seed=8, sub_rate=0.02), 1),
    # ragged lengths incl. reads too short for a single k-mer, many Ns
    "ragged_se": (dict(genera=4, species_per_genus=3, strains_per_species=2, codons=2000, seed=11),
                  dict(n_reads=3000, length=140, seed=12, n_rate=0.01, length_jitter=125), 1),
    # format-1 database (OldMetamerScanner: base-21 amino-acid part, reversed codon order)
    "format1_pe": (dict(genera=4, species_per_genus=4, strains_per_species=2, codons=2500, seed=17, species_div=0.05, kmer_format=1),
                   dict(n_reads=3000, length=150, seed=18, n_rate=0.002, sub_rate=0.02, paired=True), 2),
    # syncmer databases (SyncmerScanner.h: only closed syncmers are indexed and queried; paths may skip up to 8 - s codons;
    # SURVEY §8f N3 — the reference's current default DB type)
    "sync_se": (dict(genera=6, species_per_genus=4, strains_per_species=2, codons=3000, seed=23, eukaryote_genera=1, syncmer=1, smer_len=5),
                dict(n_reads=4000, length=150, seed=24, n_rate=0.002, sub_rate=0.02), 1),
    "sync_pe": (dict(genera=4, species_per_genus=4, strains_per_species=2, codons=2500, seed=27, species_div=0.05, syncmer=1, smer_len=6),
                dict(n_reads=3000, length=151, seed=28, n_rate=0.003, sub_rate=0.02, paired=True), 2),
    # long reads (seq-mode 3: denominator 1000)
    "long": (dict(genera=3, species_per_genus=3, strains_per_species=2, codons=6000, seed=13),
             dict(n_reads=300, length=6000, seed=14, sub_rate=0.05), 3),
}


# cases checked on the CPU only so far (oracle vs the reference binary's TSV); to be promoted to CASES (GPU parity) next
CPU_CASES = {
    # mates of different, partly too short lengths: a pair counts only when BOTH mates hold a k-mer (KmerExtractor.cpp:436-471)
    # non-default classify flags (thresholds, tie ratio, minimum path depth) on a database with close species and eukaryotes
    "flags_se": (dict(genera=4, species_per_genus=5, strains_per_species=2, codons=2500, seed=37, species_div=0.03, strain_div=0.006,
                      eukaryote_genera=2),
                 dict(n_reads=4000, length=150, seed=38, sub_rate=0.03, n_rate=0.001), 1),
    # accession-level databases (Accession_level 1 in db.parameters): by default the "accession" leaves are pruned from the clade
    # descent (Taxonomer.cpp:256-267, accessionLevel 2); with --accession-level 1 reads go down to the accession
    "acc_prune_se": (dict(genera=4, species_per_genus=3, strains_per_species=3, codons=2500, seed=43, strain_div=0.03, accession_leaves=True),
                     dict(n_reads=3000, length=150, seed=44, sub_rate=0.01), 1),
    "acc_lvl1_se": (dict(genera=4, species_per_genus=3, strains_per_species=3, codons=2500, seed=43, strain_div=0.03, accession_leaves=True),
                    dict(n_reads=3000, length=150, seed=44, sub_rate=0.01), 1),
    # long ragged reads (BASELINE configs[3]/[4] shape in miniature): 3-46 kbp, 8 % substitutions, seq-mode 3
    "ont_ragged": (dict(genera=3, species_per_genus=3, strains_per_species=2, codons=20000, seed=41),
                   dict(n_reads=48, length=50000, seed=42, sub_rate=0.08, n_rate=0.0005, length_jitter=47000), 3),
    # database written without Skip_redundancy (older builds): bit 31 of an info entry is a flag that classify masks away
    # (KmerMatcher.cpp:204-205, 381); the case sets it on every third entry
    "redund_se": (dict(genera=4, species_per_genus=3, strains_per_species=2, codons=2000, seed=47),
                  dict(n_reads=3000, length=150, seed=48, sub_rate=0.02), 1),
    # --lineage 1: an extra column with the lineage of the classification (Reporter.cpp:37-79, TaxonomyWrapper::taxLineage2)
    "lineage_se": (dict(genera=5, species_per_genus=3, strains_per_species=2, codons=1500, seed=53, eukaryote_genera=1),
                   dict(n_reads=1500, length=150, seed=54, sub_rate=0.02, n_rate=0.002), 1),
    "ragged_pe": (dict(genera=4, species_per_genus=3, strains_per_species=2, codons=2000, seed=31),
                  dict(n_reads=3000, length=150, seed=32, n_rate=0.004, paired=True, length_jitter=118, mate2_jitter=30), 2),
}


# classify flags of a case (reference spelling -> value); cases without an entry run the defaults
FLAGS = {
    "flags_se": {"--min-score": 0.3, "--min-sp-score": 0.6, "--tie-ratio": 0.9, "--min-cons-cnt": 6, "--min-cons-cnt-euk": 11},
    "acc_lvl1_se": {"--accession-level": 1},
    "lineage_se": {"--lineage": 1},
}


def oracle_flags(name):
    f = FLAGS.get(name, {})
    return dict(min_score=f.get("--min-score", 0.0), min_sp_score=f.get("--min-sp-score", 0.0), tie_ratio=f.get("--tie-ratio", 0.95),
                min_cons=f.get("--min-cons-cnt", 4), min_cons_euk=f.get("--min-cons-cnt-euk", 9),
                accession_level=f.get("--accession-level", 0), lineage=f.get("--lineage", 0))


def build(name):
    from metabuli_b200 import synth
    dbkw, rkw, seq_mode = (CASES.get(name) or CPU_CASES[name])
    sdb = synth.make_db(**dbkw)
    if name == "redund_se":
        sdb.database.info[::3] |= np.int32(-2147483648)           # bit 31
        sdb.database.params.skip_redundancy = 0
    reads = synth.make_reads(sdb, **rkw)
    return sdb, reads, seq_mode


def fingerprint(sdb, reads) -> str:
    h = hashlib.md5()
    h.update(np.ascontiguousarray(sdb.database.diff_idx).tobytes())
    h.update(np.ascontiguousarray(sdb.database.info).tobytes())
    h.update(sdb.taxonomy_blob)
    for a in reads:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def write_fasta(path, bases, offsets):
    with open(path, "w") as f:
        for i in range(offsets.size - 1):
            f.write(">r%d\n%s\n" % (i, bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode()))


def names(n):
    return ["r%d" % i for i in range(n)]
