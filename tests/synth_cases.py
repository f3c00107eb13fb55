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


# Hand-made genomes that push the index decoder and the merge kernel off their common paths (code no test reached).
# Built by _edge_db below: amino-acid stretches with independently drawn synonymous codons per strain give ONE amino-acid
# group with thousands of distinct DNA parts; genomes restricted to two far-apart residue sets give 5-fragment deltas.
EDGE_CASES = {
    # poly-Leu (6 codons) stretch of 1000 codons x 8 strains: an amino-acid group of ~7.9 k k-mers, larger than a shared-memory
    # tile (2048 k-mers) => a jumbo tile pre-decoded in HBM; a second, 200-codon poly-Ser stretch gives a 1.5 k group that stays
    # in shared memory but expands to more (query, candidate) pairs than one pair window of the merge kernel holds
    "jumbo_se": (dict(codons=2600, seed=61, stretches=(("L", 300, 1000), ("S", 1700, 200))),
                 dict(n_reads=600, length=150, seed=62, sub_rate=0.01, random_frac=0.1), 1),
    # every codon codes for one of {A, R} in the first half of a genome and one of {T, W, Y, V, stop} in the second: the first
    # k-mer of the stream needs five 15-bit fragments (value >= 2^60) and so does the jump between the two populations
    "fivefrag_se": (dict(codons=2400, seed=67, residue_sets=(("T", "W", "Y", "V", "X"), ("A", "R")), species_div=0.1),
                    dict(n_reads=2000, length=150, seed=68, sub_rate=0.02), 1),
    # only {T, W, Y, V, stop}: every value is >= 2^63, so the very first k-mer of the stream takes five fragments
    "fivefrag_first_se": (dict(codons=2400, seed=71, residue_sets=(("T", "W", "Y", "V", "X"),), species_div=0.1),
                          dict(n_reads=2000, length=150, seed=72, sub_rate=0.02), 1),
    # --mask 1 (tantan masking of the queries, KmerExtractor.cpp:308-314): genomes with low-complexity DNA — amino-acid
    # homopolymers (period 3 with synonymous noise) and exact tandem repeats of 1-14 codons shared by all genomes — so that
    # masking removes k-mers that would match; reads also carry injected tandem repeats, lower-case and IUPAC letters
    "mask_se": (dict(codons=2600, seed=73, stretches=(("L", 200, 150), ("S", 900, 120), ("G", 1500, 90)),
                     tandems=((400, 120, 1), (700, 150, 2), (1100, 160, 5), (1700, 200, 9), (2100, 210, 14))),
                dict(n_reads=3000, length=150, seed=74, sub_rate=0.01, n_rate=0.001), 1),
    "mask_pe": (dict(codons=2600, seed=75, stretches=(("L", 200, 150), ("S", 900, 120)),
                     tandems=((400, 120, 1), (700, 150, 3), (1100, 160, 6), (1700, 200, 11))),
                dict(n_reads=2000, length=150, seed=76, sub_rate=0.01, paired=True, length_jitter=60), 2),
}


def _edge_db(codons, seed, stretches=(), residue_sets=(), tandems=(), species_div=0.12, strain_div=0.01):
    import torch
    from metabuli_b200 import synth
    tx = synth.make_taxonomy(2, 2, 2)
    genomes = synth.make_genomes(tx, codons, species_div, strain_div, seed)
    gen = torch.Generator(); gen.manual_seed(seed + 1000)
    aa_of = torch.as_tensor(synth._AA)

    def codons_of(letters):
        want = torch.as_tensor([synth._AA_ORDER.index(x) for x in letters])
        return torch.nonzero((aa_of[:, None] == want[None, :]).any(1)).flatten().to(torch.uint8)
    if residue_sets:
        # position-wise residue classes: block b of the genome only uses codons of residue_sets[b]; redraw every codon inside its
        # class, keeping the genus / species / strain relatedness (same random choice where the genomes agreed)
        n_blocks = len(residue_sets)
        edges = [codons * b // n_blocks for b in range(n_blocks + 1)]
        for b, letters in enumerate(residue_sets):
            pool = codons_of(letters)
            blk = genomes[:, edges[b]:edges[b + 1]].long()
            genomes[:, edges[b]:edges[b + 1]] = pool[blk % pool.numel()]
    for letter, start, length in stretches:
        pool = codons_of(letter)
        pick = torch.randint(0, pool.numel(), (genomes.shape[0], length), generator=gen)
        genomes[:, start:start + length] = pool[pick]
    for start, length, period in tandems:
        unit = torch.randint(0, 61, (period,), generator=gen).to(torch.uint8)
        unit = torch.as_tensor([c for c in range(64) if synth._AA[c] != synth._AA_ORDER.index("X")], dtype=torch.uint8)[unit.long()]   # no stop codons
        genomes[:, start:start + length] = unit[torch.arange(length) % period]
    return synth.build_db(tx, genomes)


def _inject_read_repeats(reads, seed):
    """Every fifth read gets a window overwritten by a tandem repeat (period 1..40, 0-8 % substitutions); a few letters become
    lower-case or IUPAC codes (masking must keep them as they are unless masked; NucleotideMatrix.cpp:18-57 maps them)."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for k in range(0, len(reads), 2):
        bases, offsets = reads[k].copy(), reads[k + 1]
        for r in range(0, offsets.size - 1, 5):
            b0, b1 = int(offsets[r]), int(offsets[r + 1])
            if b1 - b0 < 30:
                continue
            a = int(rng.integers(0, b1 - b0 - 20))
            ln = int(rng.integers(12, b1 - b0 - a + 1))
            unit = letters[rng.integers(0, 4, int(rng.integers(1, 41)))]
            rep = np.resize(unit, ln)
            mut = rng.random(ln) < rng.choice([0.0, 0.03, 0.08])
            bases[b0 + a:b0 + a + ln] = np.where(mut, letters[rng.integers(0, 4, ln)], rep)
        odd = rng.integers(0, bases.size, bases.size // 400)
        bases[odd] = np.frombuffer(b"acgtRYKMSWn", dtype=np.uint8)[rng.integers(0, 11, odd.size)]
        out += [bases, offsets]
    return tuple(out)


def mask_misc_reads():
    """Odd inputs for the masking alone (no database): empty and one-letter reads, reads of Ns / lower case / IUPAC codes, exact
    and noisy tandem repeats with periods around the 50-offset limit, and multi-kilobase reads (the forward and backward sweeps
    rescale every 16 letters, so long reads walk through thousands of rescalings)."""
    rng = np.random.default_rng(97)
    L = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [b"", b"A", b"N", b"ac", b"NNNNNNNNNNNNNNNNNNNN", b"A" * 15, b"A" * 16, b"A" * 17, b"ACGT" * 12 + b"T", b"acgtnRYKM*-x" * 9]
    for per in (1, 2, 3, 7, 16, 33, 49, 50, 51, 64):
        unit = L[rng.integers(0, 4, per)]
        for noise in (0.0, 0.05):
            n = int(rng.integers(60, 500))
            rep = np.resize(unit, n)
            seqs.append(bytes(np.where(rng.random(n) < noise, L[rng.integers(0, 4, n)], rep)))
    for n in (1000, 1001, 4097, 20000):
        s_ = L[rng.integers(0, 4, n)]
        for _ in range(n // 400):
            a = int(rng.integers(0, n - 50)); b = min(n, a + int(rng.integers(10, 600)))
            s_[a:b] = np.resize(L[rng.integers(0, 4, int(rng.integers(1, 56)))], b - a)
        seqs.append(bytes(s_))
    for _ in range(300):
        seqs.append(bytes(L[rng.integers(0, 4, int(rng.integers(1, 300)))]))
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(x) for x in seqs])
    return np.frombuffer(b"".join(seqs), dtype=np.uint8).copy(), offsets


MASK_MISC_PROBS = (0.5, 0.9, 0.99)


# classify flags of a case (reference spelling -> value); cases without an entry run the defaults
FLAGS = {
    "flags_se": {"--min-score": 0.3, "--min-sp-score": 0.6, "--tie-ratio": 0.9, "--min-cons-cnt": 6, "--min-cons-cnt-euk": 11},
    "acc_lvl1_se": {"--accession-level": 1},
    "lineage_se": {"--lineage": 1},
    "mask_se": {"--mask": 1},
    "mask_pe": {"--mask": 1, "--mask-prob": 0.5},
}


def oracle_flags(name):
    f = FLAGS.get(name, {})
    return dict(min_score=f.get("--min-score", 0.0), min_sp_score=f.get("--min-sp-score", 0.0), tie_ratio=f.get("--tie-ratio", 0.95),
                min_cons=f.get("--min-cons-cnt", 4), min_cons_euk=f.get("--min-cons-cnt-euk", 9),
                accession_level=f.get("--accession-level", 0), lineage=f.get("--lineage", 0))


def mask_flags(name):
    """(--mask, --mask-prob) of a case."""
    f = FLAGS.get(name, {})
    return int(f.get("--mask", 0)), float(f.get("--mask-prob", 0.9))


def build(name):
    from metabuli_b200 import synth
    if name in EDGE_CASES:
        dbkw, rkw, seq_mode = EDGE_CASES[name]
        sdb = _edge_db(**dbkw)
        reads = synth.make_reads(sdb, **rkw)
        if name.startswith("mask_"):
            reads = _inject_read_repeats(reads, rkw["seed"] + 500)
        return sdb, reads, seq_mode
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
