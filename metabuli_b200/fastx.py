"""FASTA/FASTQ reader with kseq semantics (reference: lib/mmseqs/src/commons/KSeqWrapper.h:9-33 as used in
KmerExtractor.cpp:429-481): name = first whitespace-delimited token of the header, sequence = all graphic
characters of the record's sequence lines.  Produces the SoA layout of mbl_batch (bases + offsets)."""
from __future__ import annotations

import gzip

import numpy as np


def read_fastx(path: str):
    """-> (names: list[str], bases: np.uint8[total], offsets: np.uint64[n+1])"""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    names, seqs = [], []
    lines = data.split(b"\n")
    i, n = 0, len(lines)
    while i < n:
        ln = lines[i]
        if not ln or ln[:1] not in (b">", b"@"):
            i += 1
            continue
        hdr = ln[1:].rstrip(b"\r")
        # kseq: the name is what stands between the tag and the first whitespace — "> x" has an EMPTY name
        k = 0
        while k < len(hdr) and hdr[k] not in b" \t\v\f\r\n":
            k += 1
        names.append(hdr[:k].decode())
        i += 1
        parts = []
        while i < n and lines[i][:1] not in (b">", b"@", b"+"):
            parts.append(lines[i].rstrip(b"\r"))
            i += 1
        seq = b"".join(parts)
        seq = bytes(c for c in seq if 33 <= c <= 126) if any(c < 33 or c > 126 for c in seq) else seq
        if i < n and lines[i][:1] == b"+":                  # kseq reads qualities after a '+' line whatever the record's tag was
            i += 1
            ql = 0
            while i < n and ql < len(seq):
                ql += len(lines[i].rstrip(b"\r"))
                i += 1
        seqs.append(seq)
        if not seq or not names[-1]:                        # QueryIndexer.cpp:50-53 / KmerExtractor.cpp:447-451: the reference exits
            raise ValueError(f"{len(seqs)}th entry has no sequence or name.")
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offsets[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, dtype=np.uint8)
    return names, bases, offsets
