"""The C++ host's parallel FASTA/FASTQ reader and TSV formatter (metabuli_b200/csrc/host/fastx_tsv.hpp) against the Python
mirror of the reference semantics (kseq: KmerExtractor.cpp:429-481; Reporter.cpp:35-80).  CPU only."""
import ctypes as C
import gzip
import os
import subprocess
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "host_io_test.cpp")
HDR = os.path.join(ROOT, "metabuli_b200", "csrc", "host", "fastx_tsv.hpp")
OUT = os.path.join(ROOT, "tests", "host", "_build", "libhost_io.so")


@pytest.fixture(scope="module")
def hio():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in (SRC, HDR)):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", OUT, SRC, "-lz", "-lpthread"])
    L = C.CDLL(OUT)
    L.hio_load.argtypes = [C.c_char_p, C.c_uint]
    L.hio_load.restype = C.c_longlong
    L.hio_load_stream.argtypes = [C.c_char_p, C.c_uint, C.c_ulonglong, C.c_ulonglong]
    L.hio_load_stream.restype = C.c_longlong
    L.hio_total_bases.restype = C.c_ulonglong
    L.hio_copy.argtypes = [C.c_void_p, C.c_void_p]
    L.hio_name.argtypes = [C.c_ulonglong]
    L.hio_name.restype = C.c_char_p
    L.hio_format.argtypes = [C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
    L.hio_format.restype = C.c_ulonglong
    L.hio_text.restype = C.c_char_p
    L.hio_report.argtypes = [C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.c_int32] + [C.c_void_p] * 6
    L.hio_report.restype = C.c_ulonglong
    return L


def _load(hio, path, threads):
    n = hio.hio_load(path.encode(), threads)
    assert n >= 0, hio.hio_text()
    bases = np.zeros(hio.hio_total_bases(), dtype=np.uint8)
    offs = np.zeros(n + 1, dtype=np.uint64)
    hio.hio_copy(bases.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p))
    return [hio.hio_name(i).decode() for i in range(n)], bases, offs


def _write_cases(tmp_path):
    rng = np.random.default_rng(9)
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    n = 60000
    lens = rng.integers(30, 260, n)
    seqs = [bytes(letters[rng.integers(0, 5, int(l))]) for l in lens]
    quals = [bytes(rng.integers(33, 74, int(l)).astype(np.uint8)) for l in lens]      # '!'..'I': lines may start with '@' or '+'
    fq = tmp_path / "reads.fastq"
    with open(fq, "wb") as f:
        for i, (s, q) in enumerate(zip(seqs, quals)):
            f.write(b"@r%d some comment\n%s\n+\n%s\n" % (i, s, q))
    fa = tmp_path / "reads.fna"
    with open(fa, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">s%d\tdesc\r\n" % i)
            for k in range(0, len(s), 70):                                            # multi-line FASTA with CRLF
                f.write(s[k:k + 70] + b"\r\n")
    gz = tmp_path / "reads.fq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(open(fq, "rb").read())
    return [str(fq), str(fa), str(gz)]


@pytest.mark.parametrize("threads", [1, 7])
def test_parallel_reader_matches_python_mirror(hio, tmp_path, fixtures_dir, threads):
    from metabuli_b200.fastx import read_fastx
    paths = _write_cases(tmp_path) + [os.path.join(fixtures_dir, "reads", "ERR9594652_5000_1.fna.gz"),
                                      os.path.join(fixtures_dir, "reads", "ERR9594652_5000_2.fq.gz")]
    for p in paths:
        names, bases, offs = _load(hio, p, threads)
        wn, wb, wo = read_fastx(p)
        assert names == wn, p
        assert np.array_equal(offs, wo) and np.array_equal(bases, wb), p


@pytest.mark.parametrize("batch_reads,chunk_bytes", [(1000, 1 << 16), (7, 4096), (100000, 1 << 20), (5000, 300)])
def test_streaming_reader_equals_whole_file_load(hio, tmp_path, fixtures_dir, batch_reads, chunk_bytes):
    """FastxStream (bounded raw chunks cut at record starts, batches of at most N records, carry-over between chunks) must
    deliver exactly the records of the whole-file reader, whatever the chunk and batch sizes — including chunks smaller than
    one record and FASTQ quality lines that start with '@' or '+'."""
    paths = _write_cases(tmp_path) + [os.path.join(fixtures_dir, "reads", "ERR9594652_5000_2.fq.gz")]
    if chunk_bytes <= 4096:
        paths = paths[:2] if batch_reads > 100 else paths[2:3] + paths[3:]
    for p in paths:
        want = _load(hio, p, 3)
        n = hio.hio_load_stream(p.encode(), 3, batch_reads, chunk_bytes)
        assert n == len(want[0]), hio.hio_text()
        bases = np.zeros(hio.hio_total_bases(), dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        hio.hio_copy(bases.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p))
        assert [hio.hio_name(i).decode() for i in range(0, n, 997)] == want[0][::997]
        assert np.array_equal(offs, want[2]) and np.array_equal(bases, want[1]), p


def test_header_with_leading_blank_has_an_empty_name(hio, tmp_path):
    """kseq takes the name up to the first whitespace after the tag, so "> x" has an empty name and the reference stops with
    "1th entry has no sequence or name." (checked against the reference binary, round 2); all three readers agree."""
    from metabuli_b200.fastx import read_fastx
    p = tmp_path / "blank.fna"
    p.write_bytes(b"> x\nACGTACGT\n")
    assert hio.hio_load(str(p).encode(), 1) == -1 and b"1th entry has no sequence or name." in hio.hio_text()
    with pytest.raises(ValueError, match="1th entry has no sequence or name"):
        read_fastx(str(p))
    # qualities after a '+' line are skipped whatever the record's tag was (kseq grammar)
    q = tmp_path / "plus.fna"
    q.write_bytes(b">a d\nACGT\n+\n>III\n>b\nGGCC\n")
    names, bases, offs = read_fastx(str(q))
    assert names == ["a", "b"] and bytes(bases) == b"ACGTGGCC"
    assert hio.hio_load(str(q).encode(), 1) == 2 and hio.hio_total_bases() == 8


def test_reader_rejects_empty_entries(hio, tmp_path):
    p = tmp_path / "bad.fna"
    p.write_bytes(b">a\nACGT\n>b\n>c\nAC\n")
    assert hio.hio_load(str(p).encode(), 1) == -1
    assert b"2th entry has no sequence or name." in hio.hio_text()
    assert hio.hio_load_stream(str(p).encode(), 1, 1, 8) == -1
    assert b"2th entry has no sequence or name." in hio.hio_text()


@pytest.mark.parametrize("lineage", [False, True])
@pytest.mark.parametrize("threads", [1, 5])
def test_rows_match_python_formatter(hio, tmp_path, threads, lineage):
    from metabuli_b200 import _ffi
    from metabuli_b200.classifier import Classifier
    n = 20000
    fa = tmp_path / "names.fna"
    with open(fa, "wb") as f:
        for i in range(n):
            f.write(b">read_%d/1 x\nACGT\n" % i)
    names, _, _ = _load(hio, str(fa), threads)
    rng = np.random.default_rng(3)
    res = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    res["is_classified"] = rng.integers(0, 2, n)
    res["classification"] = np.where(res["is_classified"] == 1, rng.integers(1, 50, n), 0)
    res["query_length"] = rng.integers(24, 300, n)
    scores = np.concatenate([rng.random(n - 6).astype(np.float32), np.float32([0, 1, 0.5, 1e-5, 0.123456789, 0.0000123])])
    res["score"] = np.where(res["is_classified"] == 1, scores, np.float32(0))
    res["taxcnt_len"] = np.where(res["is_classified"] == 1, rng.integers(1, 5, n), 0)
    res["taxcnt_begin"] = np.concatenate([[0], np.cumsum(res["taxcnt_len"])[:-1]])
    pairs = np.stack([rng.integers(1, 50, int(res["taxcnt_len"].sum())), rng.integers(1, 40, int(res["taxcnt_len"].sum()))], 1).astype(np.int32)
    orig = (np.arange(50, dtype=np.int32) * 7 + 1000)
    rank_names = [b"no rank", b"species", b"genus", b"subspecies", b""]
    ranks = (C.c_char_p * 50)(*[rank_names[i % 5] for i in range(50)])
    lin = (C.c_char_p * 50)(*[b"d_Bacteria;g_G%d;s_G%d sp" % (i, i) for i in range(50)])
    length = hio.hio_format(n, res.ctypes.data_as(C.c_void_p), pairs.ctypes.data_as(C.c_void_p), orig.ctypes.data_as(C.c_void_p), ranks, threads,
                            lin if lineage else None)
    got = hio.hio_text()[:length].decode()
    tax = types.SimpleNamespace(original=lambda t: int(orig[t]), rank_of=lambda t: rank_names[t % 5].decode(),
                                lineage=lambda t: "d_Bacteria;g_G%d;s_G%d sp" % (t, t))
    fake = types.SimpleNamespace(db=types.SimpleNamespace(tax=tax))
    want = Classifier.format_tsv(fake, names, res, pairs, header=False, lineage=lineage)
    assert got == want


def test_python_formatter_with_lineage_matches_reference(golden_dir):
    """Classifier.format_tsv(lineage=True) over the oracle's results == the reference binary's TSV with --lineage 1."""
    import types as _t
    import oracle
    import synth_cases
    from metabuli_b200.classifier import Classifier
    sdb, reads, seq_mode = synth_cases.build("lineage_se")
    odb = oracle.OracleDb.from_synth(sdb)
    v, q, cov1, cov2 = oracle.extract(*reads)
    sv, sq = oracle.sort_kmers(v, q)
    res, pairs = odb.score(oracle.sort_matches(odb.match(sv, sq)), cov1, None, seq_mode=seq_mode)
    fake = _t.SimpleNamespace(db=sdb.database)
    tsv = Classifier.format_tsv(fake, synth_cases.names(reads[1].size - 1), res, pairs, lineage=True).encode()
    assert tsv == gzip.open(os.path.join(golden_dir, "synth", "lineage_se.tsv.gz"), "rb").read()
    odb.close()


@pytest.mark.parametrize("name", ["multi_se", "ties_se", "multi_pe", "acc_prune_se", "flags_se"])
def test_cpp_report_matches_reference(hio, name, golden_dir):
    """mblhost::write_report (the C++ host's <jobid>_report.tsv) over the oracle's per-read classifications == the reference
    binary's report, ties between equally large clades included (children keep node order, std::sort as in the reference)."""
    import oracle
    import synth_cases
    sdb, reads, seq_mode = synth_cases.build(name)
    odb = oracle.OracleDb.from_synth(sdb)
    fl = synth_cases.oracle_flags(name)
    fl.pop("lineage", None)
    if sdb.database.params.accession_level_db == 1 and fl["accession_level"] == 0:
        fl["accession_level"] = 2
    v, q, cov1, cov2 = oracle.extract(*reads, kmer_format=sdb.database.params.kmer_format)
    sv, sq = oracle.sort_kmers(v, q)
    res, _ = odb.score(oracle.sort_matches(odb.match(sv, sq)), cov1, cov2 if seq_mode == 2 else None, seq_mode=seq_mode, **fl)
    t = sdb.database.tax
    cls = np.ascontiguousarray(res["classification"], dtype=np.int32)
    n = t.max_nodes
    ranks = (C.c_char_p * n)(*[t.node_rank_name[i].encode() for i in range(n)])
    names = (C.c_char_p * n)(*[t.string(int(t.node_name_idx[i])).encode() for i in range(n)])
    orig = np.ascontiguousarray(t.internal2org, dtype=np.int32) if t.internal_ids else None
    length = hio.hio_report(cls.size, cls.ctypes.data_as(C.c_void_p), n, t.max_taxid, t.node_taxid.ctypes.data_as(C.c_void_p),
                            t.node_parent.ctypes.data_as(C.c_void_p), t.D.ctypes.data_as(C.c_void_p),
                            orig.ctypes.data_as(C.c_void_p) if orig is not None else None, ranks, names)
    got = hio.hio_text()[:length]
    assert got == gzip.open(os.path.join(golden_dir, "synth", name + ".report.gz"), "rb").read()
    odb.close()


@pytest.mark.parametrize("db", ["in", "ex"])
def test_cpp_report_on_the_reference_fixture(hio, db, fixtures_dir, golden_dir):
    """Same on the regression fixture database (original taxids differ from internal ones, the report walk starts at internal
    taxid 1 = Viruses rather than at the root, Q11): per-read classifications parsed back from the reference's own TSV."""
    from metabuli_b200 import load_database
    d = load_database(os.path.join(fixtures_dir, f"db_{db}"))
    t = d.tax
    org2int = {int(o): i for i, o in enumerate(t.internal2org) if t.D[i] != -1} if t.internal_ids else None
    rows = gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_se_classifications.tsv.gz"), "rt").read().split("\n")[1:]
    cls = np.array([(org2int[int(r.split("\t")[2])] if org2int and int(r.split("\t")[2]) else int(r.split("\t")[2])) for r in rows if r], dtype=np.int32)
    n = t.max_nodes
    ranks = (C.c_char_p * n)(*[t.node_rank_name[i].encode() for i in range(n)])
    names = (C.c_char_p * n)(*[t.string(int(t.node_name_idx[i])).encode() for i in range(n)])
    orig = np.ascontiguousarray(t.internal2org, dtype=np.int32) if t.internal_ids else None
    length = hio.hio_report(cls.size, cls.ctypes.data_as(C.c_void_p), n, t.max_taxid, t.node_taxid.ctypes.data_as(C.c_void_p),
                            t.node_parent.ctypes.data_as(C.c_void_p), t.D.ctypes.data_as(C.c_void_p),
                            orig.ctypes.data_as(C.c_void_p) if orig is not None else None, ranks, names)
    assert hio.hio_text()[:length] == gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_se_report.tsv.gz"), "rb").read()


def test_fast_float_formatting_equals_printf_g(hio):
    """The TSV's score column is `ostream << float` = printf("%g") (Reporter.cpp:62).  The formatter's integer fast path covers
    1e-4 <= v < 10: compared with snprintf on EVERY float of that range (139 M values) and on values outside it."""
    hio.hio_check_float_g.restype = C.c_ulonglong
    hio.hio_check_float_g.argtypes = [C.c_uint, C.c_uint]
    assert hio.hio_check_float_g(1, os.cpu_count() or 4) == 0


def _gzip_streams():
    """(label, compressed, plain) over block types, levels, strategies, window sizes and contents."""
    import zlib
    rng = np.random.default_rng(11)
    dna = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 300000)])
    fastq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, dna[i * 100:i * 100 + 100], bytes(rng.integers(33, 74, 100).astype(np.uint8))) for i in range(1500))
    texts = {
        "empty": b"", "one": b"A", "dna": dna, "fastq": fastq, "random": bytes(rng.integers(0, 256, 200000).astype(np.uint8)),
        "zeros": bytes(400000), "run": b"AC" * 150000, "period7": b"ACGTTGA" * 40000,
        "far": dna[:40000] + dna[:40000] + dna[5000:38000],                         # distances close to 32 KiB
        "skew": bytes(np.minimum(rng.geometric(0.02, 300000), 255).astype(np.uint8)),   # long Huffman codes (sub-tables)
    }
    out = []
    for name, t in texts.items():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                for wbits in (31, 25):
                    if wbits == 25 and (level != 6 or strategy != zlib.Z_DEFAULT_STRATEGY):
                        continue
                    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
                    out.append(("%s/l%d/s%d/w%d" % (name, level, strategy, wbits), c.compress(t) + c.flush(), t))
    # several members, an FEXTRA / FNAME / FCOMMENT / FHCRC header, sync-flush blocks (empty stored blocks), trailing zero padding
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = c.compress(dna[:1000]) + c.flush(zlib.Z_SYNC_FLUSH) + c.compress(dna[1000:70000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(dna[70000:]) + c.flush()
    out.append(("flushes", parts, dna))
    m1 = zlib.compressobj(9, zlib.DEFLATED, 31); a = m1.compress(fastq[:5000]) + m1.flush()
    m2 = zlib.compressobj(1, zlib.DEFLATED, 31); b = m2.compress(fastq[5000:]) + m2.flush()
    out.append(("members", a + b + bytes(37), fastq))
    raw = zlib.compressobj(6, zlib.DEFLATED, -15); body = raw.compress(dna) + raw.flush()
    import struct
    hdr = b"\x1f\x8b\x08" + bytes([4 | 8 | 16 | 2]) + b"\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC\x02\x00\x10\x00" + b"name.fq\0" + b"a comment\0" + b"\x12\x34"
    out.append(("header-fields", hdr + body + struct.pack("<II", zlib.crc32(dna), len(dna) & 0xFFFFFFFF), dna))
    return out


def test_own_inflate_equals_zlib(hio):
    """The reader's gzip decoder (csrc/host/fast_inflate.hpp) against zlib-written streams: stored, fixed and dynamic blocks, every
    strategy, 8 KiB..32 KiB windows, long codes, far distances, members / header fields / flushes / padding; the output is taken
    in one call, in odd pieces and byte by byte (matches resumed across calls, history in front of the buffer)."""
    hio.hio_inflate_check.restype = C.c_int
    hio.hio_inflate_check.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_ulonglong]
    for label, comp, plain in _gzip_streams():
        pieces = (0, 1, 263, 40000) if len(plain) <= 70000 or label.startswith(("far", "members", "flushes")) else (0, 263, 40000)
        for piece in pieces:
            rc = hio.hio_inflate_check(comp, len(comp), plain, len(plain), piece)
            assert rc == 0, (label, piece, rc, hio.hio_text())


def test_own_inflate_rejects_damaged_streams(hio):
    """Truncations and bit flips: an error or (for flips the format cannot see) a CRC / length mismatch — never a crash, never
    silently different output."""
    import zlib
    hio.hio_inflate_check.restype = C.c_int
    hio.hio_inflate_check.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_ulonglong]
    rng = np.random.default_rng(12)
    plain = b"".join(b">r%d\n%s\n" % (i, bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 150)])) for i in range(400))
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    comp = c.compress(plain) + c.flush()
    assert hio.hio_inflate_check(comp, len(comp), plain, len(plain), 0) == 0
    for cut in (1, 5, 10, 11, 50, len(comp) // 2, len(comp) - 9, len(comp) - 8, len(comp) - 1):
        assert hio.hio_inflate_check(comp[:cut], cut, plain, len(plain), 0) == 1, cut
    for k in range(600):
        pos = int(rng.integers(0, len(comp)))
        bad = bytearray(comp)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        rc = hio.hio_inflate_check(bytes(bad), len(bad), plain, len(plain), int(rng.choice([0, 97])))
        assert rc == 1 or (rc == 0 and 4 <= pos < 10), (pos, rc)      # MTIME / XFL / OS bytes of the header carry no checked information


def test_parallel_inflate_equals_zlib(hio):
    """GzParallel (one gzip stream decoded by several threads: block-start search, 16-bit symbols with window markers, in-order
    resolution) must give zlib's bytes for every thread count — on sequence text at several levels (many dynamic blocks), on a
    file of several members, on data whose blocks cannot be found (binary: falls back to the serial path) and on small inputs."""
    import zlib
    hio.hio_parallel_inflate_check.restype = C.c_int
    hio.hio_parallel_inflate_check.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_uint, C.c_ulonglong]
    rng = np.random.default_rng(21)
    L = np.frombuffer(b"ACGT", dtype=np.uint8)
    n = 150000
    seqs = L[rng.integers(0, 4, (n, 150))]
    qual = rng.integers(35, 74, (n, 150)).astype(np.uint8)
    fastq = b"".join(b"@read%d/1\n%s\n+\n%s\n" % (i, seqs[i].tobytes(), qual[i].tobytes()) for i in range(n))       # ~47 MB
    fasta = b"".join(b">r%d\n%s\n" % (i, seqs[i].tobytes()) for i in range(n))
    cases = []
    for level in (1, 6):
        c = zlib.compressobj(level, zlib.DEFLATED, 31)
        cases.append(("fastq-l%d" % level, c.compress(fastq) + c.flush(), fastq))
    c = zlib.compressobj(1, zlib.DEFLATED, 31)
    cases.append(("fasta-l1", c.compress(fasta) + c.flush(), fasta))
    parts = []
    for k in range(3):
        c = zlib.compressobj(4, zlib.DEFLATED, 31)
        parts.append(c.compress(fastq[k * len(fastq) // 3:(k + 1) * len(fastq) // 3]) + c.flush())
    cases.append(("members", b"".join(parts), fastq))
    bgzf = []
    for k in range(0, 12_000_000, 65000):                                          # bgzip-style: hundreds of small members
        c = zlib.compressobj(6, zlib.DEFLATED, 31)
        bgzf.append(c.compress(fastq[k:k + 65000]) + c.flush())
    cases.append(("small-members", b"".join(bgzf), fastq[:12_025_000]))
    import struct
    blocks = []
    for k in range(0, 9_000_000, 65280):                                           # BGZF as bgzip writes it: BC subfield, 64 KiB blocks, end marker
        chunk = fastq[k:k + 65280]
        raw = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = raw.compress(chunk) + raw.flush()
        bsize = 12 + 6 + len(body) + 8
        blocks.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + body +
                      struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    blocks.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    cases.append(("bgzf", b"".join(blocks), fastq[:len(range(0, 9_000_000, 65280)) * 65280][:9_000_000 + 65280]))
    binary = bytes(rng.integers(0, 256, 9_000_000).astype(np.uint8))
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    cases.append(("binary", c.compress(binary) + c.flush(), binary))
    mixed = fastq[:6_000_000] + binary[:5_000_000] + fastq[6_000_000:14_000_000]
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    cases.append(("mixed", c.compress(mixed) + c.flush(), mixed))
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    cases.append(("small", c.compress(fastq[:100000]) + c.flush(), fastq[:100000]))
    cases.append(("empty", zlib.compressobj(6, zlib.DEFLATED, 31).flush(), b""))
    for label, comp, plain in cases:
        for threads, piece in ((2, 64 << 20), (3, 1 << 20), (8, 64 << 20), (1, 64 << 20)):
            rc = hio.hio_parallel_inflate_check(comp, len(comp), plain, len(plain), threads, piece)
            assert rc == 0, (label, threads, piece, rc, hio.hio_text())
            words = hio.hio_text().decode().split()
            groups, segments = int(words[1]), int(words[3])
            if threads >= 2 and label.startswith(("fastq", "fasta", "members")):
                assert groups >= 1 and segments >= 2 * groups, (label, threads, hio.hio_text())     # the parallel path really ran
            if threads == 1 or label in ("small", "empty"):
                assert groups == 0
            if label == "bgzf":
                assert (int(words[7]) >= 1) == (threads >= 2), hio.hio_text()
    # damage inside the parallel region is an error, not different output
    label, comp, plain = cases[0]
    for pos in (len(comp) // 3, len(comp) // 2, len(comp) - 100):
        bad = bytearray(comp)
        bad[pos] ^= 0x10
        assert hio.hio_parallel_inflate_check(bytes(bad), len(bad), plain, len(plain), 4, 64 << 20) == 1, pos
