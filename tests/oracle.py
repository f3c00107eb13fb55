"""ctypes wrapper of the CPU oracle (oracle/_build/libmbl_oracle.so).  TEST INFRASTRUCTURE ONLY — the
checker the CUDA path is compared against; never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libmbl_oracle.so")

RESULT_DTYPE = np.dtype([("classification", "<i4"), ("score", "<f4"), ("hamming", "<i4"), ("query_length", "<i4"),
                         ("taxcnt_begin", "<u4"), ("taxcnt_len", "<u4"), ("is_classified", "u1"), ("pad", "u1", (3,))])
MATCH_DTYPE = np.dtype([("qinfo", "<u8"), ("target_id", "<i4"), ("species_id", "<i4"), ("dna_encoding", "<u4"),
                        ("right_end_hamming", "<u2"), ("hamming", "u1"), ("pad", "u1")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB)
        vp, sz = C.c_void_p, C.c_size_t
        L.orc_db_open.argtypes = [C.c_char_p, C.c_char_p, sz]
        L.orc_db_open.restype = vp
        L.orc_db_close.argtypes = [vp]
        L.orc_db_from_arrays.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, C.c_int, C.c_int, C.c_char_p, sz]
        L.orc_db_from_arrays.restype = vp
        L.orc_db_kmer_format.argtypes = [vp]
        L.orc_extract.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_int, vp, vp, sz, C.POINTER(sz), vp, vp]
        L.orc_extract2.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_int, C.c_int, C.c_int, vp, vp, sz, C.POINTER(sz), vp, vp]
        L.orc_db_set_syncmer.argtypes = [vp, C.c_int, C.c_int]
        L.orc_db_syncmer.argtypes = [vp]
        L.orc_sort_kmers.argtypes = [vp, vp, sz, C.c_int]
        L.orc_match.argtypes = [vp, vp, vp, sz, vp, sz, C.POINTER(sz), C.c_int]
        L.orc_sort_matches.argtypes = [vp, sz, C.c_int]
        L.orc_score.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, vp, sz, C.c_uint32,
                                vp, vp, vp, vp, sz, C.POINTER(sz), C.c_int]
        L.orc_classify_arrays.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_uint32, vp, C.POINTER(sz), C.POINTER(sz)]
        L.orc_classify_arrays.restype = C.c_double
        L.orc_classify_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(sz),
                                         C.POINTER(sz), C.c_char_p, sz]
        L.orc_set_flags.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_set_lineage.argtypes = [C.c_int]
        L.orc_next_target_kmer.argtypes = [C.c_uint64, vp, C.POINTER(sz)]
        L.orc_next_target_kmer.restype = C.c_uint64
        L.orc_hamming_sum.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_hamming_sum.restype = C.c_uint8
        L.orc_hammings_fwd.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_hammings_fwd.restype = C.c_uint16
        L.orc_hammings_rev.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_hammings_rev.restype = C.c_uint16
        L.orc_max_covered_length.argtypes = [C.c_int]
        L.orc_query_kmer_number.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleDb:
    def __init__(self, db_dir: str | None, handle=None):
        if handle is not None:
            self.h = handle
        else:
            err = C.create_string_buffer(512)
            self.h = lib().orc_db_open(db_dir.encode(), err, 512)
            if not self.h:
                raise RuntimeError(err.value.decode())
        self.kmer_format = lib().orc_db_kmer_format(self.h)
        self.smer_len = lib().orc_db_syncmer(self.h)            # 0 = not a syncmer database

    @classmethod
    def from_synth(cls, sdb):
        """Build from a metabuli_b200.synth.SynthDb without going through files."""
        d = sdb.database
        diff = np.ascontiguousarray(d.diff_idx); info = np.ascontiguousarray(d.info); split = np.ascontiguousarray(d.split)
        blob = np.frombuffer(sdb.taxonomy_blob, dtype=np.uint8)
        tl = np.ascontiguousarray(sdb.taxid_list, dtype=np.int32)
        err = C.create_string_buffer(512)
        h = lib().orc_db_from_arrays(_p(diff), diff.size, _p(info), info.size, _p(split), split.size // 3, _p(blob), blob.size,
                                     _p(tl), tl.size, d.params.kmer_format, d.params.skip_redundancy, err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        lib().orc_db_set_syncmer(h, int(d.params.syncmer), int(d.params.smer_len))
        return cls(None, handle=h)

    def close(self):
        if self.h:
            lib().orc_db_close(self.h)
            self.h = None

    def match(self, value, qinfo, threads=4):
        v = np.ascontiguousarray(value, dtype=np.uint64)
        q = np.ascontiguousarray(qinfo, dtype=np.uint64)
        cap = max(1024, v.size)
        while True:
            out = np.zeros(cap, dtype=MATCH_DTYPE)
            n = C.c_size_t(0)
            rc = lib().orc_match(self.h, _p(v), _p(q), v.size, _p(out), cap, C.byref(n), threads)
            if rc == 1:
                cap = n.value + 16
                continue
            if rc != 0:
                raise RuntimeError(f"orc_match failed {rc}")
            return out[: n.value]

    def score(self, sorted_matches, cov1, cov2=None, seq_mode=1, min_score=0.0, min_sp_score=0.0, tie_ratio=0.95,
              min_cons=4, min_cons_euk=9, accession_level=0, threads=4):
        m = np.ascontiguousarray(sorted_matches, dtype=MATCH_DTYPE)
        c1 = np.ascontiguousarray(cov1, dtype=np.int32)
        c2 = np.ascontiguousarray(cov2, dtype=np.int32) if cov2 is not None else None
        n = c1.size
        out = np.zeros(n, dtype=RESULT_DTYPE)
        cap = 1 << 16
        while True:
            pairs = np.zeros((cap, 2), dtype=np.int32)
            used = C.c_size_t(0)
            rc = lib().orc_score(self.h, seq_mode, min_score, min_sp_score, tie_ratio, min_cons, min_cons_euk, accession_level,
                                 _p(m), m.size, n, _p(c1), _p(c2), _p(out), _p(pairs), cap, C.byref(used), threads)
            if rc == 2:
                cap = used.value + 16
                continue
            if rc != 0:
                raise RuntimeError(f"orc_score failed {rc}")
            return out, pairs[: used.value]

    def classify_arrays(self, bases1, off1, bases2=None, off2=None, seq_mode=1, threads=1, want_results=True):
        n = off1.size - 1
        res = np.zeros(n, dtype=RESULT_DTYPE) if want_results else None
        nk, nm = C.c_size_t(0), C.c_size_t(0)
        sec = lib().orc_classify_arrays(self.h, seq_mode, threads, _p(bases1), _p(off1), _p(bases2), _p(off2), n, _p(res),
                                        C.byref(nk), C.byref(nm))
        if sec < 0:
            raise RuntimeError("orc_classify_arrays failed")
        return sec, res, nk.value, nm.value


def extract(bases1, off1, bases2=None, off2=None, kmer_format=2, syncmer=0, smer_len=5):
    b1 = np.ascontiguousarray(bases1, dtype=np.uint8)
    o1 = np.ascontiguousarray(off1, dtype=np.uint64)
    b2 = np.ascontiguousarray(bases2, dtype=np.uint8) if bases2 is not None else None
    o2 = np.ascontiguousarray(off2, dtype=np.uint64) if off2 is not None else None
    n = o1.size - 1
    cov1 = np.zeros(n, dtype=np.int32)
    cov2 = np.zeros(n, dtype=np.int32)
    cnt = C.c_size_t(0)
    lib().orc_extract2(_p(b1), _p(o1), _p(b2), _p(o2), n, kmer_format, syncmer, smer_len, None, None, 0, C.byref(cnt), _p(cov1), _p(cov2))
    value = np.zeros(cnt.value, dtype=np.uint64)
    qinfo = np.zeros(cnt.value, dtype=np.uint64)
    rc = lib().orc_extract2(_p(b1), _p(o1), _p(b2), _p(o2), n, kmer_format, syncmer, smer_len, _p(value), _p(qinfo), cnt.value,
                            C.byref(cnt), _p(cov1), _p(cov2))
    assert rc == 0
    return value, qinfo, cov1, cov2


def sort_kmers(value, qinfo, threads=4):
    v = np.ascontiguousarray(value, dtype=np.uint64).copy()
    q = np.ascontiguousarray(qinfo, dtype=np.uint64).copy()
    lib().orc_sort_kmers(_p(v), _p(q), v.size, threads)
    return v, q


def sort_matches(m, threads=4):
    m = np.ascontiguousarray(m, dtype=MATCH_DTYPE).copy()
    lib().orc_sort_matches(_p(m), m.size, threads)
    return m


def classify_files(q1, q2, db_dir, seq_mode, out_path, threads=1, min_score=0.0, min_sp_score=0.0, tie_ratio=0.95, min_cons=4,
                   min_cons_euk=9, accession_level=0, lineage=0):
    lib().orc_set_flags(min_score, min_sp_score, tie_ratio, min_cons, min_cons_euk, accession_level)
    lib().orc_set_lineage(lineage)
    err = C.create_string_buffer(512)
    nk, nm = C.c_size_t(0), C.c_size_t(0)
    rc = lib().orc_classify_files(q1.encode(), q2.encode() if q2 else None, db_dir.encode(), seq_mode, threads, out_path.encode(),
                                  C.byref(nk), C.byref(nm), err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return nk.value, nm.value
