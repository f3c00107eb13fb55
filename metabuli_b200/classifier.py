"""Host-side mirror of the reference's Classifier for the classify hot path
(src/commons/Classifier.{h,cpp}: ctor :6-32, startClassify :44-164; Reporter.cpp:35-80 for the TSV).

Same names and argument meaning as the reference's operator surface; every stage runs in
libmetabuli_b200.so through the C-ABI (include/metabuli_b200.h).  Python here is plumbing: it loads the
on-disk DB into host arrays, hands pointers across the boundary and formats the per-read TSV.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _ffi
from .dbio import Database, load_database
from .fastx import read_fastx


@dataclass
class ClassifyOptions:
    """classify flags the path reads (defaults: src/workflow/classify.cpp:10-37)."""
    seq_mode: int = 2
    min_score: float = 0.0
    min_sp_score: float = 0.0
    tie_ratio: float = 0.95
    min_cons_cnt: int = 4
    min_cons_cnt_euk: int = 9
    accession_level: int = 0
    match_per_kmer: int = 4
    device: int = 0
    mask: int = 0                 # --mask 1: tantan masking of the queries before extraction (KmerExtractor.cpp:308-314)
    mask_prob: float = 0.9
    mask_on_host: bool = False    # False: the library masks every uploaded batch on the device (mbl_config.mask_mode, k0_mask.cu);
                                  # True: mbl_mask_reads on the host threads before the upload — same letters either way
    threads: int = 0              # host threads of the host-side masking (0 = all)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Classifier:
    """Classifier(par): loads the DB (loadDbParameters, loadTaxonomy, KmerMatcher::loadTaxIdList) onto the GPU."""

    def __init__(self, db_dir: str | None, opt: ClassifyOptions | None = None, database: Database | None = None, shard=None,
                 total_kmers: int | None = None):
        """shard: an _ffi.Shard from sharded.plan_shards -> only that value range of the index is uploaded (mbl_load_db_shard).
        total_kmers: k-mers of the WHOLE index when `database` holds one rank's part only (sizes the presence filter, which
        every rank must lay out identically)."""
        self.opt = opt or ClassifyOptions()
        self.lib = _ffi.load_library()
        self.db = database if database is not None else load_database(db_dir)
        p = self.db.params
        # loadDbParameters (common.cpp:88-133): "Accession_level 1" in the DB turns a default 0 into 2
        acc = self.opt.accession_level
        if p.accession_level_db == 1 and acc == 0:
            acc = 2
        if p.accession_level_db == 0 and acc == 1:
            acc = 0
        self.cfg = _ffi.Config(kmer_format=p.kmer_format, reduced_aa=p.reduced_aa, skip_redundancy=p.skip_redundancy,
                               syncmer=p.syncmer, smer_len=p.smer_len, seq_mode=self.opt.seq_mode,
                               min_score=self.opt.min_score, min_sp_score=self.opt.min_sp_score, tie_ratio=self.opt.tie_ratio,
                               min_cons_cnt=self.opt.min_cons_cnt, min_cons_cnt_euk=self.opt.min_cons_cnt_euk,
                               accession_level=acc, device=self.opt.device, match_per_kmer=self.opt.match_per_kmer,
                               mask_mode=1 if (self.opt.mask and not self.opt.mask_on_host) else 0, mask_prob=self.opt.mask_prob)
        self.ctx = C.c_void_p()
        rc = self.lib.mbl_create(C.byref(self.cfg), C.byref(self.ctx))
        if rc != _ffi.MBL_OK:
            raise _ffi.MblError(rc, "mbl_create failed (no usable CUDA device?)" if rc == _ffi.MBL_E_NO_DEVICE else "mbl_create failed")
        t = self.db.tax
        self._keep = [np.ascontiguousarray(self.db.diff_idx), np.ascontiguousarray(self.db.info), np.ascontiguousarray(self.db.split)]
        dbs = _ffi.Db(_ptr(self._keep[0]), self._keep[0].size, _ptr(self._keep[1]), max(self._keep[1].size, int(total_kmers or 0)),
                      _ptr(self._keep[2]), self._keep[2].size // 3)
        tx = _ffi.Taxonomy(t.max_nodes, t.max_taxid, t.eukaryota, _ptr(t.D), _ptr(t.E), _ptr(t.L), _ptr(t.H), _ptr(t.M), t.M_k,
                           _ptr(t.node_taxid), _ptr(t.node_parent), _ptr(t.node_prune), _ptr(t.node_rank),
                           _ptr(self.db.taxid2species))
        if shard is None:
            self._check(self.lib.mbl_load_db(self.ctx, C.byref(dbs), C.byref(tx)))
        else:
            self._check(self.lib.mbl_load_db_shard(self.ctx, C.byref(dbs), C.byref(tx), C.byref(shard)))

    # ---------------------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != _ffi.MBL_OK:
            raise _ffi.MblError(rc, self.lib.mbl_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self._dl_free()
            self.lib.mbl_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mask_reads(self, bases, offsets, mask_prob: float | None = None):
        """SeqIterator::maskLowComplexityRegions (SeqIterator.cpp:154-175) over a batch: a masked COPY of `bases` (letters whose
        tantan repeat probability reaches --mask-prob, and letters that are no nucleotides, become 'N').  Host work
        (mbl_mask_reads); the reads handed to the extractor are these, the lengths the TSV prints are unchanged."""
        b = np.array(bases, dtype=np.uint8, copy=True, order="C")
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self.lib.mbl_mask_reads(_ptr(b), _ptr(o), o.size - 1, C.c_float(self.opt.mask_prob if mask_prob is None else mask_prob),
                                            self.opt.threads))
        return b

    def make_batch(self, bases1, off1, bases2=None, off2=None):
        b1 = np.ascontiguousarray(bases1, dtype=np.uint8)
        o1 = np.ascontiguousarray(off1, dtype=np.uint64)
        b2 = np.ascontiguousarray(bases2, dtype=np.uint8) if bases2 is not None else None
        o2 = np.ascontiguousarray(off2, dtype=np.uint64) if off2 is not None else None
        if self.opt.mask and self.opt.mask_on_host:
            b1 = self.mask_reads(b1, o1)
            b2 = self.mask_reads(b2, o2) if b2 is not None else None
        batch = _ffi.Batch(_ptr(b1), _ptr(o1), _ptr(b2), _ptr(o2), o1.size - 1)
        return batch, (b1, o1, b2, o2)

    _n_resident = 0

    # ---- whole path -----------------------------------------------------------------------------
    def classify_batch(self, bases1, off1, bases2=None, off2=None):
        """One QuerySplit through extract/sort/match/sort/score.  -> (results[n], taxcnt_pairs[k,2])"""
        batch, keep = self.make_batch(bases1, off1, bases2, off2)
        n = batch.n_reads
        out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        cap = 1 << 16
        while True:
            pairs = np.zeros((cap, 2), dtype=np.int32)
            used = C.c_size_t(0)
            rc = self.lib.mbl_classify_batch(self.ctx, C.byref(batch), _ptr(out), _ptr(pairs), cap, C.byref(used))
            if rc == _ffi.MBL_E_CAPACITY:
                cap = int(used.value) + 16
                continue
            self._check(rc)
            del keep
            return out, pairs[: used.value]

    def classify_stream(self, batches):
        """The QuerySplit loop over many batches: yields (results, taxcnt_pairs) per batch of `batches` (an iterable of
        (bases1, off1[, bases2, off2]) tuples) while the NEXT batch's reads are already on their way to the device
        (mbl_prefetch_batch / mbl_classify_prefetched)."""
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        cur_b, cur_keep = self.make_batch(*cur)
        self._check(self.lib.mbl_prefetch_batch(self.ctx, C.byref(cur_b)))
        while cur is not None:
            nxt = next(it, None)
            nxt_b, nxt_keep = self.make_batch(*nxt) if nxt is not None else (None, None)
            n = cur_b.n_reads
            out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
            cap = max(1 << 16, 5 * n)
            used = C.c_size_t(0)
            pairs = np.zeros((cap, 2), dtype=np.int32)
            rc = self.lib.mbl_classify_prefetched(self.ctx, C.byref(nxt_b) if nxt_b is not None else None, _ptr(out), _ptr(pairs), cap, C.byref(used))
            if rc == _ffi.MBL_E_CAPACITY:             # the batch is classified and resident: only the download is repeated
                cap = int(used.value) + 16
                pairs = np.zeros((cap, 2), dtype=np.int32)
                rc = self.lib.mbl_download_results(self.ctx, _ptr(out), _ptr(pairs), cap, C.byref(used))
            self._check(rc)
            yield out, pairs[: used.value]
            cur, cur_b, cur_keep = nxt, nxt_b, nxt_keep

    def release_host_index(self):
        """Drop this object's references to the host copies of diffIdx / info (they are only read by mbl_load_db); several ranks
        on one node would otherwise each keep ~the index size in host memory."""
        self._keep = None
        self.db.diff_idx = np.zeros(0, dtype=np.uint16)
        self.db.info = np.zeros(0, dtype=np.int32)

    def download_results(self):
        """mbl_download_results of the resident batch -> (results[n], taxcnt_pairs[k,2]).  The arrays are views of two pinned
        buffers owned by this object and are overwritten by the next call."""
        n = int(self._n_resident)
        if self._dl_out is None or self._dl_out.size < n:
            self._dl_free()
            self._dl_out = np.zeros(max(16, n + n // 8), dtype=_ffi.RESULT_DTYPE)
            self._dl_pairs = np.zeros((max(16, 5 * n), 2), dtype=np.int32)
            for a in (self._dl_out, self._dl_pairs):
                self.lib.mbl_host_register(_ptr(a), a.nbytes)
        while True:
            used = C.c_size_t(0)
            rc = self.lib.mbl_download_results(self.ctx, _ptr(self._dl_out), _ptr(self._dl_pairs), self._dl_pairs.shape[0], C.byref(used))
            if rc == _ffi.MBL_E_CAPACITY:
                self.lib.mbl_host_unregister(_ptr(self._dl_pairs))
                self._dl_pairs = np.zeros((int(used.value) + int(used.value) // 8 + 16, 2), dtype=np.int32)
                self.lib.mbl_host_register(_ptr(self._dl_pairs), self._dl_pairs.nbytes)
                continue
            self._check(rc)
            return self._dl_out[:n], self._dl_pairs[: used.value]

    _dl_out = None
    _dl_pairs = None

    def _dl_free(self):
        for a in (self._dl_out, self._dl_pairs):
            if a is not None:
                self.lib.mbl_host_unregister(_ptr(a))
        self._dl_out = self._dl_pairs = None

    def stats(self) -> dict:
        s = _ffi.Stats()
        self.lib.mbl_get_stats(self.ctx, C.byref(s))
        d = {f"ms_{n}": float(s.ms[i]) for i, n in enumerate(_ffi.STAGE_NAMES)}
        d.update(ms_merge_kernel=float(s.merge_kernel_ms), n_query_kmers=int(s.n_query_kmers), n_matches=int(s.n_matches), merge_bytes=int(s.merge_bytes),
                 merge_launches=int(s.merge_launches), kernel_launches=int(s.kernel_launches),
                 overflow_retries=int(s.overflow_retries), sub_batches=int(s.sub_batches), ms_bucket_kmers=float(s.ms_bucket_kmers),
                 ms_bucket_matches=float(s.ms_bucket_matches), n_merge_queries=int(s.n_merge_queries),
                 ms_push_kmers=float(s.ms_push_kmers), ms_push_matches=float(s.ms_push_matches), ms_mask=float(s.ms_mask))
        return d

    def db_info(self) -> dict:
        s = _ffi.DbInfo()
        self.lib.mbl_get_db_info(self.ctx, C.byref(s))
        return dict(n_tiles=int(s.n_tiles), n_jumbo=int(s.n_jumbo), n_kmers=int(s.n_kmers), n_u16=int(s.n_u16), hbm_bytes=int(s.hbm_bytes))

    # ---- stages (parity tests) ------------------------------------------------------------------
    def extract(self, bases1, off1, bases2=None, off2=None):
        batch, keep = self.make_batch(bases1, off1, bases2, off2)
        n = C.c_size_t(0)
        rc = self.lib.mbl_extract(self.ctx, C.byref(batch), None, None, 0, C.byref(n))
        if rc not in (_ffi.MBL_OK, _ffi.MBL_E_CAPACITY):
            self._check(rc)
        value = np.zeros(n.value, dtype=np.uint64)
        qinfo = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self.lib.mbl_extract(self.ctx, C.byref(batch), _ptr(value), _ptr(qinfo), n.value, C.byref(n)))
        del keep
        return value, qinfo

    def sort_kmers(self, value, qinfo):
        v = np.ascontiguousarray(value, dtype=np.uint64).copy()
        q = np.ascontiguousarray(qinfo, dtype=np.uint64).copy()
        self._check(self.lib.mbl_sort_kmers(self.ctx, _ptr(v), _ptr(q), v.size))
        return v, q

    def match(self, value, qinfo, cap=None):
        v = np.ascontiguousarray(value, dtype=np.uint64)
        q = np.ascontiguousarray(qinfo, dtype=np.uint64)
        cap = cap or max(1024, 4 * v.size)
        while True:
            out = np.zeros(cap, dtype=_ffi.MATCH_DTYPE)
            n = C.c_size_t(0)
            rc = self.lib.mbl_match(self.ctx, _ptr(v), _ptr(q), v.size, _ptr(out), cap, C.byref(n))
            if rc == _ffi.MBL_E_MATCH_OVERFLOW and n.value > cap:
                cap = int(n.value) + 16
                continue
            self._check(rc)
            return out[: n.value]

    def sort_matches(self, matches):
        m = np.ascontiguousarray(matches, dtype=_ffi.MATCH_DTYPE).copy()
        self._check(self.lib.mbl_sort_matches(self.ctx, _ptr(m), m.size))
        return m

    def score(self, sorted_matches, cov1, cov2=None):
        m = np.ascontiguousarray(sorted_matches, dtype=_ffi.MATCH_DTYPE)
        c1 = np.ascontiguousarray(cov1, dtype=np.int32)
        c2 = np.ascontiguousarray(cov2, dtype=np.int32) if cov2 is not None else None
        n = c1.size
        out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        cap = 1 << 16
        while True:
            pairs = np.zeros((cap, 2), dtype=np.int32)
            used = C.c_size_t(0)
            rc = self.lib.mbl_score(self.ctx, _ptr(m), m.size, n, _ptr(c1), _ptr(c2), _ptr(out), _ptr(pairs), cap, C.byref(used))
            if rc == _ffi.MBL_E_CAPACITY:
                cap = int(used.value) + 16
                continue
            self._check(rc)
            return out, pairs[: used.value]

    # ---- Reporter::writeReadClassification (Reporter.cpp:35-80, printLineage 0) --------------------
    def format_tsv(self, names, results, pairs, header=True, lineage=False) -> str:
        """lineage: --lineage 1 (an extra column with TaxonomyWrapper::taxLineage2 of the classification, Reporter.cpp:37-79)"""
        t = self.db.tax
        rows = []
        if header:
            rows.append("#is_classified\tname\ttaxID\tquery_length\tscore\trank" + ("\tlineage" if lineage else "") + "\ttaxID:match_count\n")
        for i, name in enumerate(names):
            r = results[i]
            score = "%g" % float(r["score"])                     # ostream << float, precision 6 (Q12)
            cls = int(r["classification"])
            if r["is_classified"]:
                b, ln = int(r["taxcnt_begin"]), int(r["taxcnt_len"])
                cnt = "".join(f"{t.original(int(pairs[k, 0]))}:{int(pairs[k, 1])} " for k in range(b, b + ln))
                lin = (t.lineage(cls) + "\t") if lineage else ""
                rows.append(f"1\t{name}\t{t.original(cls)}\t{int(r['query_length'])}\t{score}\t{t.rank_of(cls)}\t{lin}{cnt}\n")
            else:
                rows.append(f"0\t{name}\t{t.original(cls)}\t{int(r['query_length'])}\t{score}\t-\t{'-' + chr(9) if lineage else ''}-\t\n")
        return "".join(rows)

    def classify_files(self, q1: str, q2: str | None = None) -> str:
        """startClassify over whole files -> the text of <jobid>_classifications.tsv."""
        names, b1, o1 = read_fastx(q1)
        b2 = o2 = None
        if q2:
            names2, b2, o2 = read_fastx(q2)
            if len(names2) != len(names):
                raise ValueError("The number of reads in the two files are not equal.")
        res, pairs = self.classify_batch(b1, o1, b2, o2)
        return self.format_tsv(names, res, pairs)
