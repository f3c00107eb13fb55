"""ctypes binding of the C-ABI in include/metabuli_b200.h (libmetabuli_b200.so, built in-tree).

The library is the product: there is no Python or CPU fallback.  Importing this module never touches a
GPU; calling into it without a usable CUDA device returns MBL_E_NO_DEVICE, which surfaces as MblError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MBL_LIB_PATH") or os.path.join(_HERE, "_lib", "libmetabuli_b200.so")   # MBL_LIB_PATH: development builds side by side

MBL_OK = 0
MBL_E_MATCH_OVERFLOW = 1
MBL_E_CAPACITY = 2
MBL_E_NO_DEVICE = -1
MBL_E_CUDA = -2
MBL_E_BAD_ARG = -3
MBL_E_BAD_DB = -4
MBL_E_UNSUPPORTED = -5
MBL_E_HOST = -6

IPC_HANDLE_BYTES = 64
STAGE_NAMES = ["h2d", "extract", "sort", "merge", "match_sort", "score", "d2h"]


class MblError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"metabuli_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("kmer_format", C.c_int), ("reduced_aa", C.c_int), ("skip_redundancy", C.c_int), ("syncmer", C.c_int),
                ("smer_len", C.c_int), ("seq_mode", C.c_int), ("min_score", C.c_float), ("min_sp_score", C.c_float),
                ("tie_ratio", C.c_float), ("min_cons_cnt", C.c_int), ("min_cons_cnt_euk", C.c_int),
                ("accession_level", C.c_int), ("device", C.c_int), ("match_per_kmer", C.c_int),
                ("mask_mode", C.c_int), ("mask_prob", C.c_float)]


class Db(C.Structure):
    _fields_ = [("diff_idx", C.c_void_p), ("n_u16", C.c_size_t), ("info", C.c_void_p), ("n_kmers", C.c_size_t),
                ("split", C.c_void_p), ("n_split", C.c_size_t)]


class Taxonomy(C.Structure):
    _fields_ = [("max_nodes", C.c_size_t), ("max_taxid", C.c_int32), ("eukaryota", C.c_int32),
                ("D", C.c_void_p), ("E", C.c_void_p), ("L", C.c_void_p), ("H", C.c_void_p), ("M", C.c_void_p),
                ("M_k", C.c_int32), ("node_taxid", C.c_void_p), ("node_parent", C.c_void_p), ("node_prune", C.c_void_p),
                ("node_rank", C.c_void_p), ("taxid2species", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("offsets", C.c_void_p), ("bases2", C.c_void_p), ("offsets2", C.c_void_p),
                ("n_reads", C.c_uint32)]


class ReadResult(C.Structure):
    _fields_ = [("classification", C.c_int32), ("score", C.c_float), ("hamming", C.c_int32), ("query_length", C.c_int32),
                ("taxcnt_begin", C.c_uint32), ("taxcnt_len", C.c_uint32), ("is_classified", C.c_uint8), ("pad", C.c_uint8 * 3)]


class Stats(C.Structure):
    _fields_ = [("ms", C.c_float * 7), ("merge_kernel_ms", C.c_float), ("n_query_kmers", C.c_uint64), ("n_matches", C.c_uint64), ("merge_bytes", C.c_uint64),
                ("merge_launches", C.c_uint32), ("kernel_launches", C.c_uint32), ("overflow_retries", C.c_uint32),
                ("sub_batches", C.c_uint32), ("ms_bucket_kmers", C.c_float), ("ms_bucket_matches", C.c_float), ("n_merge_queries", C.c_uint64),
                ("ms_push_kmers", C.c_float), ("ms_push_matches", C.c_float), ("ms_mask", C.c_float), ("reserved0", C.c_uint32)]


class Shard(C.Structure):
    _fields_ = [("first_value", C.c_uint64), ("base_value", C.c_uint64), ("diff_begin", C.c_uint64), ("diff_end", C.c_uint64),
                ("info_begin", C.c_uint64), ("info_end", C.c_uint64), ("holds_db_tail", C.c_int32), ("pad", C.c_int32)]


class DbInfo(C.Structure):
    _fields_ = [("n_tiles", C.c_uint64), ("n_jumbo", C.c_uint64), ("n_kmers", C.c_uint64), ("n_u16", C.c_uint64),
                ("hbm_bytes", C.c_uint64)]


# numpy dtypes matching the C structs
import numpy as np  # noqa: E402

RESULT_DTYPE = np.dtype([("classification", "<i4"), ("score", "<f4"), ("hamming", "<i4"), ("query_length", "<i4"),
                         ("taxcnt_begin", "<u4"), ("taxcnt_len", "<u4"), ("is_classified", "u1"), ("pad", "u1", (3,))])
MATCH_DTYPE = np.dtype([("qinfo", "<u8"), ("target_id", "<i4"), ("species_id", "<i4"), ("dna_encoding", "<u4"),
                        ("right_end_hamming", "<u2"), ("hamming", "u1"), ("pad", "u1")])
assert RESULT_DTYPE.itemsize == C.sizeof(ReadResult) == 28
assert MATCH_DTYPE.itemsize == 24

EXPORTS = ["mbl_create", "mbl_destroy", "mbl_last_error", "mbl_load_db", "mbl_classify_batch", "mbl_upload_batch",
           "mbl_classify_resident", "mbl_download_results", "mbl_prefetch_batch", "mbl_classify_prefetched", "mbl_extract", "mbl_sort_kmers", "mbl_match", "mbl_sort_matches",
           "mbl_score", "mbl_get_stats", "mbl_get_db_info", "mbl_host_register", "mbl_host_unregister",
           "mbl_plan_shards", "mbl_load_db_shard", "mbl_shard_extract", "mbl_shard_match", "mbl_shard_score",
           "mbl_shard_pack_kmers", "mbl_shard_pack_matches", "mbl_shard_recv_buffers", "mbl_shard_attach_peer", "mbl_shard_detach_peers", "mbl_shard_push_kmers",
           "mbl_shard_push_matches", "mbl_shard_filter", "mbl_shard_filter_or", "mbl_mask_reads", "mbl_download_reads"]

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no non-CUDA fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, sz = C.c_void_p, C.c_size_t
    lib.mbl_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.mbl_destroy.argtypes = [vp]
    lib.mbl_destroy.restype = None
    lib.mbl_last_error.argtypes = [vp]
    lib.mbl_last_error.restype = C.c_char_p
    lib.mbl_load_db.argtypes = [vp, C.POINTER(Db), C.POINTER(Taxonomy)]
    lib.mbl_classify_batch.argtypes = [vp, C.POINTER(Batch), vp, vp, sz, C.POINTER(sz)]
    lib.mbl_upload_batch.argtypes = [vp, C.POINTER(Batch)]
    lib.mbl_prefetch_batch.argtypes = [vp, C.POINTER(Batch)]
    lib.mbl_classify_prefetched.argtypes = [vp, C.POINTER(Batch), vp, vp, sz, C.POINTER(sz)]
    lib.mbl_classify_resident.argtypes = [vp]
    lib.mbl_download_results.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    lib.mbl_extract.argtypes = [vp, C.POINTER(Batch), vp, vp, sz, C.POINTER(sz)]
    lib.mbl_sort_kmers.argtypes = [vp, vp, vp, sz]
    lib.mbl_match.argtypes = [vp, vp, vp, sz, vp, sz, C.POINTER(sz)]
    lib.mbl_sort_matches.argtypes = [vp, vp, sz]
    lib.mbl_score.argtypes = [vp, vp, sz, C.c_uint32, vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.mbl_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.mbl_get_db_info.argtypes = [vp, C.POINTER(DbInfo)]
    lib.mbl_plan_shards.argtypes = [C.POINTER(Db), C.c_uint32, C.POINTER(Shard)]
    lib.mbl_load_db_shard.argtypes = [vp, C.POINTER(Db), C.POINTER(Taxonomy), C.POINTER(Shard)]
    lib.mbl_shard_extract.argtypes = [vp, C.POINTER(Batch), C.c_uint64, C.c_uint32, vp, vp]
    lib.mbl_shard_match.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, vp, vp]
    lib.mbl_shard_score.argtypes = [vp, vp, C.c_uint64]
    lib.mbl_shard_pack_kmers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.mbl_shard_pack_matches.argtypes = [vp, C.POINTER(vp)]
    lib.mbl_shard_recv_buffers.argtypes = [vp, C.c_uint64, C.c_uint64, C.POINTER(vp), C.POINTER(vp), vp, vp]
    lib.mbl_shard_attach_peer.argtypes = [vp, C.c_uint32, vp, vp, vp, vp]
    lib.mbl_shard_detach_peers.argtypes = [vp]
    lib.mbl_shard_push_kmers.argtypes = [vp, vp, vp]
    lib.mbl_shard_push_matches.argtypes = [vp, vp]
    lib.mbl_shard_filter.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    lib.mbl_shard_filter_or.argtypes = [vp, vp, C.c_uint64, C.c_int]
    lib.mbl_mask_reads.argtypes = [vp, vp, C.c_uint32, C.c_float, C.c_int]
    lib.mbl_download_reads.argtypes = [vp, C.c_int, vp, sz]
    lib.mbl_host_register.argtypes = [vp, sz]
    lib.mbl_host_unregister.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("mbl_destroy", "mbl_last_error"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
