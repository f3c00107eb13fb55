/*
 * metabuli_b200.h — C-ABI of the B200-native `metabuli classify` hot path.
 *
 * The reference (steineggerlab/Metabuli @ 22e7026) has no FFI; the seam this library replaces is the
 * four calls inside Classifier::startClassify (src/commons/Classifier.cpp:105,114,117,118) plus the
 * constructor's loads (Classifier.cpp:6-32).  Each entry point below cites the reference interface it
 * stands in for.  Everything crosses the boundary as plain pointers and sizes; no C++ or torch types.
 *
 * Conventions: one context per process and device, calls serialized by the caller; return 0 = ok,
 * >0 = recoverable (MBL_E_*), <0 = fatal, text via mbl_last_error().  Host buffers are caller-owned.
 * There is no CPU fallback: every function fails with MBL_E_NO_DEVICE when no CUDA device is usable.
 */
#ifndef METABULI_B200_H
#define METABULI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBL_OK                 0
#define MBL_E_MATCH_OVERFLOW   1   /* KmerMatcher::matchKmers returning false (KmerMatcher.cpp:474-476)  */
#define MBL_E_CAPACITY         2   /* caller buffer too small; *n / *used holds the required size          */
#define MBL_E_NO_DEVICE       -1
#define MBL_E_CUDA            -2
#define MBL_E_BAD_ARG         -3
#define MBL_E_BAD_DB          -4   /* Q2: target k-mer with taxid 0 / unmapped species (KmerMatcher.cpp:292-300) */
#define MBL_E_UNSUPPORTED     -5
#define MBL_E_HOST         (-6)   /* host-side failure inside the library (allocation, thread creation); message in mbl_last_error */

typedef struct mbl_ctx mbl_ctx;

/* LocalParameters fields the path reads (src/workflow/classify.cpp:10-37 defaults; db.parameters
 * overrides via loadDbParameters, src/commons/common.cpp:88-133). */
typedef struct {
    int   kmer_format;        /* 1|2 (Kmer_format; classify.cpp:13 default 1)                         */
    int   reduced_aa;         /* must be 0 (ReducedKmerMatcher is out of scope)         */
    int   skip_redundancy;    /* Skip_redundancy: 0 => info & ~(1<<31) (KmerMatcher.cpp:204-205)       */
    int   syncmer;            /* Syncmer: 1 = only closed syncmers are queries (SyncmerScanner.h:9-103), needs    */
    int   smer_len;           /* kmer_format 2; S-mer_len (2..7): paths may skip up to 8 - s codons (Taxonomer.cpp:34-42) */
    int   seq_mode;           /* 1 SE, 2 PE, 3 long (denominator 100/100/1000, Taxonomer.cpp:44-48)    */
    float min_score;          /* --min-score                                                           */
    float min_sp_score;       /* --min-sp-score                                                        */
    float tie_ratio;          /* --tie-ratio (0.95)                                                    */
    int   min_cons_cnt;       /* --min-cons-cnt (4)                                                    */
    int   min_cons_cnt_euk;   /* --min-cons-cnt-euk (9)                                                */
    int   accession_level;    /* 0, or 2 = prune ""/"accession" leaves (Taxonomer.cpp:256-267)         */
    int   device;             /* CUDA device ordinal                                                   */
    int   match_per_kmer;     /* initial match-buffer factor (--match-per-kmer, 4); grows on overflow  */
    int   mask_mode;          /* --mask 1: every uploaded batch is tantan-masked on the device before extraction        */
                              /* (KmerExtractor.cpp:308-314); 0 when off or when the caller masked with mbl_mask_reads  */
    float mask_prob;          /* --mask-prob (0.9)                                                     */
} mbl_config;

/* The on-disk index as the reference reads it (KmerMatcher.cpp:137-139, 212-217): whole files. */
typedef struct {
    const uint16_t* diff_idx; size_t n_u16;     /* <db>/diffIdx                                        */
    const int32_t*  info;     size_t n_kmers;   /* <db>/info                                           */
    const uint64_t* split;    size_t n_split;   /* <db>/split: n_split x {ADkmer,diffOff,infoOff}      */
} mbl_db;

/* The arrays of taxonomyDB as stored (TaxonomyWrapper.cpp:363-421; NcbiTaxonomy.cpp:250-330). */
typedef struct {
    size_t  max_nodes;
    int32_t max_taxid;
    int32_t eukaryota;                 /* TaxonomyWrapper::getEukaryotaTaxID, 0 if absent              */
    const int32_t *D, *E, *L, *H, *M;  /* D[max_taxid+1], E/L[2*max_nodes], H[max_nodes], M[2*max_nodes][M_k] */
    int32_t M_k;
    const int32_t *node_taxid;         /* TaxonNode::taxId per node                                     */
    const int32_t *node_parent;        /* TaxonNode::parentTaxId per node                               */
    const uint8_t *node_prune;         /* 1 when rank is "" or "accession" (Taxonomer.cpp:259), else 0  */
    const int8_t  *node_rank;          /* NcbiRanks index of the node's rank (NcbiTaxonomy.h:52-80), -1 none */
    const int32_t *taxid2species;      /* dense [max_taxid+1], KmerMatcher::loadTaxIdList (:56-120)     */
} mbl_taxonomy;

/* One QuerySplit of reads, SoA (replaces KSeqWrapper entries + vector<Query>, KmerExtractor.cpp:429-481).
 * offsets has n_reads+1 entries into bases; mate 2 arrays are NULL for seq_mode 1/3. */
typedef struct {
    const char*     bases;   const uint64_t* offsets;
    const char*     bases2;  const uint64_t* offsets2;
    uint32_t        n_reads;
} mbl_batch;

/* Query fields the path produces (common.h:94-122). taxcnt_* index the flat (taxid,count) pair array. */
typedef struct {
    int32_t  classification;   /* internal taxid                                                        */
    float    score;
    int32_t  hamming;          /* always 0 (Q6)                                                         */
    int32_t  query_length;     /* covered length, mate1 + mate2 (Q8)                                    */
    uint32_t taxcnt_begin;
    uint32_t taxcnt_len;
    uint8_t  is_classified;
    uint8_t  pad[3];
} mbl_read_result;

/* Match.h:9-26 without the vptr. */
typedef struct {
    uint64_t qinfo;            /* pos[31:0] | seqID[60:32] | frame[63:61] (Kmer.h:11-16)                */
    int32_t  target_id;
    int32_t  species_id;
    uint32_t dna_encoding;
    uint16_t right_end_hamming;
    uint8_t  hamming;
    uint8_t  pad;
} mbl_match_rec;

/* ---- lifetime ----------------------------------------------------------------------------------- */
/* Classifier::Classifier (Classifier.cpp:6-32) */
int  mbl_create(const mbl_config* cfg, mbl_ctx** out);
void mbl_destroy(mbl_ctx* ctx);
const char* mbl_last_error(const mbl_ctx* ctx);

/* loadTaxonomy + KmerMatcher ctor + the per-thread fopen/fread of diffIdx/info (common.cpp:50-86,
 * KmerMatcher.cpp:56-120, 206-217): uploads the index and builds the in-HBM tile directory. */
int  mbl_load_db(mbl_ctx* ctx, const mbl_db* db, const mbl_taxonomy* tax);

/* ---- whole path --------------------------------------------------------------------------------- */
/* One iteration of the QuerySplit loop (Classifier.cpp:81-140): extract, sort, match (with the
 * matchPerKmer overflow retry handled inside), sort matches, score.  out has n_reads entries;
 * taxcnt_pairs receives (taxid,count) int32 pairs in ascending internal taxid order per read. */
int  mbl_classify_batch(mbl_ctx* ctx, const mbl_batch* batch, mbl_read_result* out,
                        int32_t* taxcnt_pairs, size_t taxcnt_cap_pairs, size_t* taxcnt_used_pairs);

/* Same work with the reads already resident in HBM (bench "value" leg). */
int  mbl_upload_batch(mbl_ctx* ctx, const mbl_batch* batch);
int  mbl_classify_resident(mbl_ctx* ctx);
int  mbl_download_results(mbl_ctx* ctx, mbl_read_result* out, int32_t* taxcnt_pairs,
                          size_t taxcnt_cap_pairs, size_t* taxcnt_used_pairs);

/* Streaming over many batches (the QuerySplit loop, Classifier.cpp:81-140, with the next split's reads on their way while
 * the current split is classified): mbl_prefetch_batch starts the asynchronous upload of a batch into a staging buffer on a
 * copy stream and returns; mbl_classify_prefetched makes the staged batch resident, starts the upload of `next` (may be
 * NULL) and classifies.  Host buffers of a prefetched batch must stay valid (and should be pinned) until the
 * mbl_classify_prefetched call that consumes it returns. */
int  mbl_prefetch_batch(mbl_ctx* ctx, const mbl_batch* batch);
int  mbl_classify_prefetched(mbl_ctx* ctx, const mbl_batch* next, mbl_read_result* out, int32_t* taxcnt_pairs,
                             size_t taxcnt_cap_pairs, size_t* taxcnt_used_pairs);

/* ---- stage-level entry points (parity tests against the oracle) -------------------------------- */
/* KmerExtractor::fillQueryKmerBufferParallel[_paired] (KmerExtractor.cpp:83-290): fills every reserved
 * slot; blank slots (N windows) are value = UINT64_MAX, qinfo = 0.  *n = number of slots. */
int  mbl_extract(mbl_ctx* ctx, const mbl_batch* batch, uint64_t* value, uint64_t* qinfo, size_t cap, size_t* n);
/* SORT_PARALLEL(compareQueryKmer) (KmerExtractor.cpp:79): ascending by amino-acid part (value >> 24);
 * order inside an amino-acid group is unspecified (parity-safe). In place. */
int  mbl_sort_kmers(mbl_ctx* ctx, uint64_t* value, uint64_t* qinfo, size_t n);
/* KmerMatcher::matchKmers (KmerMatcher.cpp:123-481): value/qinfo sorted as by mbl_sort_kmers. */
int  mbl_match(mbl_ctx* ctx, const uint64_t* value, const uint64_t* qinfo, size_t n,
               mbl_match_rec* out, size_t cap, size_t* n_match);
/* KmerMatcher::sortMatches (KmerMatcher.cpp:1071-1078, 1149-1166). In place. */
int  mbl_sort_matches(mbl_ctx* ctx, mbl_match_rec* m, size_t n);
/* Classifier::assignTaxonomy (Classifier.cpp:166-208): matches sorted; cov_len* = covered lengths. */
int  mbl_score(mbl_ctx* ctx, const mbl_match_rec* sorted, size_t n_match, uint32_t n_reads,
               const int32_t* cov_len1, const int32_t* cov_len2, mbl_read_result* out,
               int32_t* taxcnt_pairs, size_t taxcnt_cap_pairs, size_t* taxcnt_used_pairs);

/* ---- index-sharded mode ------------------------------------------------------------ */
/* The reference's OpenMP threads all walk one diffIdx (KmerMatcher.cpp:156-217, one DiffIdxSplit checkpoint per thread,
 * Kmer.h:111-119).  Across GPUs the same checkpoints cut the index into contiguous value ranges, one per GPU/context; query
 * metamers travel to the shard that owns their amino-acid part, matches travel back to the rank that owns the read.  The three
 * phases below are what one rank runs around the two exchanges (all-to-all #1: value + qinfo, #2: 24-byte match rows).  The
 * exchange is either the caller's collective (NCCL through torch.distributed, metabuli_b200/sharded.py) fed by mbl_shard_pack_*,
 * or the library's own peer-memory stores (mbl_shard_push_*).  Device pointers returned by a call stay valid until the
 * context's next phase. */
typedef struct {
    uint64_t first_value;      /* amino-acid-group-aligned lower bound of the shard's value range (0 for shard 0,
                                  UINT64_MAX for an empty trailing shard)                                         */
    uint64_t base_value;       /* value the shard's first delta is relative to (DiffIdxSplit::ADkmer of the
                                  preceding k-mer; 0 at the file start)                                           */
    uint64_t diff_begin, diff_end;   /* u16 range of <db>/diffIdx                                                 */
    uint64_t info_begin, info_end;   /* k-mer range of <db>/info                                                  */
    int32_t  holds_db_tail;    /* the shard ends with the numerically last k-mer of the DB (Q1 applies here only) */
    int32_t  pad;
} mbl_shard;
#define MBL_MAX_SHARDS 64
/* Host only (no device needed): cut the index into n_shards value ranges of near-equal diffIdx+info bytes at amino-acid-group
 * starts, from the `split` checkpoints (IndexCreator.cpp:817-872) or, when those are too few, from one scan of the stream. */
int  mbl_plan_shards(const mbl_db* db, uint32_t n_shards, mbl_shard* out);
/* mbl_load_db for one shard: uploads diffIdx[diff_begin, diff_end) and info[info_begin, info_end) only. */
int  mbl_load_db_shard(mbl_ctx* ctx, const mbl_db* db, const mbl_taxonomy* tax, const mbl_shard* shard);
/* Presence filter in sharded mode: a shard's filter covers only its own value range, so phase 1 may use it only after the ranks
 * have merged their parts.  mbl_shard_filter returns this rank's part (device pointer, same size on every rank);
 * mbl_shard_filter_or ORs another rank's part into it (d_other may be NULL) and, with complete != 0, declares the filter
 * whole.  Without this step phase 1 simply sends every metamer. */
int  mbl_shard_filter(mbl_ctx* ctx, void** d_words, uint64_t* n_bytes);
int  mbl_shard_filter_or(mbl_ctx* ctx, const void* d_other, uint64_t n_bytes, int complete);
/* Phase 1, read owner: upload + extract (A0-A3') and bucket the metamers by owning shard.  seq_base = index of the batch's
 * first read among the reads of all ranks (seqIDs are global on the wire); shard_first_value[n_shards] from mbl_plan_shards.
 * -> send_counts[n_shards]: metamers bound for every shard. */
int  mbl_shard_extract(mbl_ctx* ctx, const mbl_batch* batch, uint64_t seq_base, uint32_t n_shards, const uint64_t* shard_first_value,
                       uint64_t* send_counts);
/* Phase 2, shard owner: sort (A4) and merge (A5-A8) the received metamers (device pointers) against the resident shard, then
 * bucket the matches by read owner: owner o holds the reads [owner_first_read[o], owner_first_read[o+1]).
 * -> send_counts[n_owners]: matches bound for every owner. */
int  mbl_shard_match(mbl_ctx* ctx, const uint64_t* d_value, const uint64_t* d_qinfo, uint64_t n, uint32_t n_owners,
                     const uint64_t* owner_first_read, uint64_t* send_counts);
/* Phase 3, read owner: sort (A9) and score (A10-A12) the received matches (device pointer) of the batch given to phase 1;
 * fetch with mbl_download_results. */
int  mbl_shard_score(mbl_ctx* ctx, const mbl_match_rec* d_match, uint64_t n_match);

/* Transport A (collective library): pack the buckets of the last phase 1 / phase 2 into contiguous send buffers (bucket b
 * starts at sum(send_counts[0..b))) for a variable-count all-to-all (NCCL). */
int  mbl_shard_pack_kmers(mbl_ctx* ctx, const uint64_t** d_send_value, const uint64_t** d_send_qinfo);
int  mbl_shard_pack_matches(mbl_ctx* ctx, const mbl_match_rec** d_send_match);

/* Transport B (peer memory, fused gather + all-to-all): every rank owns two receive buffers that its peers' gather kernels
 * store into directly over NVLink.  mbl_shard_recv_buffers allocates them (kmer_rows x 16 B laid out as values | qinfo,
 * match_rows x 24 B) and returns CUDA IPC handles; mbl_shard_attach_peer maps a peer's buffers from its handles (other
 * process) or takes raw device pointers (same process, or the rank itself).  mbl_shard_push_* then replace pack + all-to-all:
 * bucket b's rows land at row dst_row_offset[b] of peer b's buffer; dst_total_rows[b] = rows peer b receives from all ranks
 * (start of its qinfo half).  The caller orders a push before the receiver's next phase with a barrier. */
#define MBL_IPC_HANDLE_BYTES 64
int  mbl_shard_recv_buffers(mbl_ctx* ctx, uint64_t kmer_rows, uint64_t match_rows, void** d_kmers, void** d_matches,
                            uint8_t* handle_kmers, uint8_t* handle_matches);
int  mbl_shard_attach_peer(mbl_ctx* ctx, uint32_t peer, const uint8_t* handle_kmers, const uint8_t* handle_matches,
                           void* raw_kmers, void* raw_matches);
int  mbl_shard_detach_peers(mbl_ctx* ctx);      /* unmap every peer buffer (before the owners re-allocate theirs) */
int  mbl_shard_push_kmers(mbl_ctx* ctx, const uint64_t* dst_row_offset, const uint64_t* dst_total_rows);
int  mbl_shard_push_matches(mbl_ctx* ctx, const uint64_t* dst_row_offset);

/* `--mask 1` (KmerExtractor.cpp:308-314, SeqIterator::maskLowComplexityRegions, SeqIterator.cpp:154-175): tantan's repeat
 * probability per letter of every read; letters at or above mask_prob (and letters that are not nucleotides) become 'N' in
 * place.  Host work (no device needed); call it on a batch before mbl_classify_batch / mbl_prefetch_batch — or set
 * mbl_config.mask_mode and let the device do the same, letter for letter, after the upload (k0_mask.cu). */
int  mbl_mask_reads(char* bases, const uint64_t* offsets, uint32_t n_reads, float mask_prob, int threads);

/* The letters of the resident batch as the extractor sees them (mate 1 or 2; after the device masking when mask_mode is 1):
 * n_bytes = offsets[n_reads] of that mate.  Diagnostics and tests. */
int  mbl_download_reads(mbl_ctx* ctx, int mate, char* out, size_t n_bytes);

/* Pin / unpin a caller buffer (cudaHostRegister) so the copies inside mbl_classify_batch run at PCIe
 * speed; purely an optimisation, pageable buffers work too. */
int  mbl_host_register(void* ptr, size_t bytes);
int  mbl_host_unregister(void* ptr);

/* ---- measurement -------------------------------------------------------------------------------- */
#define MBL_STAGE_H2D      0
#define MBL_STAGE_EXTRACT  1
#define MBL_STAGE_SORT     2
#define MBL_STAGE_MERGE    3   /* the merge kernel alone (CUDA events around its launch)               */
#define MBL_STAGE_MSORT    4
#define MBL_STAGE_SCORE    5
#define MBL_STAGE_D2H      6
#define MBL_STAGE_COUNT    7
typedef struct {
    float    ms[MBL_STAGE_COUNT];     /* CUDA-event time of each stage in the last classify call       */
    float    merge_kernel_ms;         /* the merge kernel launches alone, CUDA events on their stream   */
    uint64_t n_query_kmers;           /* non-blank metamers extracted                                   */
    uint64_t n_matches;
    uint64_t merge_bytes;             /* algorithmic bytes of the merge launches: S_diff+4K+16Nq+24Nm   */
    uint32_t merge_launches;
    uint32_t kernel_launches;         /* launches of this library's own kernels in the last call        */
    uint32_t overflow_retries;
    uint32_t sub_batches;
    float    ms_bucket_kmers;         /* sharded mode: bucketing + packing of the metamers / of the matches    */
    float    ms_bucket_matches;
    uint64_t n_merge_queries;         /* metamers that reached the sort and the merge (after the amino-acid presence filter) */
    float    ms_push_kmers;           /* sharded mode, peer transport: the push kernels alone (stores into the peers' buffers) */
    float    ms_push_matches;
    float    ms_mask;                 /* mask_mode 1: the masking kernel of the last upload              */
    uint32_t reserved0;
} mbl_stats;
int  mbl_get_stats(const mbl_ctx* ctx, mbl_stats* out);

/* DB figures after mbl_load_db: number of tiles, jumbo tiles, k-mers, bytes resident. */
typedef struct { uint64_t n_tiles, n_jumbo, n_kmers, n_u16, hbm_bytes; } mbl_db_info;
int  mbl_get_db_info(const mbl_ctx* ctx, mbl_db_info* out);

#ifdef __cplusplus
}
#endif
#endif
