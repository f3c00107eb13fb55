// Launch interfaces of the classify-path kernels (definitions in k1_extract.cu, k2_sort.cu, k3_index.cu,
// k3_merge.cu, k5_score.cu).  Host code in mbl_api.cu strings them together.
#pragma once
#include "mbl_common.cuh"

namespace mbl {

constexpr uint32_t kItemQueries = 1u << 15; // queries per merge work item (hot tiles are split)

struct TileDirectory {
    Tile* tiles = nullptr;          // [n_tiles]
    uint64_t* cell_k = nullptr;     // [n_cells] k-mer index of the first k-mer ending in the cell
    uint64_t* cell_v = nullptr;     // [n_cells] value of the last k-mer ending before the cell
    uint64_t* jumbo_vals = nullptr; // pre-decoded values of jumbo tiles
    uint64_t n_tiles = 0, n_cells = 0, n_jumbo = 0, n_jumbo_kmers = 0;
    // geometry chosen at load time: nominal tile = tile_cells cells; tiles with more than max_u16 fragments or
    // max_kmers k-mers are jumbo (pre-decoded in HBM)
    uint32_t tile_cells = 4, max_u16 = 0, max_kmers = 0;
    // lowest value bit the query sort must cover so that almost no tile lies inside a single sort bucket (40, 32 or 24;
    // see tile_sort_bit in k3_index.cu) — the merge only needs queries grouped per tile, not fully ordered
    int sort_begin_bit = 24;
    uint64_t n_kmers_decoded = 0;   // number of end flags in the stream (must equal the info count)
    uint32_t* filter = nullptr;     // amino-acid presence filter over every k-mer of the stream (AaFilter), or null
    uint32_t filter_lines = 0;      // 128-byte lines
    int filter_minimizer = 0;
};

struct DeviceTaxonomy {             // reference arrays as stored in taxonomyDB
    const int32_t *D = nullptr, *E = nullptr, *L = nullptr, *H = nullptr, *M = nullptr;
    const int32_t *node_taxid = nullptr, *node_parent = nullptr;
    const uint8_t* node_prune = nullptr;
    const int8_t* node_rank = nullptr;
    const int32_t* taxid2species = nullptr;
    int32_t max_taxid = 0, M_k = 0, eukaryota = 0;
    uint32_t max_nodes = 0;
};

struct ScoreParams {
    float min_score, min_sp_score, tie_ratio;
    int min_cons_cnt, min_cons_cnt_euk, accession_level, denominator, kmer_format;
    int max_codon_shift, dna_shift;   // Taxonomer.cpp:34-42: 1 / 3, or (8 - s) / 3 (8 - s) for syncmer databases
    int force_scratch_dp;      // tests only: skip the register-resident DP fast path
};

// K0 (--mask 1): tantan masking of the resident reads, in place, one warp per read
struct MaskPlan { uint32_t blocks; uint64_t stride, prob_floats, scale_doubles; };
MaskPlan plan_mask(uint32_t n_reads, uint64_t max_len, int sm_count);
void launch_mask(uint8_t* bases, const uint64_t* off, uint32_t n_reads, float mask_prob, const MaskPlan& plan, float* prob_scratch,
                 double* scale_scratch, unsigned long long* counter, cudaStream_t st);

// K1
void launch_read_meta(const uint64_t* off1, const uint64_t* off2, uint32_t n_reads, int32_t* cov1, int32_t* cov2,
                      int32_t* w1, int32_t* w2, uint64_t* slots, uint32_t* quot_cnt, cudaStream_t st);
void launch_extract(int format, const uint8_t* bases1, const uint64_t* off1, const uint8_t* bases2, const uint64_t* off2,
                    uint32_t n_reads, const int32_t* cov1, const int32_t* w1, const int32_t* w2, const uint64_t* slot_off,
                    const uint8_t* base_code, const uint8_t* codon, uint64_t* value, uint64_t* qinfo, uint32_t* slot_idx,
                    unsigned long long* n_valid, int sm_count, cudaStream_t st, AaFilter filter = AaFilter(),
                    unsigned long long* out_cursor = nullptr, uint64_t out_cap = 0, int smer_len = 0);
// filtered extraction packs the surviving metamers from slot 0 upwards in per-warp chunks of kExtractChunk slots (unused chunk
// tails are blank) and never writes at or beyond out_cap (a cursor beyond out_cap tells the host to redo it with more room);
// extract_filtered_capacity = the slots it can need when `pass` of the `slots` reserved slots survive
constexpr uint32_t kExtractChunk = 512;
uint64_t extract_filtered_capacity(uint64_t pass, int sm_count);

// K2 / K4 (radix sorts) and scans
size_t sort_kmers_temp_bytes(size_t n);
void sort_kmers(void* tmp, size_t tmp_bytes, uint64_t* key_a, uint64_t* key_b, uint64_t* val_a, uint64_t* val_b, size_t n,
                int& result_in_b, cudaStream_t st);
// pipeline: the payload is the 64-bit qinfo itself (16 bytes per element through the passes), so the merge reads it as a stream;
// the index-sharded mode sorts (value, 32-bit position in the receive buffer) and gathers qinfo for hits only
void sort_kmers_qinfo(void* tmp, size_t tmp_bytes, uint64_t* key_a, uint64_t* key_b, uint64_t* val_a, uint64_t* val_b, size_t n, int begin_bit,
                      int& result_in_b, cudaStream_t st);
void sort_kmers_idx(void* tmp, size_t tmp_bytes, uint64_t* key_a, uint64_t* key_b, uint32_t* idx_a, uint32_t* idx_b, size_t n, int begin_bit,
                    int& result_in_b, cudaStream_t st);
size_t scan_temp_bytes(size_t n);
void exclusive_sum_u64(void* tmp, size_t tmp_bytes, const uint64_t* in, uint64_t* out, size_t n, cudaStream_t st);
void exclusive_sum_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st);
// full reference order (seqID, species, frame, pos, hamming, dna): sorted copy of `in` in `out`
size_t sort_matches_temp_bytes(size_t n);
// seg_begin / seg_end (optional, [n_reads]): when given and the single-key path is taken, the per-read segments are written by
// the same kernel that gathers the rows; returns true in that case (the caller then skips launch_segments)
bool sort_matches(void* tmp, size_t tmp_bytes, const mbl_match_rec* in, mbl_match_rec* out, size_t n, uint32_t n_reads,
                  int32_t max_taxid, uint32_t max_pos, bool codon_spaced, uint64_t* key_a, uint64_t* key_b, uint32_t* idx_a,
                  uint32_t* idx_b, cudaStream_t st, uint64_t* seg_begin = nullptr, uint64_t* seg_end = nullptr);
size_t order_reads_temp_bytes(size_t n);
const uint32_t* order_reads_by_matches(void* tmp, size_t tmp_bytes, const uint64_t* seg_b, const uint64_t* seg_e, uint32_t n,
                                       uint32_t chunk_reads, uint32_t* key_a, uint32_t* key_b, uint32_t* idx_a, uint32_t* idx_b, cudaStream_t st);
void launch_seq_bounds(const mbl_match_rec* sorted, size_t n, uint32_t chunk_reads, uint32_t n_chunks, uint64_t* bounds, cudaStream_t st);
void launch_segments(const mbl_match_rec* sorted, size_t n, uint32_t n_reads, uint64_t* seg_begin, uint64_t* seg_end, cudaStream_t st);

// K3 directory (load time)
// filter_bits_per_kmer > 0: also build the amino-acid presence filter (about that many bits per k-mer of the stream)
void build_tile_directory(const uint16_t* d_diff, uint64_t n_u16, uint64_t n_kmers, int sm_count, uint32_t tile_cells,
                          cudaStream_t st, TileDirectory& dir, uint64_t base_value = 0, bool holds_db_tail = true,
                          int filter_bits_per_kmer = 0, uint64_t filter_total_kmers = 0, int filter_minimizer = 0);
// words |= other (merging the shards' presence filters)
void launch_filter_or(uint32_t* words, const uint32_t* other, uint64_t n_words, cudaStream_t st);
void free_tile_directory(TileDirectory& dir);

// K3 merge (per batch)
struct MergeArgs {
    const uint16_t* diff;
    const int32_t* info;
    uint32_t info_mask;
    const Tile* tiles;
    uint64_t n_tiles;
    const uint64_t* cell_k;
    const uint64_t* cell_v;
    const uint64_t* jumbo_vals;
    const uint64_t* q_value;        // sorted by amino-acid part
    const uint64_t* q_info;
    const uint32_t* q_idx;          // optional (index-sharded mode): q_info is indexed through q_idx[sorted position]
    uint64_t n_query;               // non-blank
    const int32_t* taxid2species;
    int32_t max_taxid;
    const uint16_t* ham_pair;       // 4096-entry two-codon table in HBM
    const uint8_t* ham_single;      // 64-entry single-codon distances
    uint32_t max_u16, max_kmers, n_buckets;   // shared-memory tile geometry (see smem_layout in k3_merge.cu)
    int kmer_format;
    mbl_match_rec* out;
    uint64_t out_cap;
    unsigned long long* out_count;  // total matches found (may exceed out_cap => overflow)
    unsigned int* error_flag;       // Q2
    // work list
    uint64_t* q_lo;                 // [2 * (n_tiles + 1)]: per tile first query / one past the last query of its prefix range
    int prefix_shift;               // queries are ordered by value >> prefix_shift only
    int cta_threads;                // 512 (2 CTAs per SM, default) or 256 (up to 4 CTAs per SM)
    uint32_t* item_cnt;             // [n_tiles + 1]
    uint32_t* item_off;             // [n_tiles + 1]
    MergeItem* items;
    uint64_t items_cap;
    unsigned int* item_cursor;
    void* scan_tmp;
    size_t scan_tmp_bytes;
};
void launch_merge_plan(const MergeArgs& a, cudaStream_t st);      // partition + work items
void launch_merge(const MergeArgs& a, int sm_count, cudaStream_t st);
size_t merge_smem_bytes(uint32_t max_u16, uint32_t max_kmers, uint32_t n_buckets, int cta_threads);

// K5
struct ScoreArgs {
    const mbl_match_rec* matches;       // sorted
    uint64_t n_match;
    uint32_t read_begin;            // reads [read_begin, read_begin + n_reads) of the sub-batch are scored by one launch
    const uint32_t* read_perm;      // optional: thread i scores read read_perm[read_begin + i] (reads ordered by match count)
    uint32_t n_reads;
    const uint64_t* seg_begin;      // per read (seqID - 1)
    const uint64_t* seg_end;
    const int32_t* cov1;
    const int32_t* cov2;
    const uint32_t* quot_off;       // [n_reads + 1] exclusive scan of per-read quotient-table sizes
    DeviceTaxonomy tax;
    ScoreParams par;
    // scratch, one entry per match
    float* l_score; int32_t* l_start; int32_t* l_ham; int32_t* l_depth; uint32_t* l_smatch; uint8_t* l_conn;
    int32_t* p_start; int32_t* p_end; float* p_score; int32_t* p_ham; int32_t* p_depth; uint32_t* p_smatch; uint32_t* p_ematch;
    int32_t* c_start; int32_t* c_end; float* s_score;
    // flat task lists (first match index of every (species, frame) group / (read, species) group); optional
    const uint32_t* fg_list; uint32_t n_fg; const uint32_t* sp_list; uint32_t n_sp; uint32_t* g_np; uint64_t match_end;
    const uint32_t* fg_order;           // optional: permutation of the frame-group tasks (similar lengths per warp)
    // optional direct maps of the CUDA pipeline (k5_score.cu) in place of binary searches over the task lists:
    const uint32_t* sp_fg;              // [n_sp] index in fg_list of a species group's first frame group
    const uint32_t* read_sp;            // [reads of the sub-batch] index in sp_list of a read's first species group
    float* sp_score;                    // [n_sp] species scores by task index (dense) instead of s_score[first match]
    // scratch, per quotient
    int32_t* q_tax; uint8_t* q_ham; uint8_t* q_has;
    // outputs
    mbl_read_result* results;
    int32_t* taxcnt_pairs;          // 2 x int32 per entry, region of a read starts at quot_off[r]
};
void launch_score(const ScoreArgs& a, cudaStream_t st);
// flat pipeline over the matches [match_begin, a.match_end) of the reads [a.read_begin, +a.n_reads)
struct ScoreFlatScratch {
    uint8_t* flags_fg; uint8_t* flags_sp;       // per match; reused as the length keys of the frame-group ordering
    uint32_t* fg_list; uint32_t* sp_list;
    uint32_t* fg_ord;                           // [2 * matches]: frame-group task order, double buffered
    uint32_t* sp_fg;                            // [matches]: first frame-group task of every species task
    uint32_t* read_sp;                          // [reads of the sub-batch]: first species task of every read
    float* sp_score;                            // [matches]: species scores by task index
    uint32_t* counts; void* cub_tmp; size_t cub_tmp_bytes;
};
size_t score_flat_temp_bytes(size_t n_matches);
void launch_score_flat(ScoreArgs a, uint64_t match_begin, const ScoreFlatScratch& s, cudaStream_t st);
void launch_compact_taxcnt(const mbl_read_result* results, uint32_t n_reads, const uint32_t* quot_off,
                           const int32_t* pairs_in, const uint32_t* out_off, uint32_t pair_base, int32_t* pairs_out,
                           mbl_read_result* results_out, cudaStream_t st);
// taxcnt_begin += delta (mod 2^32) for a range of reads: re-bases a sub-batch's pair offsets inside the batch's pair array
void launch_shift_taxcnt(mbl_read_result* results, uint32_t n_reads, uint32_t delta, cudaStream_t st);
void launch_taxcnt_len(const mbl_read_result* results, uint32_t n_reads, uint32_t* len, cudaStream_t st);

// K6 exchange staging of the index-sharded mode (k6_shard.cu)
constexpr uint32_t kMaxShards = 64;
struct ShardBounds {                // ascending lower bounds of the buckets: shard first values (amino-acid part) / owners' first reads
    uint32_t n;
    uint64_t bound[kMaxShards];
};
struct PushDst {                    // where the rows of bucket b go: rows [begin[b], begin[b+1]) of the send order land at
    uint32_t n;                     // base[b] + row_off[b] (peer-mapped or local memory); total[b] = rows the receiver gets
    uint64_t begin[kMaxShards + 1];
    void* base[kMaxShards];
    uint64_t row_off[kMaxShards];
    uint64_t total[kMaxShards];
};
void push_kmers(const uint32_t* idx, uint64_t n_send, const uint64_t* value, const uint64_t* qinfo, uint64_t seq_add, const PushDst& d,
                cudaStream_t st);
void push_matches(const uint32_t* idx, uint64_t n_send, const mbl_match_rec* in, const PushDst& d, cudaStream_t st);
size_t bucket_sort_temp_bytes(size_t n);
// stable partition by bucket; return the permutation (idx_a or idx_b); d_begin[0..b.n] = bucket starts, d_begin[b.n] = elements kept
const uint32_t* bucket_kmers(void* tmp, size_t tmp_bytes, const uint64_t* value, uint64_t n, const ShardBounds& b, uint8_t* key_a, uint8_t* key_b,
                             uint32_t* idx_a, uint32_t* idx_b, uint64_t* d_begin, cudaStream_t st);
const uint32_t* bucket_matches(void* tmp, size_t tmp_bytes, const mbl_match_rec* m, uint64_t n, const ShardBounds& b, uint8_t* key_a,
                               uint8_t* key_b, uint32_t* idx_a, uint32_t* idx_b, uint64_t* d_begin, cudaStream_t st);
void gather_kmers(const uint32_t* idx, uint64_t n_send, const uint64_t* value, const uint64_t* qinfo, uint64_t seq_add, uint64_t* out_value,
                  uint64_t* out_qinfo, cudaStream_t st);
void gather_matches(const uint32_t* idx, uint64_t n_send, const mbl_match_rec* in, mbl_match_rec* out, cudaStream_t st);
void localize_matches(const mbl_match_rec* in, uint64_t n, uint64_t seq_sub, mbl_match_rec* out, cudaStream_t st);
void launch_iota(uint32_t* idx, uint64_t n, cudaStream_t st);

}  // namespace mbl
