// C-ABI of the B200 classify path (include/metabuli_b200.h): context, index upload + tile directory,
// and the batch pipeline  K1 extract -> K2 sort -> K3 merge -> K4 match sort -> K5 score
// (reference: Classifier::startClassify, src/commons/Classifier.cpp:44-164).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "mbl_tables.hpp"

using namespace mbl;

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    template <class T>
    T* get(size_t count) {
        size_t bytes = count * sizeof(T) + 64;
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr; cap = 0;
            size_t want = bytes + bytes / 16;
            MBL_CUDA(cudaMalloc(&p, want));
            cap = want;
        }
        return reinterpret_cast<T*>(p);
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct SubBatch { uint32_t r0, r1; uint64_t slots, quots; uint32_t max_pos; uint32_t b0, b1; };   // reads [r0, r1) = blocks [b0, b1)
struct ReadBlock { uint64_t slots, quots; uint32_t max_pos; uint64_t max_len; };
struct BatchSummary { uint32_t n_reads = 0, block_reads = 1; uint64_t max_len = 0; std::vector<ReadBlock> blocks; };
constexpr uint32_t kScoreChunkReads = 1u << 20;   // reads scored per launch (bounds the per-match scratch)

}  // namespace

struct mbl_ctx {
    mbl_config cfg{};
    std::string err;
    int sm_count = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[16] = {};
    // tables
    uint8_t *d_base_code = nullptr, *d_codon = nullptr;
    uint16_t* d_ham_pair = nullptr;
    uint8_t* d_ham_single = nullptr;
    uint32_t tile_cells = 2;            // MBL_TILE_CELLS (measured best on the 8 GiB benchmark index: 3 CTAs per SM)
    int filter_minimizer = 1;           // MBL_FILTER_MINIMIZER=0: filter line from the whole amino-acid part instead of its minimizer
    int filter_bits = 16;               // MBL_FILTER_BITS: bits per index k-mer of the amino-acid presence filter, 0 = no filter
    uint64_t arena_S8 = 0;              // slot stride of the phase-1 arena layout (set by whoever fills it)
    int merge_threads = 512;            // MBL_MERGE_THREADS: 512 (2 CTAs per SM) or 256 (up to 4 CTAs per SM) threads per merge CTA
    bool budget_low = false;            // slots_budget found (almost) no workspace left next to the index
    bool no_probe = false;              // the index-sharded phases work on whole batches: no probe sub-batch
    int force_sort_bit = 0;             // MBL_SORT_BIT: override the load-time choice of TileDirectory::sort_begin_bit
    // test hooks: force the capacity guesses low so that the retry paths run on small inputs (tests/test_gpu_edge_paths.py)
    uint64_t test_match_cap = 0;        // MBL_TEST_MATCH_CAP: first match-buffer capacity (rows) of a merge
    uint64_t test_items_cap = 0;        // MBL_TEST_ITEMS_CAP: first work-list capacity of a merge
    uint64_t test_pack_slots = 0;       // MBL_TEST_PACK_SLOTS: first capacity of the packed (filtered) extraction
    // index
    uint16_t* d_diff = nullptr;
    int32_t* d_info = nullptr;
    uint64_t n_u16 = 0, n_kmers = 0;
    TileDirectory dir;
    DeviceTaxonomy tax;
    std::vector<void*> tax_allocs;
    size_t db_bytes = 0;
    bool db_loaded = false;
    // resident batch
    Buf bases1, bases2, off1, off2;
    uint32_t n_reads = 0;
    bool paired = false;
    std::vector<SubBatch> subs;
    BatchSummary summary, staged_summary;   // per-block totals of the resident / the staged batch (plan_sub_batches)
    bool plan_has_probe = false;            // c->subs still holds the probe cut of an earlier run of the resident batch
    bool probe_first = false, staged_probe_first = false;   // the plan starts with a small probe sub-batch; plan the rest again after it
    // staging copy of the NEXT batch (mbl_prefetch_batch): uploaded on its own stream while the resident batch is classified
    Buf stage_bases1, stage_bases2, stage_off1, stage_off2;
    cudaStream_t copy_st = nullptr;
    cudaEvent_t copy_ev[2] = {};
    bool staged = false, staged_paired = false;
    uint32_t staged_reads = 0;
    std::vector<SubBatch> staged_subs;
    // workspace
    Buf cov1, cov2, w1, w2, slots, slot_off, quot_cnt, quot_off, seg_b, seg_e, res_sub, tax_len, tax_off;
    Buf val_a, val_b, qi_a, qi_b, cub_tmp;
    Buf arena, chunk_bounds, order_keys, g_np, flag_fg, flag_sp, fg_list, sp_list, fg_ord, flat_tmp, sp_fg, read_sp;     // phase 1: k-mer keys/payloads (32 B per slot); phase 2: match sort buffers
    Buf m_raw, m_sorted, key_a, key_b, idx_a, idx_b;
    Buf l_score, l_start, l_ham, l_depth, l_smatch, l_conn, p_start, p_end, p_score, p_ham, p_depth, p_smatch, p_ematch,
        c_start, c_end, s_score;
    Buf q_tax, q_ham, q_has, pairs_raw;
    Buf q_lo, item_cnt, item_off, items, counters;
    Buf mask_prob, mask_scale, mask_counter;            // K0 scratch (mask_mode 1)
    float ms_mask = 0.f;                                // masking kernel of the last upload (kept across the stats reset of a classify call)
    // index-sharded mode (mbl_shard_*): bucket keys / permutations, send buffers, bucket starts
    Buf sh_key_a, sh_key_b, sh_idx_a, sh_idx_b, sh_begin, send_value, send_qinfo, send_match, sh_tmp;
    uint64_t seq_base = 0;              // index of the resident batch's first read among all ranks' reads
    const uint32_t* sh_perm = nullptr;  // bucket permutation of the last mbl_shard_extract / mbl_shard_match
    uint32_t sh_n = 0;                  // its bucket count
    uint64_t sh_S8 = 0;                 // arena layout of the last mbl_shard_extract
    uint64_t sh_begin_h[kMaxShards + 2] = {};   // its bucket starts
    // peer-memory transport: this rank's receive buffers and the peers' (IPC-mapped or same-process pointers)
    void *recv_kmers = nullptr, *recv_matches = nullptr;
    uint64_t recv_kmer_rows = 0, recv_match_rows = 0;
    void* peer_kmers[kMaxShards] = {};
    void* peer_matches[kMaxShards] = {};
    bool peer_opened[kMaxShards][2] = {};
    mbl_shard shard{};                  // the value range this context holds (whole index: first_value 0)
    bool is_shard = false;
    bool filter_complete = false;       // the presence filter covers the whole index (always for mbl_load_db; shards: after the OR)
    // results of the whole batch
    Buf results, pairs, pairs_final;
    uint64_t n_pairs = 0;               // pairs written by this lane
    const void* pairs_out = nullptr;    // what mbl_download_results copies (pairs, or pairs_final after a two-lane run)
    uint64_t n_pairs_total = 0;
    // second pipeline lane: a context with its own stream and workspace that shares the index, the tables, the resident
    // reads and the result array.  Sub-batches alternate between the lanes (two host threads), so the latency-bound
    // scoring kernels of one sub-batch overlap the bandwidth-bound sorts of the other.
    mbl_ctx* shadow = nullptr;
    bool is_shadow = false;
    int pipeline_parts = 2;             // MBL_PIPELINE_PARTS: sub-batches a large batch is cut into when the lanes are on
    int pipeline = 0;                   // MBL_PIPELINE=1 switches the second lane on (measured slower: the lanes contend and the index is streamed twice)
    uint32_t pipeline_min_reads = 1u << 21;   // MBL_PIPELINE_MIN_READS: smaller batches stay on one lane
    double match_ratio = 0.0;   // matches per slot seen so far (sizes the match buffer)
    double shard_match_ratio = 0.0;   // sharded phase 2: matches per received metamer
    double pass_ratio = 0.0;    // slots the packed (filtered) extraction used per reserved slot, seen so far
    mbl_stats stats{};
};

namespace {

// blank tails of the merge kernel's warp-private output chunks (k3_merge.cu: kOutChunk = 1024 slots, <= 4 CTAs x 8 warps per SM)
uint64_t out_slack(const mbl_ctx* c) { return (uint64_t)c->sm_count * 4 * 8 * 1024 + 65536; }

int fail(mbl_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
// no exception may cross the C boundary (a std::bad_alloc from a planning vector, a std::system_error from thread creation
// would abort a ctypes / C caller): every entry point ends with this after its CudaError handler
#define MBL_CATCH_HOST(c)                                                                                       \
    catch (const std::bad_alloc&) { return fail(c, MBL_E_HOST, "host memory allocation failed"); }              \
    catch (const std::exception& e) { return fail(c, MBL_E_HOST, std::string("host error: ") + e.what()); }     \
    catch (...) { return fail(c, MBL_E_HOST, "unknown host error"); }
int fail_cuda(mbl_ctx* c, const CudaError& e) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)e.code, cudaGetErrorString(e.code), e.file, e.line);
    cudaGetLastError();
    return fail(c, MBL_E_CUDA, buf);
}

template <class T>
T* upload(mbl_ctx* c, const T* h, size_t n, size_t pad_elems = 16) {
    T* d = nullptr;
    MBL_CUDA(cudaMalloc(&d, (n + pad_elems) * sizeof(T)));
    MBL_CUDA(cudaMemsetAsync(d, 0, (n + pad_elems) * sizeof(T), c->st));
    if (n) MBL_CUDA(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, c->st));
    return d;
}

struct StageTimer {
    mbl_ctx* c;
    int stage;
    cudaEvent_t a, b;
    StageTimer(mbl_ctx* ctx, int s) : c(ctx), stage(s) {
        a = c->ev[2 * (s % 7)]; b = c->ev[2 * (s % 7) + 1];
        cudaEventRecord(a, c->st);
    }
    void stop() {
        cudaEventRecord(b, c->st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        c->stats.ms[stage] += ms;
    }
};

void free_db(mbl_ctx* c) {
    cudaFree(c->d_diff); cudaFree(c->d_info);
    c->d_diff = nullptr; c->d_info = nullptr;
    free_tile_directory(c->dir);
    for (void* p : c->tax_allocs) cudaFree(p);
    c->tax_allocs.clear();
    c->tax = DeviceTaxonomy();
    c->db_loaded = false;
    c->db_bytes = 0;
}

// host-side planning of sub-batches from the read lengths (QueryIndexer::indexQueryFile analogue, QueryIndexer.cpp:30-147, with
// the HBM budget in place of --max-ram).  The reads are summarised once per batch in blocks of `block_reads` reads (slot and
// quotient totals, largest position); sub-batches are runs of whole blocks, so the plan can be redone for the rest of a batch
// when the first sub-batch has shown how many metamers survive the filter and how many matches a slot yields — the two ratios
// the budget rests on, unknown before the first sub-batch against a new index (a 40 GiB index yields about twice the matches
// per slot of an 8 GiB one) — and a sub-batch that still runs out of memory can be halved.
void summarize_reads(const mbl_batch* b, BatchSummary& sum) {
    const uint32_t n = b->n_reads;
    sum.n_reads = n;
    sum.block_reads = std::max<uint32_t>(1, std::min<uint32_t>(4096, n / 256));
    const size_t nb = ((size_t)n + sum.block_reads - 1) / sum.block_reads;
    sum.blocks.assign(nb, ReadBlock{0, 0, 0, 0});
    auto work = [&](size_t b0, size_t b1) {
        for (size_t k = b0; k < b1; ++k) {
            ReadBlock acc{0, 0, 0, 0};
            const uint32_t r1 = (uint32_t)std::min<uint64_t>(n, (uint64_t)(k + 1) * sum.block_reads);
            for (uint32_t r = (uint32_t)(k * sum.block_reads); r < r1; ++r) {
                const int l1 = (int)(b->offsets[r + 1] - b->offsets[r]);
                acc.max_len = std::max<uint64_t>(acc.max_len, b->offsets[r + 1] - b->offsets[r]);
                if (b->offsets2) acc.max_len = std::max<uint64_t>(acc.max_len, b->offsets2[r + 1] - b->offsets2[r]);
                int w1 = windows_per_frame(l1), c1 = max_covered_length(l1), w2 = 0, c2 = 0;
                if (b->offsets2) {
                    const int l2 = (int)(b->offsets2[r + 1] - b->offsets2[r]);
                    w2 = windows_per_frame(l2); c2 = max_covered_length(l2);
                }
                const bool empty = w1 < 1 || (b->offsets2 && w2 < 1);
                acc.slots += empty ? 0 : 6ull * (uint64_t)(w1 + w2);
                const int ql = c1 + c2;
                acc.quots += ql + 3 > 0 ? (uint64_t)((ql + 3) / 3 + 1) : 1;
                acc.max_pos = std::max(acc.max_pos, (uint32_t)std::max(0, c1 + 3 + c2 + 8));
            }
            sum.blocks[k] = acc;
        }
    };
    const unsigned T = n > (1u << 18) ? 8u : 1u;          // runs while the reads are still on their way to the device
    if (T > 1) {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back(work, nb * t / T, nb * (t + 1) / T);
        work(0, nb / T);
        for (auto& x : th) x.join();
    } else {
        work(0, nb);
    }
    sum.max_len = 0;
    for (const ReadBlock& x : sum.blocks) sum.max_len = std::max(sum.max_len, x.max_len);
}

SubBatch sub_of_blocks(const BatchSummary& sum, size_t b0, size_t b1) {
    SubBatch sb{(uint32_t)std::min<uint64_t>(sum.n_reads, (uint64_t)b0 * sum.block_reads),
                (uint32_t)std::min<uint64_t>(sum.n_reads, (uint64_t)b1 * sum.block_reads), 0, 0, 0, (uint32_t)b0, (uint32_t)b1};
    for (size_t k = b0; k < b1; ++k) { sb.slots += sum.blocks[k].slots; sb.quots += sum.blocks[k].quots; sb.max_pos = std::max(sb.max_pos, sum.blocks[k].max_pos); }
    return sb;
}

// sub-batches over the blocks [first_block, end) appended to subs.  probe: the ratios behind max_slots are guesses (nothing has
// run against this index yet) => a small first sub-batch, after which the caller plans the rest again
void plan_sub_batches(mbl_ctx* c, const BatchSummary& sum, size_t first_block, uint64_t max_slots, bool probe, std::vector<SubBatch>& subs) {
    const size_t nb = sum.blocks.size();
    if (first_block >= nb) return;
    const uint32_t n = sum.n_reads;
    if (first_block == 0 && c->pipeline && n >= c->pipeline_min_reads && n >= 16) {
        // large batches are cut in (at least) pipeline_parts pieces so that the two lanes overlap (mbl_classify_resident); each
        // lane then owns half of the workspace budget
        const unsigned P = (unsigned)std::max(2, c->pipeline_parts);
        std::vector<SubBatch> h;
        bool fits = true;
        for (unsigned p = 0; p < P; ++p) {
            const size_t b0 = nb * p / P, b1 = nb * (p + 1) / P;
            if (b1 == b0) continue;
            h.push_back(sub_of_blocks(sum, b0, b1));
            fits = fits && h.back().slots <= max_slots / 2 && h.back().quots <= 0xF0000000ull;
        }
        if (fits) { for (const SubBatch& d : h) subs.push_back(d); return; }
        max_slots /= 2;
    }
    size_t b0 = first_block;
    if (probe) {
        const size_t pb = std::max<size_t>(1, (size_t)(262144 / sum.block_reads));
        if (nb - first_block > 2 * pb) {
            subs.push_back(sub_of_blocks(sum, first_block, first_block + pb));
            b0 = first_block + pb;
        }
    }
    uint64_t slots = 0, quots = 0;
    size_t start = b0;
    for (size_t k = b0; k < nb; ++k) {
        const ReadBlock& rb = sum.blocks[k];
        if (k > start && (slots + rb.slots > max_slots || quots + rb.quots > 0xF0000000ull)) {
            subs.push_back(sub_of_blocks(sum, start, k));
            start = k; slots = 0; quots = 0;
        }
        slots += rb.slots; quots += rb.quots;
    }
    subs.push_back(sub_of_blocks(sum, start, nb));
}

uint64_t slots_budget(mbl_ctx* c) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // what the workspace already holds is reusable
    size_t held = 0;
    for (Buf* b : {&c->arena, &c->m_raw, &c->l_score, &c->l_start, &c->l_ham, &c->l_depth, &c->l_smatch, &c->l_conn, &c->p_start, &c->p_end,
                   &c->p_score, &c->p_ham, &c->p_depth, &c->p_smatch, &c->p_ematch, &c->c_start, &c->c_end, &c->s_score, &c->cub_tmp,
                   &c->q_tax, &c->q_ham, &c->q_has, &c->pairs_raw, &c->seg_b, &c->seg_e, &c->res_sub})
        held += b->cap;
    // per slot: 32 B keys+payloads (arena; reused by the match sort: 48 B per match) + 24 B per raw match + ~4 B per-read tables;
    // fixed: scoring scratch of one chunk + results
    const double r = std::max(0.5, c->match_ratio * 1.3);
    // with the presence filter the phase-1 arena only holds the survivors (the guess of run_sub_batch)
    const double kept = (c->dir.filter && c->filter_complete) ? std::min(1.0, c->pass_ratio > 0 ? 1.3 * c->pass_ratio + 0.05 : 0.45) : 1.0;
    const double per_slot = std::max(32.0 * kept, 48.0 * r) + 24.0 * r + 4.0;
    // fixed part: scoring scratch of one chunk, results, CUB scratch — 10 GB on a B200, never more than 6 % of the device
    double budget = 0.88 * (double)(free_b + held) - std::min(10.0e9, 0.06 * (double)total_b);
    uint64_t s = budget > 0 ? (uint64_t)(budget / per_slot) : 0;
    s = std::min<uint64_t>(s, (uint64_t)(3.9e9 / r));            // 32-bit match permutation
    s = std::min<uint64_t>(s, 4000000000ull);                    // 32-bit slot index (K2 payload)
    c->budget_low = s < (1ull << 22);                          // a few thousand reads per sub-batch: the run would look hung
    s = std::max<uint64_t>(s, 1ull << 16);
    return s;
}

// ---- the pipeline over one sub-batch of resident reads, in three stages -------------------------------------------------
// (the index-sharded mode runs the same stages with an exchange between them, see mbl_shard_* below)

// K1: per-read metadata + metamer extraction into the phase-1 arena (value A | value B | qinfo A | qinfo B)
// use_filter: metamers whose amino-acid part is not in the index are dropped and the survivors packed from slot 0 (K1 + filter)
// filter_pass: expected number of survivors (sizes the packed buffers; a too small guess is detected and redone by the caller)
void stage_extract(mbl_ctx* c, const SubBatch& sb, bool use_filter, uint64_t filter_pass = 0, uint64_t forced_cap = 0) {
    cudaStream_t st = c->st;
    const uint32_t n = sb.r1 - sb.r0;
    const uint64_t S = sb.slots;
    use_filter = use_filter && c->dir.filter != nullptr;
    const uint64_t S8 = ((use_filter ? (forced_cap ? forced_cap : extract_filtered_capacity(std::min(filter_pass, S), c->sm_count)) : S) + 31) & ~31ull;
    c->arena_S8 = S8;
    const uint8_t* bases1 = (const uint8_t*)c->bases1.p;
    const uint8_t* bases2 = c->paired ? (const uint8_t*)c->bases2.p : nullptr;
    const uint64_t* off1 = (const uint64_t*)c->off1.p + sb.r0;
    const uint64_t* off2 = c->paired ? (const uint64_t*)c->off2.p + sb.r0 : nullptr;
    int32_t *cov1 = c->cov1.get<int32_t>(n), *cov2 = c->cov2.get<int32_t>(n), *w1 = c->w1.get<int32_t>(n), *w2 = c->w2.get<int32_t>(n);
    uint64_t *slots = c->slots.get<uint64_t>(n + 1), *slot_off = c->slot_off.get<uint64_t>(n + 1);
    uint32_t *quot_cnt = c->quot_cnt.get<uint32_t>(n + 1), *quot_off = c->quot_off.get<uint32_t>(n + 1);
    const size_t scan_bytes = std::max(scan_temp_bytes(n + 1), scan_temp_bytes(c->dir.n_tiles + 2));
    const size_t sortk_bytes = sort_kmers_temp_bytes(S);
    unsigned long long* counters = c->counters.get<unsigned long long>(8);   // [0] n_valid [1] reserved [2] matches [3] err|cursor
    {
        StageTimer t(c, MBL_STAGE_EXTRACT);
        MBL_CUDA(cudaMemsetAsync(counters, 0, 64, st));
        MBL_CUDA(cudaMemsetAsync(slots + n, 0, 8, st));
        MBL_CUDA(cudaMemsetAsync(quot_cnt + n, 0, 4, st));
        launch_read_meta(off1, off2, n, cov1, cov2, w1, w2, slots, quot_cnt, st);
        void* tmp = c->cub_tmp.get<uint8_t>(std::max(scan_bytes, sortk_bytes));
        exclusive_sum_u64(tmp, c->cub_tmp.cap, slots, slot_off, n + 1, st);
        exclusive_sum_u32(tmp, c->cub_tmp.cap, quot_cnt, quot_off, n + 1, st);
        // phase-1 layout of the arena (32 B per slot): value A | value B | qinfo A | qinfo B (the sort's second buffers)
        uint64_t* ar = c->arena.get<uint64_t>(4 * S8 + 64);
        uint64_t *va = ar, *qa = ar + 2 * S8;
        AaFilter flt;
        if (use_filter) { flt.words = c->dir.filter; flt.n_lines = c->dir.filter_lines; flt.minimizer = c->dir.filter_minimizer; }
        launch_extract(c->cfg.kmer_format, bases1, off1, bases2, off2, n, cov1, w1, w2, slot_off, c->d_base_code, c->d_codon,
                       va, qa, nullptr, counters, c->sm_count, st, flt, counters + 4, S8, c->cfg.syncmer ? c->cfg.smer_len : 0);
        c->stats.kernel_launches += 2;
        t.stop();
    }
}

// K2 + K3: sort the S slots in the phase-1 arena (keys in `value A`, slot indices in `slot idx A`) and merge them against the
// resident index; q_info is the array the slot indices point into.  -> rows written to m_raw (blank tails included), matches
// ratio: matches per cap_basis slot seen so far on this kind of input (sizes the match buffer; updated)
int stage_sort_merge(mbl_ctx* c, uint64_t S, uint64_t cap_basis, double* ratio, const uint64_t* q_info, bool count_valid_on_device,
                     uint64_t* reserved_out, uint64_t* n_match_out) {
    cudaStream_t st = c->st;
    const uint64_t S8 = c->arena_S8;
    const size_t scan_bytes = scan_temp_bytes(c->dir.n_tiles + 2);
    const size_t sortk_bytes = sort_kmers_temp_bytes(S);
    unsigned long long* counters = c->counters.get<unsigned long long>(8);
    uint64_t *qv = nullptr;
    uint32_t* qidx = nullptr;
    bool direct = false;
    {
        StageTimer t(c, MBL_STAGE_SORT);
        c->cub_tmp.get<uint8_t>(std::max(scan_bytes, sortk_bytes));
        uint64_t* ar = (uint64_t*)c->arena.p;
        uint64_t *va = ar, *vb = ar + S8;
        uint32_t *ia = reinterpret_cast<uint32_t*>(ar + 3 * S8), *ib = ia + S8;
        int in_b = 0;
        // qinfo travels through the sort with the value when it sits in the arena (i.e. not for the receive buffer of the sharded
        // mode, which is sorted as (value, position) pairs); its second buffer is the region of the two slot-index arrays
        direct = q_info == ar + 2 * S8;
        if (direct) {
            uint64_t *qa = ar + 2 * S8, *qb = ar + 3 * S8;
            if (S) sort_kmers_qinfo(c->cub_tmp.p, c->cub_tmp.cap, va, vb, qa, qb, S, c->dir.sort_begin_bit, in_b, st);
            q_info = in_b ? qb : qa;
        } else if (S) {
            sort_kmers_idx(c->cub_tmp.p, c->cub_tmp.cap, va, vb, ia, ib, S, c->dir.sort_begin_bit, in_b, st);
        }
        qv = in_b ? vb : va;
        qidx = direct ? nullptr : (in_b ? ib : ia);
        t.stop();
    }
    const uint64_t* qi = q_info;
    unsigned long long h_cnt[4] = {0, 0, 0, 0};
    uint64_t n_query = S;                          // exchange buffers hold no blanks
    if (count_valid_on_device) {
        MBL_CUDA(cudaMemcpyAsync(h_cnt, counters, 8, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        n_query = h_cnt[0];
    }
    c->stats.n_merge_queries += n_query;

    // ---- K3 ------------------------------------------------------------------------------------------
    uint64_t cap = (uint64_t)((double)cap_basis * std::max(*ratio * 1.25, 0.125 * (double)std::max(1, c->cfg.match_per_kmer))) + out_slack(c);
    if (c->test_match_cap) cap = c->test_match_cap;
    uint64_t reserved = 0, n_match = 0;
    MergeArgs ma{};
    ma.diff = c->d_diff; ma.info = c->d_info;
    ma.info_mask = ~((uint32_t)(c->cfg.skip_redundancy == 0) << 31);
    ma.tiles = c->dir.tiles; ma.n_tiles = c->dir.n_tiles; ma.cell_k = c->dir.cell_k; ma.cell_v = c->dir.cell_v;
    ma.jumbo_vals = c->dir.jumbo_vals;
    ma.q_value = qv; ma.q_info = qi; ma.q_idx = qidx; ma.n_query = n_query;
    ma.taxid2species = c->tax.taxid2species; ma.max_taxid = c->tax.max_taxid;
    ma.ham_pair = c->d_ham_pair; ma.kmer_format = c->cfg.kmer_format;
    ma.ham_single = c->d_ham_single; ma.max_u16 = c->dir.max_u16 + 16; ma.max_kmers = c->dir.max_kmers;
    ma.n_buckets = 2; while (ma.n_buckets <= ma.max_kmers) ma.n_buckets <<= 1;    // hash table load <= 50 %
    ma.out_count = counters + 1;
    ma.error_flag = reinterpret_cast<unsigned int*>(counters + 3);
    ma.item_cursor = reinterpret_cast<unsigned int*>(counters + 3) + 1;
    ma.q_lo = c->q_lo.get<uint64_t>(2 * c->dir.n_tiles + 2);
    ma.prefix_shift = c->dir.sort_begin_bit;
    ma.cta_threads = c->merge_threads;
    ma.item_cnt = c->item_cnt.get<uint32_t>(c->dir.n_tiles + 2);
    ma.item_off = c->item_off.get<uint32_t>(c->dir.n_tiles + 2);
    ma.items_cap = c->dir.n_tiles + n_query / kItemQueries + 2;
    if (c->test_items_cap) ma.items_cap = c->test_items_cap;
    ma.items = c->items.get<MergeItem>(ma.items_cap);
    ma.scan_tmp = c->cub_tmp.p; ma.scan_tmp_bytes = c->cub_tmp.cap;
    for (int attempt = 0;; ++attempt) {
        ma.out = c->m_raw.get<mbl_match_rec>(cap);
        ma.out_cap = cap;
        MBL_CUDA(cudaMemsetAsync(counters + 1, 0, 24, st));
        {
            StageTimer t(c, MBL_STAGE_MERGE);
            if (n_query && c->dir.n_tiles) {
                launch_merge_plan(ma, st);
                cudaEventRecord(c->ev[14], st);
                launch_merge(ma, c->sm_count, st);
                cudaEventRecord(c->ev[15], st);
                cudaEventSynchronize(c->ev[15]);
                float kms = 0;
                cudaEventElapsedTime(&kms, c->ev[14], c->ev[15]);
                c->stats.merge_kernel_ms += kms;
                c->stats.kernel_launches += 4;
                c->stats.merge_launches += 1;
            }
            t.stop();
        }
        MBL_CUDA(cudaMemcpyAsync(h_cnt, counters, 32, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        MBL_CUDA(cudaGetLastError());
        reserved = h_cnt[1]; n_match = h_cnt[2];
        if ((uint32_t)h_cnt[3] & 1u) return fail(c, MBL_E_BAD_DB, "target k-mer with taxid 0 or unmapped species (reference exits, KmerMatcher.cpp:292-300)");
        if ((uint32_t)h_cnt[3] & 2u) {          // overlapping prefix ranges produced more work items than planned for
            uint32_t need = 0;
            MBL_CUDA(cudaMemcpy(&need, ma.item_off + c->dir.n_tiles, 4, cudaMemcpyDeviceToHost));
            ma.items_cap = (uint64_t)need + 2;
            ma.items = c->items.get<MergeItem>(ma.items_cap);
            c->stats.overflow_retries += 1;
            if (attempt > 4) return fail(c, MBL_E_MATCH_OVERFLOW, "merge work list overflow persists");
            continue;
        }
        if (reserved <= cap) break;
        // Classifier.cpp:127-130: the reference bumps matchPerKmer and restarts; here only the merge is redone
        c->stats.overflow_retries += 1;
        cap = reserved + reserved / 8 + (1u << 16);
        if (attempt > 4) return fail(c, MBL_E_MATCH_OVERFLOW, "match buffer overflow persists");
    }
    c->stats.n_matches += n_match;
    c->stats.merge_bytes += 2 * c->n_u16 + 4 * c->n_kmers + 16 * n_query + 24 * n_match;
    if (cap_basis) *ratio = std::max(*ratio, (double)reserved / (double)cap_basis);
    if (reserved >= (1ull << 32)) return fail(c, MBL_E_UNSUPPORTED, "more than 2^32 matches in one sub-batch");
    *reserved_out = reserved; *n_match_out = n_match;
    return MBL_OK;
}

// K4 + K5: order the M rows in m_raw (blank rows allowed) and score the sub-batch's reads
int stage_sort_score(mbl_ctx* c, const SubBatch& sb, uint64_t M) {
    cudaStream_t st = c->st;
    const uint32_t n = sb.r1 - sb.r0;
    const size_t scan_bytes = std::max(scan_temp_bytes(n + 1), scan_temp_bytes(c->dir.n_tiles + 2));
    int32_t *cov1 = (int32_t*)c->cov1.p, *cov2 = (int32_t*)c->cov2.p;
    uint32_t* quot_off = (uint32_t*)c->quot_off.p;
    // the k-mer buffers are dead now: phase-2 layout of the arena = sorted matches | key A | key B | idx A | idx B
    const uint64_t M8 = (M + 32) & ~31ull;
    uint8_t* ar2 = c->arena.get<uint8_t>(48 * M8 + 256);
    mbl_match_rec* sorted = reinterpret_cast<mbl_match_rec*>(ar2);
    uint64_t* key_a = reinterpret_cast<uint64_t*>(ar2 + 24 * M8);
    uint64_t* key_b = key_a + M8;
    uint32_t* idx_a = reinterpret_cast<uint32_t*>(key_b + M8);
    uint32_t* idx_b = idx_a + M8;
    uint64_t *seg_b = c->seg_b.get<uint64_t>(n + 1), *seg_e = c->seg_e.get<uint64_t>(n + 1);
    {
        StageTimer t(c, MBL_STAGE_MSORT);
        const size_t sm_bytes = sort_matches_temp_bytes(M);
        void* tmp = c->cub_tmp.get<uint8_t>(std::max(sm_bytes, scan_bytes));
        const bool have_segments = sort_matches(tmp, c->cub_tmp.cap, (const mbl_match_rec*)c->m_raw.p, sorted, M, n, c->tax.max_taxid, sb.max_pos,
                                                true, key_a, key_b, idx_a, idx_b, st, seg_b, seg_e);
        if (!have_segments) launch_segments(sorted, M, n, seg_b, seg_e, st);
        c->stats.kernel_launches += M ? 4 : 0;
        t.stop();
    }
    // ---- K5 ------------------------------------------------------------------------------------------
    {
        StageTimer t(c, MBL_STAGE_SCORE);
        ScoreArgs sa{};
        sa.matches = sorted; sa.n_match = M; sa.seg_begin = seg_b; sa.seg_end = seg_e;
        sa.cov1 = cov1; sa.cov2 = cov2; sa.quot_off = quot_off;
        sa.tax = c->tax;
        sa.par.min_score = c->cfg.min_score; sa.par.min_sp_score = c->cfg.min_sp_score; sa.par.tie_ratio = c->cfg.tie_ratio;
        sa.par.min_cons_cnt = c->cfg.min_cons_cnt; sa.par.min_cons_cnt_euk = c->cfg.min_cons_cnt_euk;
        sa.par.accession_level = c->cfg.accession_level;
        sa.par.denominator = (c->cfg.seq_mode == 1 || c->cfg.seq_mode == 2) ? 100 : 1000;      // Taxonomer.cpp:44-48
        sa.par.kmer_format = c->cfg.kmer_format;
        sa.par.max_codon_shift = c->cfg.syncmer ? 8 - c->cfg.smer_len : 1;               // Taxonomer.cpp:34-42
        sa.par.dna_shift = 3 * sa.par.max_codon_shift;
        sa.q_tax = c->q_tax.get<int32_t>(sb.quots + 1); sa.q_ham = c->q_ham.get<uint8_t>(sb.quots + 1); sa.q_has = c->q_has.get<uint8_t>(sb.quots + 1);
        sa.taxcnt_pairs = c->pairs_raw.get<int32_t>(2 * (sb.quots + 1));
        sa.results = c->res_sub.get<mbl_read_result>(n);
        // reads are scored in chunks so the per-match scratch only has to cover one chunk
        const uint32_t chunk_reads = kScoreChunkReads;
        const uint32_t n_chunks = (n + chunk_reads - 1) / chunk_reads;
        uint64_t* d_bounds = c->chunk_bounds.get<uint64_t>(n_chunks + 2);
        std::vector<uint64_t> bounds(n_chunks + 1, 0);
        launch_seq_bounds(sorted, M, chunk_reads, n_chunks, d_bounds, st);
        MBL_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds, 8 * (size_t)(n_chunks + 1), cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        uint64_t max_chunk = 0;
        for (uint32_t k = 0; k < n_chunks; ++k) max_chunk = std::max(max_chunk, bounds[k + 1] - bounds[k]);
        // inside a chunk, threads take the reads in order of match count (less divergence between the lanes of a warp)
        {
            void* tmp = c->cub_tmp.get<uint8_t>(std::max(order_reads_temp_bytes(n), scan_bytes));
            uint32_t* ok = c->order_keys.get<uint32_t>(4 * (size_t)(n + 1));
            sa.read_perm = order_reads_by_matches(tmp, c->cub_tmp.cap, seg_b, seg_e, n, chunk_reads, ok, ok + (n + 1), ok + 2 * (size_t)(n + 1),
                                                  ok + 3 * (size_t)(n + 1), st);
            c->stats.kernel_launches += 1;
        }
        const size_t Mp = max_chunk + 1;
        float *l_score = c->l_score.get<float>(Mp), *p_score = c->p_score.get<float>(Mp), *s_score = c->s_score.get<float>(Mp);
        int32_t *l_start = c->l_start.get<int32_t>(Mp), *l_ham = c->l_ham.get<int32_t>(Mp), *l_depth = c->l_depth.get<int32_t>(Mp);
        uint32_t *l_smatch = c->l_smatch.get<uint32_t>(Mp), *p_smatch = c->p_smatch.get<uint32_t>(Mp), *p_ematch = c->p_ematch.get<uint32_t>(Mp);
        uint8_t* l_conn = c->l_conn.get<uint8_t>(Mp);
        int32_t *p_start = c->p_start.get<int32_t>(Mp), *p_end = c->p_end.get<int32_t>(Mp), *p_ham = c->p_ham.get<int32_t>(Mp),
                *p_depth = c->p_depth.get<int32_t>(Mp), *c_start = c->c_start.get<int32_t>(Mp), *c_end = c->c_end.get<int32_t>(Mp);
        uint32_t* g_np = c->g_np.get<uint32_t>(Mp);
        ScoreFlatScratch flat{};
        flat.flags_fg = c->flag_fg.get<uint8_t>(Mp); flat.flags_sp = c->flag_sp.get<uint8_t>(Mp);
        flat.fg_list = c->fg_list.get<uint32_t>(Mp); flat.sp_list = c->sp_list.get<uint32_t>(Mp);
        flat.fg_ord = c->fg_ord.get<uint32_t>(2 * Mp);
        flat.sp_fg = c->sp_fg.get<uint32_t>(Mp);
        flat.read_sp = c->read_sp.get<uint32_t>(n + 1);
        flat.sp_score = s_score;                    // the per-match array, indexed by species task (the flat passes use nothing else of it)
        flat.counts = reinterpret_cast<uint32_t*>(c->counters.get<unsigned long long>(8)) + 12;
        flat.cub_tmp = c->flat_tmp.get<uint8_t>(score_flat_temp_bytes(Mp)); flat.cub_tmp_bytes = c->flat_tmp.cap;
        for (uint32_t k = 0; k < n_chunks; ++k) {
            const uint64_t f = bounds[k];                 // scratch is indexed by (match index - f)
            sa.read_begin = k * chunk_reads;
            sa.n_reads = std::min(chunk_reads, n - sa.read_begin);
            sa.l_score = l_score - f; sa.l_start = l_start - f; sa.l_ham = l_ham - f; sa.l_depth = l_depth - f; sa.l_smatch = l_smatch - f;
            sa.l_conn = l_conn - f; sa.p_start = p_start - f; sa.p_end = p_end - f; sa.p_score = p_score - f; sa.p_ham = p_ham - f;
            sa.p_depth = p_depth - f; sa.p_smatch = p_smatch - f; sa.p_ematch = p_ematch - f; sa.c_start = c_start - f; sa.c_end = c_end - f;
            sa.s_score = s_score - f;
            sa.g_np = g_np - f;
            sa.match_end = bounds[k + 1];
            launch_score_flat(sa, f, flat, st);
            c->stats.kernel_launches += 6;
        }
        // compact the (taxid,count) lists behind the pairs of earlier sub-batches
        uint32_t *tl = c->tax_len.get<uint32_t>(n + 1), *to = c->tax_off.get<uint32_t>(n + 1);
        launch_taxcnt_len(sa.results, n, tl, st);
        exclusive_sum_u32(c->cub_tmp.p, c->cub_tmp.cap, tl, to, n + 1, st);
        uint32_t total = 0;
        MBL_CUDA(cudaMemcpyAsync(&total, to + n, 4, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        // grow the batch-level pair array, keeping what is there
        if ((c->n_pairs + total) * 8 + 64 > c->pairs.cap) {
            Buf nb;
            int32_t* np = nb.get<int32_t>(2 * (c->n_pairs + total) + 2 * (size_t)c->n_reads);
            if (c->n_pairs) MBL_CUDA(cudaMemcpyAsync(np, c->pairs.p, c->n_pairs * 8, cudaMemcpyDeviceToDevice, st));
            MBL_CUDA(cudaStreamSynchronize(st));
            c->pairs.release();
            c->pairs = nb;
        }
        launch_compact_taxcnt(sa.results, n, quot_off, sa.taxcnt_pairs, to, (uint32_t)c->n_pairs, (int32_t*)c->pairs.p + 2 * c->n_pairs,
                              (mbl_read_result*)c->results.p + sb.r0, st);
        c->stats.kernel_launches += 3;
        t.stop();
        c->n_pairs += total;
    }
    MBL_CUDA(cudaGetLastError());
    return MBL_OK;
}

// K1 through the presence filter: -> slots the packed extraction handed out (what the sort / the bucketing walks).  The packed
// buffers are sized from the survivor fraction seen so far (40 % before the first batch); K1 never writes beyond them, and a
// cursor past the end means the guess was too small: redo with room for every slot.
int extract_filtered(mbl_ctx* c, const SubBatch& sb, uint64_t* n_slots_used) {
    unsigned long long h[6] = {0, 0, 0, 0, 0, 0};
    uint64_t guess = (uint64_t)((double)sb.slots * std::min(1.0, c->pass_ratio > 0 ? 1.3 * c->pass_ratio + 0.02 : 0.4)) + 4096;
    for (int attempt = 0;; ++attempt) {
        stage_extract(c, sb, true, guess, attempt == 0 ? c->test_pack_slots : 0);
        MBL_CUDA(cudaMemcpyAsync(h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost, c->st));
        MBL_CUDA(cudaStreamSynchronize(c->st));
        if (h[4] <= c->arena_S8 || attempt > 0) break;
        c->stats.overflow_retries += 1;
        // the cursor kept counting past the capacity, so it IS the room the packed output needs (chunk tails included); the same
        // reads give the same survivors, only the chunk hand-out order differs between launches: a margin covers that
        guess = std::min<uint64_t>(sb.slots, h[4] + h[4] / 8 + 65536);
    }
    if (h[4] > c->arena_S8) return fail(c, MBL_E_CUDA, "internal: packed extraction overflow");
    *n_slots_used = h[4];
    c->stats.n_query_kmers += h[5];             // valid metamers before the filter
    if (sb.slots) c->pass_ratio = std::max(c->pass_ratio, (double)h[4] / (double)sb.slots);
    return MBL_OK;
}

// front half of a sub-batch: K1 (+ presence filter), K2, K3 -> rows in m_raw
int sub_front(mbl_ctx* c, const SubBatch& sb, uint64_t* reserved_out) {
    const bool filtered = c->dir.filter != nullptr && c->filter_complete;
    uint64_t n_sort = sb.slots;
    if (filtered) {
        int rc = extract_filtered(c, sb, &n_sort);
        if (rc != MBL_OK) return rc;
    } else {
        stage_extract(c, sb, false);
    }
    uint64_t n_match = 0;
    const uint64_t nq_before = c->stats.n_merge_queries;
    int rc = stage_sort_merge(c, n_sort, sb.slots, &c->match_ratio, (const uint64_t*)c->arena.p + 2 * c->arena_S8, true, reserved_out, &n_match);
    if (!filtered) c->stats.n_query_kmers += c->stats.n_merge_queries - nq_before;
    return rc;
}
// back half: K4, K5
int sub_back(mbl_ctx* c, const SubBatch& sb, uint64_t reserved) { return stage_sort_score(c, sb, reserved); }

int run_sub_batch(mbl_ctx* c, const SubBatch& sb) {
    uint64_t reserved = 0;
    int rc = sub_front(c, sb, &reserved);
    if (rc != MBL_OK) return rc;
    return sub_back(c, sb, reserved);
}

}  // namespace

// =====================================================================================================
// mask_mode 1: the batch that has just become resident is masked in place on the context's stream (K0), once per upload
void mask_resident(mbl_ctx* c) {
    c->ms_mask = 0.f;
    if (!c->cfg.mask_mode || !c->n_reads) return;
    const MaskPlan plan = plan_mask(c->n_reads, c->summary.max_len, c->sm_count);
    if (!plan.blocks) return;
    float* prob = c->mask_prob.get<float>(plan.prob_floats);
    double* scale = c->mask_scale.get<double>(plan.scale_doubles);
    unsigned long long* counter = c->mask_counter.get<unsigned long long>(2);
    cudaEvent_t a = c->ev[14], b = c->ev[15];
    MBL_CUDA(cudaEventRecord(a, c->st));
    launch_mask((uint8_t*)c->bases1.p, (const uint64_t*)c->off1.p, c->n_reads, c->cfg.mask_prob, plan, prob, scale, counter, c->st);
    if (c->paired) launch_mask((uint8_t*)c->bases2.p, (const uint64_t*)c->off2.p, c->n_reads, c->cfg.mask_prob, plan, prob, scale, counter + 1, c->st);
    MBL_CUDA(cudaGetLastError());
    MBL_CUDA(cudaEventRecord(b, c->st));
    MBL_CUDA(cudaEventSynchronize(b));
    MBL_CUDA(cudaEventElapsedTime(&c->ms_mask, a, b));
    if (c->mask_prob.cap > (256u << 20)) { c->mask_prob.release(); c->mask_scale.release(); }      // long reads: give the scratch back
}

extern "C" {

int mbl_create(const mbl_config* cfg, mbl_ctx** out) {
    if (!cfg || !out) return MBL_E_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return MBL_E_NO_DEVICE;
    }
    mbl_ctx* c = new mbl_ctx();
    c->cfg = *cfg;
    try {
        if (cfg->reduced_aa) { delete c; return MBL_E_UNSUPPORTED; }
        if (cfg->kmer_format != 1 && cfg->kmer_format != 2) { delete c; return MBL_E_UNSUPPORTED; }
        // syncmer databases (SyncmerScanner is a MetamerScanner, KmerExtractor.cpp:18-20): format-2 k-mers, 2 <= s <= 7
        if (cfg->syncmer && (cfg->kmer_format != 2 || cfg->smer_len < 2 || cfg->smer_len > 7)) { delete c; return MBL_E_UNSUPPORTED; }
        MBL_CUDA(cudaSetDevice(cfg->device));
        cudaDeviceProp prop;
        MBL_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        c->sm_count = prop.multiProcessorCount;
        MBL_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        for (auto& e : c->ev) MBL_CUDA(cudaEventCreate(&e));
        HostTables t;
        c->d_base_code = upload(c, t.base_code, 256);
        c->d_codon = upload(c, t.codon, 512);
        c->d_ham_pair = upload(c, t.ham_pair, 4096);
        c->d_ham_single = upload(c, t.ham_sum, 64);
        if (const char* e = getenv("MBL_PIPELINE")) { int v = atoi(e); c->pipeline = v < 0 ? 0 : (v > 2 ? 2 : v); }
        if (const char* e = getenv("MBL_PIPELINE_PARTS")) { int v = atoi(e); if (v >= 2 && v <= 16) c->pipeline_parts = v; }
        // L2 fetch granularity on a DRAM miss (32, 64 or 128 bytes; the default fetches 128): the presence-filter probes of K1 and
        // the qinfo gathers of K3 are random 32-byte sector reads, and ncu shows 126 bytes of DRAM traffic per probe with the default
        if (const char* e = getenv("MBL_L2_FETCH")) {
            int v = atoi(e);
            if (v == 32 || v == 64 || v == 128) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v); cudaGetLastError(); }
        }
        if (const char* e = getenv("MBL_FILTER_MINIMIZER")) c->filter_minimizer = atoi(e) != 0;
        if (const char* e = getenv("MBL_FILTER_BITS")) { int v = atoi(e); if (v >= 0 && v <= 64) c->filter_bits = v; }
        if (const char* e = getenv("MBL_MERGE_THREADS")) { int v = atoi(e); if (v == 256 || v == 512) c->merge_threads = v; }
        if (const char* e = getenv("MBL_PIPELINE_MIN_READS")) { long v = atol(e); if (v > 0) c->pipeline_min_reads = (uint32_t)v; }
        if (const char* e = getenv("MBL_TEST_MATCH_CAP")) c->test_match_cap = strtoull(e, nullptr, 10);
        if (const char* e = getenv("MBL_TEST_ITEMS_CAP")) c->test_items_cap = strtoull(e, nullptr, 10);
        if (const char* e = getenv("MBL_TEST_PACK_SLOTS")) c->test_pack_slots = strtoull(e, nullptr, 10);
        if (const char* e = getenv("MBL_SORT_BIT")) { int v = atoi(e); if (v == 24 || v == 32 || v == 40) c->force_sort_bit = v; }
        if (const char* e = getenv("MBL_TILE_CELLS")) { int v = atoi(e); if (v >= 1 && v <= 8) c->tile_cells = (uint32_t)v; }
        MBL_CUDA(cudaStreamSynchronize(c->st));
    } catch (const CudaError& e) {
        fail_cuda(c, e);
        mbl_destroy(c);                 // stream, events and the tables uploaded so far
        return MBL_E_CUDA;
    } catch (...) {
        mbl_destroy(c);
        return MBL_E_HOST;
    }
    *out = c;
    return MBL_OK;
}

namespace {
void release_lane(mbl_ctx* c) {
    if (c->is_shadow) {             // borrowed from the owning context
        for (Buf* b : {&c->bases1, &c->bases2, &c->off1, &c->off2, &c->results}) { b->p = nullptr; b->cap = 0; }
    }
    for (Buf* b : {&c->bases1, &c->bases2, &c->off1, &c->off2, &c->cov1, &c->cov2, &c->w1, &c->w2, &c->slots, &c->slot_off, &c->quot_cnt,
                   &c->quot_off, &c->seg_b, &c->seg_e, &c->res_sub, &c->tax_len, &c->tax_off, &c->val_a, &c->val_b, &c->qi_a, &c->qi_b,
                   &c->cub_tmp, &c->arena, &c->chunk_bounds, &c->order_keys, &c->g_np, &c->flag_fg, &c->flag_sp, &c->fg_list, &c->sp_list,
                   &c->fg_ord, &c->flat_tmp, &c->sp_fg, &c->read_sp, &c->m_raw, &c->m_sorted, &c->key_a, &c->key_b, &c->idx_a, &c->idx_b, &c->l_score, &c->l_start,
                   &c->l_ham, &c->l_depth, &c->l_smatch, &c->l_conn, &c->p_start, &c->p_end, &c->p_score, &c->p_ham, &c->p_depth,
                   &c->p_smatch, &c->p_ematch, &c->c_start, &c->c_end, &c->s_score, &c->q_tax, &c->q_ham, &c->q_has, &c->pairs_raw,
                   &c->q_lo, &c->item_cnt, &c->item_off, &c->items, &c->counters, &c->results, &c->pairs, &c->pairs_final,
                   &c->sh_key_a, &c->sh_key_b, &c->sh_idx_a, &c->sh_idx_b, &c->sh_begin, &c->send_value, &c->send_qinfo, &c->send_match,
                   &c->sh_tmp})
        b->release();
    for (Buf* b : {&c->stage_bases1, &c->stage_bases2, &c->stage_off1, &c->stage_off2}) b->release();
    for (auto& e : c->copy_ev) if (e) cudaEventDestroy(e);
    if (c->copy_st) cudaStreamDestroy(c->copy_st);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->st) cudaStreamDestroy(c->st);
}

// the second lane: shares everything that is read-only during a batch
mbl_ctx* ensure_shadow(mbl_ctx* c) {
    if (!c->shadow) {
        mbl_ctx* s = new mbl_ctx();
        s->is_shadow = true;
        s->cfg = c->cfg; s->sm_count = c->sm_count;
        MBL_CUDA(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
        for (auto& e : s->ev) MBL_CUDA(cudaEventCreate(&e));
        c->shadow = s;
    }
    mbl_ctx* s = c->shadow;
    s->d_base_code = c->d_base_code; s->d_codon = c->d_codon; s->d_ham_pair = c->d_ham_pair; s->d_ham_single = c->d_ham_single;
    s->tile_cells = c->tile_cells; s->merge_threads = c->merge_threads; s->filter_bits = c->filter_bits; s->filter_minimizer = c->filter_minimizer;
    s->d_diff = c->d_diff; s->d_info = c->d_info; s->n_u16 = c->n_u16; s->n_kmers = c->n_kmers;
    s->dir = c->dir; s->tax = c->tax; s->db_loaded = c->db_loaded; s->filter_complete = c->filter_complete;
    s->bases1 = c->bases1; s->bases2 = c->bases2; s->off1 = c->off1; s->off2 = c->off2; s->results = c->results;   // borrowed
    s->n_reads = c->n_reads; s->paired = c->paired;
    s->match_ratio = c->match_ratio; s->pass_ratio = c->pass_ratio;
    s->n_pairs = 0;
    s->stats = mbl_stats{};
    return s;
}
}  // namespace

void mbl_destroy(mbl_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->st);
    if (c->shadow) {
        cudaStreamSynchronize(c->shadow->st);
        release_lane(c->shadow);
        delete c->shadow;
        c->shadow = nullptr;
    }
    free_db(c);
    for (uint32_t p = 0; p < kMaxShards; ++p) {
        if (c->peer_opened[p][0] && c->peer_kmers[p]) cudaIpcCloseMemHandle(c->peer_kmers[p]);
        if (c->peer_opened[p][1] && c->peer_matches[p]) cudaIpcCloseMemHandle(c->peer_matches[p]);
    }
    cudaFree(c->recv_kmers); cudaFree(c->recv_matches);
    cudaFree(c->d_base_code); cudaFree(c->d_codon); cudaFree(c->d_ham_pair); cudaFree(c->d_ham_single);
    release_lane(c);
    delete c;
}

const char* mbl_last_error(const mbl_ctx* c) { return c ? c->err.c_str() : "no context"; }

namespace {
int load_db_range(mbl_ctx* c, const mbl_db* db, const mbl_taxonomy* tx, const mbl_shard& sh, bool is_shard) {
    if (!c || !db || !tx || !db->diff_idx || !db->info) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (sh.diff_begin > sh.diff_end || sh.diff_end > db->n_u16 || sh.info_begin > sh.info_end || sh.info_end > db->n_kmers)
        return fail(c, MBL_E_BAD_ARG, "shard range outside the index");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        free_db(c);
        c->shard = sh; c->is_shard = is_shard;
        c->n_u16 = sh.diff_end - sh.diff_begin; c->n_kmers = sh.info_end - sh.info_begin;
        c->d_diff = upload(c, db->diff_idx + sh.diff_begin, c->n_u16, 64);
        c->d_info = upload(c, db->info + sh.info_begin, c->n_kmers, 64);
        auto up = [&](auto* h, size_t n) { auto* d = upload(c, h, n); c->tax_allocs.push_back((void*)d); return d; };
        const size_t N = tx->max_nodes, T = (size_t)tx->max_taxid + 1;
        c->tax.D = up(tx->D, T); c->tax.E = up(tx->E, 2 * N); c->tax.L = up(tx->L, 2 * N); c->tax.H = up(tx->H, N);
        c->tax.M = up(tx->M, 2 * N * (size_t)tx->M_k);
        c->tax.node_taxid = up(tx->node_taxid, N); c->tax.node_parent = up(tx->node_parent, N);
        c->tax.node_prune = up(tx->node_prune, N); c->tax.node_rank = up(tx->node_rank, N);
        c->tax.taxid2species = up(tx->taxid2species, T);
        c->tax.max_taxid = tx->max_taxid; c->tax.M_k = tx->M_k; c->tax.eukaryota = tx->eukaryota; c->tax.max_nodes = (uint32_t)N;
        MBL_CUDA(cudaStreamSynchronize(c->st));
        // the presence filter must cover every k-mer a query could match: a shard builds its part, sized for the whole index, and
        // the filter is only used once the ranks have OR-ed their parts together (mbl_shard_filter / mbl_shard_filter_or)
        build_tile_directory(c->d_diff, c->n_u16, c->n_kmers, c->sm_count, c->tile_cells, c->st, c->dir, sh.base_value, sh.holds_db_tail != 0,
                             c->filter_bits, db->n_kmers, c->cfg.kmer_format == 2 && c->filter_minimizer);
        c->filter_complete = !is_shard;
        if (c->force_sort_bit) c->dir.sort_begin_bit = c->force_sort_bit;
        // the k-mer count implied by the end flags must agree with the info file
        if (c->dir.n_kmers_decoded != c->n_kmers) {
            char msg[256];
            snprintf(msg, sizeof msg, "diffIdx holds %llu k-mers but info has %llu entries", (unsigned long long)c->dir.n_kmers_decoded,
                     (unsigned long long)c->n_kmers);
            free_db(c);
            return fail(c, MBL_E_BAD_DB, msg);
        }
        c->db_bytes = 2 * c->n_u16 + 4 * c->n_kmers + sizeof(Tile) * c->dir.n_tiles + 16 * c->dir.n_cells + 8 * c->dir.n_jumbo_kmers +
                      128 * (size_t)c->dir.filter_lines;
        c->db_loaded = true;
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

// one k-mer's delta: 15 payload bits per fragment, most significant group first (KmerMatcher.h:282-297)
inline uint64_t decode_delta(const uint16_t* d, uint64_t b, uint64_t e) {
    uint64_t v = 0;
    for (uint64_t i = b; i < e; ++i) v = (v << 15) | (uint64_t)(d[i] & 0x7FFFu);
    return v;
}
struct ShardCut { uint64_t first_value, base_value, diff_begin, info_begin; };
}  // namespace

int mbl_load_db(mbl_ctx* c, const mbl_db* db, const mbl_taxonomy* tx) {
    if (!c || !db) return fail(c, MBL_E_BAD_ARG, "null argument");
    mbl_shard whole{};
    whole.first_value = 0; whole.base_value = 0; whole.diff_begin = 0; whole.diff_end = db->n_u16; whole.info_begin = 0;
    whole.info_end = db->n_kmers; whole.holds_db_tail = 1;
    return load_db_range(c, db, tx, whole, false);
}

int mbl_load_db_shard(mbl_ctx* c, const mbl_db* db, const mbl_taxonomy* tx, const mbl_shard* shard) {
    if (!c || !shard) return fail(c, MBL_E_BAD_ARG, "null argument");
    return load_db_range(c, db, tx, *shard, true);
}

int mbl_plan_shards(const mbl_db* db, uint32_t n_shards, mbl_shard* out) {
    if (!db || !out || n_shards < 1 || n_shards > MBL_MAX_SHARDS || (db->n_u16 && !db->diff_idx)) return MBL_E_BAD_ARG;
    const uint16_t* d = db->diff_idx;
    const uint64_t NU = db->n_u16, NK = db->n_kmers;
    const long double total = 2.0L * NU + 4.0L * NK;
    auto bytes_at = [](const ShardCut& c) { return 2.0L * c.diff_begin + 4.0L * c.info_begin; };
    std::vector<ShardCut> cuts;                     // chosen boundaries, ascending; shard s+1 starts at cuts[s]
    if (n_shards > 1 && NU && NK) {
        // candidates from the split checkpoints: entry = {ADkmer = first k-mer of a new amino-acid group, u16 index just after it,
        // its info index + 1} (IndexCreator.cpp:849-857)
        std::vector<ShardCut> cand;
        if (db->split) {
            uint64_t last_info = 0;
            for (size_t k = 1; k < db->n_split; ++k) {
                const uint64_t ad = db->split[3 * k], doff = db->split[3 * k + 1], ioff = db->split[3 * k + 2];
                if (ad == 0 || ad == ~0ull || doff < 1 || doff > NU || ioff < 2 || ioff > NK || ioff - 1 <= last_info) continue;
                if (!(d[doff - 1] & 0x8000u)) continue;
                uint64_t b = doff - 1;
                while (b > 0 && !(d[b - 1] & 0x8000u) && doff - b < 5) --b;
                if (b > 0 && !(d[b - 1] & 0x8000u)) continue;               // more than 5 fragments: not a k-mer boundary
                const uint64_t delta = decode_delta(d, b, doff);
                if (delta > ad || ((ad - delta) & kAaMask) == (ad & kAaMask)) continue;   // must start an amino-acid group
                cand.push_back(ShardCut{ad, ad - delta, b, ioff - 1});
                last_info = ioff - 1;
            }
        }
        auto choose = [&](const std::vector<ShardCut>& cs) {
            std::vector<ShardCut> pick;
            size_t from = 0;
            for (uint32_t s = 1; s < n_shards && from < cs.size(); ++s) {
                const long double target = total * s / n_shards;
                size_t best = from;
                for (size_t i = from; i < cs.size(); ++i) {
                    if (fabsl(bytes_at(cs[i]) - target) < fabsl(bytes_at(cs[best]) - target)) best = i;
                    if (bytes_at(cs[i]) > target) break;
                }
                pick.push_back(cs[best]);
                from = best + 1;
            }
            return pick;
        };
        cuts = choose(cand);
        auto largest = [&](const std::vector<ShardCut>& cs) {
            long double prev = 0, worst = 0;
            for (const ShardCut& c : cs) { worst = std::max(worst, bytes_at(c) - prev); prev = bytes_at(c); }
            return std::max(worst, total - prev);
        };
        if (cuts.size() + 1 < n_shards || largest(cuts) > 1.25L * total / n_shards) {
            const std::vector<ShardCut> from_split = cuts;
            // too few usable checkpoints (small or hand-made DBs): one sequential pass over the stream, taking the first
            // amino-acid-group start at or after every byte target
            cuts.clear();
            uint64_t v = 0, k = 0, p = 0;
            uint32_t s = 1;
            while (p < NU && s < n_shards) {
                uint64_t e = p;
                while (e < NU && !(d[e] & 0x8000u)) ++e;
                if (e >= NU) break;                                          // truncated tail: ignored, the loader reports it
                const uint64_t nv = v + decode_delta(d, p, e + 1);
                if (k > 0 && (nv & kAaMask) != (v & kAaMask) && 2.0L * p + 4.0L * k >= total * s / n_shards) {
                    cuts.push_back(ShardCut{nv, v, p, k});
                    ++s;
                }
                v = nv; ++k; p = e + 1;
            }
            if (from_split.size() + 1 == n_shards && (cuts.size() + 1 < n_shards || largest(from_split) <= largest(cuts))) cuts = from_split;
        }
    }
    for (uint32_t s = 0; s < n_shards; ++s) {
        mbl_shard& o = out[s];
        o = mbl_shard{};
        const bool has_begin = s == 0 || s - 1 < cuts.size();
        const bool has_end = s < cuts.size();
        if (!has_begin) {                       // empty trailing shard
            o.first_value = ~0ull; o.base_value = 0; o.diff_begin = o.diff_end = NU; o.info_begin = o.info_end = NK;
            continue;
        }
        if (s > 0) { const ShardCut& b = cuts[s - 1]; o.first_value = b.first_value; o.base_value = b.base_value; o.diff_begin = b.diff_begin; o.info_begin = b.info_begin; }
        o.diff_end = has_end ? cuts[s].diff_begin : NU;
        o.info_end = has_end ? cuts[s].info_begin : NK;
        o.holds_db_tail = (!has_end && o.info_end > o.info_begin) ? 1 : 0;
    }
    return MBL_OK;
}

int mbl_upload_batch(mbl_ctx* c, const mbl_batch* b) {
    if (!c || !b || !b->bases || !b->offsets) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (b->n_reads >= (1u << 29)) return fail(c, MBL_E_BAD_ARG, "at most 2^29-1 reads per batch (29-bit sequenceID, Kmer.h:13)");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        StageTimer t(c, MBL_STAGE_H2D);
        const uint32_t n = b->n_reads;
        c->n_reads = n;
        c->paired = b->bases2 != nullptr && b->offsets2 != nullptr;
        const uint64_t nb1 = b->offsets[n];
        uint8_t* d1 = c->bases1.get<uint8_t>(nb1 + 64);
        MBL_CUDA(cudaMemcpyAsync(d1, b->bases, nb1, cudaMemcpyHostToDevice, c->st));
        MBL_CUDA(cudaMemcpyAsync(c->off1.get<uint64_t>(n + 1), b->offsets, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, c->st));
        if (c->paired) {
            const uint64_t nb2 = b->offsets2[n];
            uint8_t* d2 = c->bases2.get<uint8_t>(nb2 + 64);
            MBL_CUDA(cudaMemcpyAsync(d2, b->bases2, nb2, cudaMemcpyHostToDevice, c->st));
            MBL_CUDA(cudaMemcpyAsync(c->off2.get<uint64_t>(n + 1), b->offsets2, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, c->st));
        }
        summarize_reads(b, c->summary);                        // host work overlaps the copies
        c->subs.clear();
        c->probe_first = c->match_ratio == 0.0 && !c->no_probe;
        const uint64_t max_slots = slots_budget(c);
        plan_sub_batches(c, c->summary, 0, max_slots, c->probe_first, c->subs);
        if (c->budget_low && c->subs.size() > 64) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            char msg[256];
            snprintf(msg, sizeof msg, "only %.1f of %.1f GB of device memory are free next to the index: the batch would be cut into %zu "
                     "sub-batches of ~%llu k-mer slots (each streams the whole index); use a GPU with more memory or the index-sharded mode",
                     free_b / 1e9, total_b / 1e9, c->subs.size(), (unsigned long long)max_slots);
            t.stop();
            return fail(c, MBL_E_CAPACITY, msg);
        }
        c->probe_first = c->probe_first && c->subs.size() > 1;
        c->plan_has_probe = c->probe_first;
        t.stop();
        mask_resident(c);
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_classify_resident(mbl_ctx* c) {
    if (!c) return MBL_E_BAD_ARG;
    if (!c->db_loaded) return fail(c, MBL_E_BAD_ARG, "mbl_load_db has not been called");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        float h2d = c->stats.ms[MBL_STAGE_H2D];
        c->stats = mbl_stats{};
        c->stats.ms[MBL_STAGE_H2D] = h2d;
        c->n_pairs = 0;
        c->results.get<mbl_read_result>(c->n_reads + 1);
        if (c->plan_has_probe && !c->probe_first && c->match_ratio > 0.0) {
            // the resident batch is classified again (benchmarks): its plan still starts with the probe sub-batch of the first
            // run against this index; the ratios are known now, plan it whole
            c->subs.clear();
            plan_sub_batches(c, c->summary, 0, slots_budget(c), false, c->subs);
            c->plan_has_probe = false;
        }
        c->stats.sub_batches = (uint32_t)c->subs.size();
        struct SubDone { int lane; uint64_t lane_off, count; };
        std::vector<SubDone> done(c->subs.size(), SubDone{0, 0, 0});
        const bool two = c->pipeline && c->subs.size() >= 2;
        const bool staggered = two && c->pipeline == 2;
        mbl_ctx* s = two ? ensure_shadow(c) : nullptr;
        int rc1 = MBL_OK;
        std::thread lane1;
        if (staggered) {
            // Sub-batch k runs on lane k & 1.  One thread walks the FRONT halves (K1-K3: DRAM- and issue-bound), this thread the
            // BACK halves (K4-K5: radix passes and latency-bound scoring), so back(k) overlaps front(k+1) on the other lane's
            // stream; front(k+2) waits for back(k) because it reuses that lane's workspace.
            mbl_ctx* lanes[2] = {c, s};
            const size_t n = c->subs.size();
            std::vector<uint64_t> reserved(n, 0);
            std::vector<char> fdone(n, 0), bdone(n, 0);
            std::mutex mu;
            std::condition_variable cv;
            bool stop = false;
            lane1 = std::thread([&] {
                cudaSetDevice(c->cfg.device);
                for (size_t k = 0; k < n; ++k) {
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || k < 2 || bdone[k - 2]; });
                        if (stop) return;
                    }
                    int rc = MBL_OK;
                    try { rc = sub_front(lanes[k & 1], c->subs[k], &reserved[k]); } catch (const CudaError& e) { rc = fail_cuda(lanes[k & 1], e); }
                    catch (...) { rc = fail(lanes[k & 1], MBL_E_HOST, "host error in the pipeline lane"); }
                    std::lock_guard<std::mutex> lk(mu);
                    if (rc != MBL_OK) { rc1 = rc; c->err = lanes[k & 1]->err; stop = true; cv.notify_all(); return; }
                    fdone[k] = 1;
                    cv.notify_all();
                }
            });
            int rcb = MBL_OK;
            for (size_t k = 0; k < n; ++k) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || fdone[k]; });
                    if (stop) break;
                }
                mbl_ctx* L = lanes[k & 1];
                const uint64_t off = L->n_pairs;
                try { rcb = sub_back(L, c->subs[k], reserved[k]); } catch (const CudaError& e) { rcb = fail_cuda(L, e); }
                catch (...) { rcb = fail(L, MBL_E_HOST, "host error in the pipeline lane"); }
                std::lock_guard<std::mutex> lk(mu);
                if (rcb != MBL_OK) { if (L != c) c->err = L->err; stop = true; cv.notify_all(); break; }
                done[k] = SubDone{(int)(k & 1), off, L->n_pairs - off};
                bdone[k] = 1;
                cv.notify_all();
            }
            lane1.join();
            if (rcb != MBL_OK) return rcb;
            if (rc1 != MBL_OK) return rc1;
        } else if (two) {
            lane1 = std::thread([&] {
                cudaSetDevice(c->cfg.device);
                try {
                    for (size_t k = 1; k < c->subs.size(); k += 2) {
                        const uint64_t off = s->n_pairs;
                        rc1 = run_sub_batch(s, c->subs[k]);
                        if (rc1 != MBL_OK) break;
                        done[k] = SubDone{1, off, s->n_pairs - off};
                    }
                } catch (const CudaError& e) {
                    rc1 = fail_cuda(s, e);
                } catch (...) {
                    rc1 = fail(s, MBL_E_HOST, "host error in the pipeline lane");
                }
            });
        }
        int rc0 = MBL_OK;
        if (!two) {
            // one lane: sub-batches in order.  After a probe sub-batch the rest of the batch is planned again with the ratios
            // it measured; a sub-batch that runs out of device memory all the same is halved (whole blocks) and redone
            for (size_t k = 0; k < c->subs.size() && rc0 == MBL_OK; ++k) {
                const uint64_t pairs_before = c->n_pairs;
                const mbl_stats stats_before = c->stats;
                try {
                    rc0 = run_sub_batch(c, c->subs[k]);
                } catch (const CudaError& e) {
                    const SubBatch sb = c->subs[k];
                    if (e.code != cudaErrorMemoryAllocation || sb.b1 - sb.b0 < 2) { rc0 = fail_cuda(c, e); break; }
                    cudaGetLastError();
                    MBL_CUDA(cudaStreamSynchronize(c->st));
                    c->n_pairs = pairs_before;
                    c->stats = stats_before;
                    c->stats.overflow_retries += 1;
                    for (Buf* b : {&c->arena, &c->m_raw, &c->cub_tmp}) b->release();     // regrown to what the halves need
                    const uint32_t mid = sb.b0 + (sb.b1 - sb.b0) / 2;
                    c->subs[k] = sub_of_blocks(c->summary, sb.b0, mid);
                    c->subs.insert(c->subs.begin() + (ptrdiff_t)k + 1, sub_of_blocks(c->summary, mid, sb.b1));
                    --k;
                    continue;
                }
                if (rc0 == MBL_OK && k == 0 && c->probe_first) {
                    const uint32_t next_block = c->subs[0].b1;
                    c->subs.resize(1);
                    plan_sub_batches(c, c->summary, next_block, slots_budget(c), false, c->subs);
                    c->probe_first = false;
                }
            }
            c->stats.sub_batches = (uint32_t)c->subs.size();
        } else if (!staggered) try {
            for (size_t k = 0; k < c->subs.size(); k += 2) {
                const uint64_t off = c->n_pairs;
                rc0 = run_sub_batch(c, c->subs[k]);
                if (rc0 != MBL_OK) break;
                done[k] = SubDone{0, off, c->n_pairs - off};
            }
        } catch (const CudaError& e) {
            rc0 = fail_cuda(c, e);
        }
        if (two && !staggered) lane1.join();
        if (rc0 != MBL_OK) return rc0;
        if (rc1 != MBL_OK) { c->err = s->err; return rc1; }
        c->pairs_out = c->pairs.p;
        c->n_pairs_total = c->n_pairs;
        if (two) {
            // per-stage times of the two lanes add up (they overlap in wall-clock time); counts add up
            for (int i = 0; i < 7; ++i) if (i != MBL_STAGE_H2D) c->stats.ms[i] += s->stats.ms[i];
            c->stats.merge_kernel_ms += s->stats.merge_kernel_ms; c->stats.n_query_kmers += s->stats.n_query_kmers;
            c->stats.n_merge_queries += s->stats.n_merge_queries;
            c->stats.n_matches += s->stats.n_matches; c->stats.merge_bytes += s->stats.merge_bytes;
            c->stats.merge_launches += s->stats.merge_launches; c->stats.kernel_launches += s->stats.kernel_launches;
            c->stats.overflow_retries += s->stats.overflow_retries;
            c->match_ratio = std::max(c->match_ratio, s->match_ratio);
            c->pass_ratio = std::max(c->pass_ratio, s->pass_ratio);
            // the batch's pair array: the sub-batches' pairs in sub-batch order
            uint64_t total = 0;
            for (const SubDone& d : done) total += d.count;
            int32_t* fin = c->pairs_final.get<int32_t>(2 * total + 2);
            uint64_t base = 0;
            for (size_t k = 0; k < done.size(); ++k) {
                const SubDone& d = done[k];
                const mbl_ctx* src = d.lane ? s : c;
                if (d.count) MBL_CUDA(cudaMemcpyAsync(fin + 2 * base, (const int32_t*)src->pairs.p + 2 * d.lane_off, 8 * d.count, cudaMemcpyDeviceToDevice, c->st));
                if (base != d.lane_off)
                    launch_shift_taxcnt((mbl_read_result*)c->results.p + c->subs[k].r0, c->subs[k].r1 - c->subs[k].r0,
                                        (uint32_t)(base - d.lane_off), c->st);
                base += d.count;
            }
            c->pairs_out = fin;
            c->n_pairs_total = total;
            c->stats.kernel_launches += (uint32_t)done.size();
        }
        MBL_CUDA(cudaStreamSynchronize(c->st));
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_download_results(mbl_ctx* c, mbl_read_result* out, int32_t* taxcnt_pairs, size_t cap_pairs, size_t* used_pairs) {
    if (!c || !out) return fail(c, MBL_E_BAD_ARG, "null argument");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        if (used_pairs) *used_pairs = c->n_pairs_total;
        if (c->n_pairs_total > cap_pairs) return fail(c, MBL_E_CAPACITY, "taxcnt_pairs too small");
        StageTimer t(c, MBL_STAGE_D2H);
        if (c->n_reads) MBL_CUDA(cudaMemcpyAsync(out, c->results.p, sizeof(mbl_read_result) * (size_t)c->n_reads, cudaMemcpyDeviceToHost, c->st));
        if (c->n_pairs_total && taxcnt_pairs) MBL_CUDA(cudaMemcpyAsync(taxcnt_pairs, c->pairs_out, 8 * c->n_pairs_total, cudaMemcpyDeviceToHost, c->st));
        t.stop();
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_classify_batch(mbl_ctx* c, const mbl_batch* b, mbl_read_result* out, int32_t* taxcnt_pairs, size_t cap_pairs, size_t* used_pairs) {
    if (!c) return MBL_E_BAD_ARG;
    c->stats.ms[MBL_STAGE_H2D] = 0;
    int rc = mbl_upload_batch(c, b);
    if (rc != MBL_OK) return rc;
    rc = mbl_classify_resident(c);
    if (rc != MBL_OK) return rc;
    return mbl_download_results(c, out, taxcnt_pairs, cap_pairs, used_pairs);
}

// ---- streaming: the next batch is uploaded while the current one is classified ------------------------------------------
int mbl_prefetch_batch(mbl_ctx* c, const mbl_batch* b) {
    if (!c || !b || !b->bases || !b->offsets) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (b->n_reads >= (1u << 29)) return fail(c, MBL_E_BAD_ARG, "at most 2^29-1 reads per batch (29-bit sequenceID, Kmer.h:13)");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        if (!c->copy_st) {
            MBL_CUDA(cudaStreamCreateWithFlags(&c->copy_st, cudaStreamNonBlocking));
            for (auto& e : c->copy_ev) MBL_CUDA(cudaEventCreate(&e));
        }
        const uint32_t n = b->n_reads;
        c->staged_reads = n;
        c->staged_paired = b->bases2 != nullptr && b->offsets2 != nullptr;
        MBL_CUDA(cudaEventRecord(c->copy_ev[0], c->copy_st));
        const uint64_t nb1 = b->offsets[n];
        MBL_CUDA(cudaMemcpyAsync(c->stage_bases1.get<uint8_t>(nb1 + 64), b->bases, nb1, cudaMemcpyHostToDevice, c->copy_st));
        MBL_CUDA(cudaMemcpyAsync(c->stage_off1.get<uint64_t>(n + 1), b->offsets, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, c->copy_st));
        if (c->staged_paired) {
            const uint64_t nb2 = b->offsets2[n];
            MBL_CUDA(cudaMemcpyAsync(c->stage_bases2.get<uint8_t>(nb2 + 64), b->bases2, nb2, cudaMemcpyHostToDevice, c->copy_st));
            MBL_CUDA(cudaMemcpyAsync(c->stage_off2.get<uint64_t>(n + 1), b->offsets2, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, c->copy_st));
        }
        MBL_CUDA(cudaEventRecord(c->copy_ev[1], c->copy_st));
        summarize_reads(b, c->staged_summary);
        c->staged_subs.clear();
        c->staged_probe_first = c->match_ratio == 0.0 && !c->no_probe;
        plan_sub_batches(c, c->staged_summary, 0, slots_budget(c), c->staged_probe_first, c->staged_subs);
        c->staged_probe_first = c->staged_probe_first && c->staged_subs.size() > 1;
        c->staged = true;
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_classify_prefetched(mbl_ctx* c, const mbl_batch* next, mbl_read_result* out, int32_t* taxcnt_pairs, size_t cap_pairs,
                            size_t* used_pairs) {
    if (!c) return MBL_E_BAD_ARG;
    if (!c->staged) return fail(c, MBL_E_BAD_ARG, "mbl_prefetch_batch has not been called");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        // the staged batch becomes the resident one; what was resident is free (the previous classify has returned)
        MBL_CUDA(cudaEventSynchronize(c->copy_ev[1]));
        float ms = 0;
        cudaEventElapsedTime(&ms, c->copy_ev[0], c->copy_ev[1]);
        std::swap(c->bases1, c->stage_bases1); std::swap(c->bases2, c->stage_bases2);
        std::swap(c->off1, c->stage_off1); std::swap(c->off2, c->stage_off2);
        c->subs.swap(c->staged_subs);
        std::swap(c->summary, c->staged_summary);
        c->probe_first = c->staged_probe_first;
        c->plan_has_probe = c->probe_first;
        c->n_reads = c->staged_reads; c->paired = c->staged_paired;
        c->staged = false;
        c->stats.ms[MBL_STAGE_H2D] = ms;
        mask_resident(c);
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    if (next) {
        int rc = mbl_prefetch_batch(c, next);
        if (rc != MBL_OK) return rc;
    }
    int rc = mbl_classify_resident(c);
    if (rc != MBL_OK) return rc;
    return mbl_download_results(c, out, taxcnt_pairs, cap_pairs, used_pairs);
}

// ---- index-sharded mode: the phases one rank runs around the two exchanges (include/metabuli_b200.h) ---------------------
namespace {
// same-process ranks hand each other raw device pointers: a kernel of this context may only touch memory of another device
// after peer access is enabled from this context's device (cross-process peers come through cudaIpcOpenMemHandle, which does it)
void enable_peer_for(mbl_ctx* c, const void* p) {
    if (!p) return;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (at.type != cudaMemoryTypeDevice || at.device == c->cfg.device) return;
    int can = 0;
    cudaDeviceCanAccessPeer(&can, c->cfg.device, at.device);
    if (!can) throw CudaError{cudaErrorPeerAccessUnsupported, __FILE__, __LINE__};
    cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) throw CudaError{e, __FILE__, __LINE__};
    cudaGetLastError();
}
}  // namespace

int mbl_shard_filter(mbl_ctx* c, void** d_words, uint64_t* n_bytes) {
    if (!c || !d_words || !n_bytes) return fail(c, MBL_E_BAD_ARG, "null argument");
    *d_words = c->dir.filter;
    *n_bytes = 128ull * c->dir.filter_lines;
    return MBL_OK;
}

int mbl_shard_filter_or(mbl_ctx* c, const void* d_other, uint64_t n_bytes, int complete) {
    if (!c) return MBL_E_BAD_ARG;
    if (!c->dir.filter) { c->filter_complete = false; return MBL_OK; }          // MBL_FILTER_BITS=0
    if (d_other && n_bytes != 128ull * c->dir.filter_lines) return fail(c, MBL_E_BAD_ARG, "filter sizes differ between the ranks");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        enable_peer_for(c, d_other);
        if (d_other) launch_filter_or(c->dir.filter, (const uint32_t*)d_other, n_bytes / 4, c->st);
        MBL_CUDA(cudaStreamSynchronize(c->st));
        MBL_CUDA(cudaGetLastError());
        if (complete) c->filter_complete = true;
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_extract(mbl_ctx* c, const mbl_batch* b, uint64_t seq_base, uint32_t n_shards, const uint64_t* shard_first_value,
                      uint64_t* send_counts) {
    if (!c || !b || !shard_first_value || !send_counts) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (n_shards < 1 || n_shards > kMaxShards) return fail(c, MBL_E_BAD_ARG, "1..64 shards");
    if (seq_base + b->n_reads >= (1ull << 29)) return fail(c, MBL_E_BAD_ARG, "global read index exceeds the 29-bit sequenceID (Kmer.h:13)");
    c->stats.ms[MBL_STAGE_H2D] = 0;
    const int pipeline = c->pipeline;
    c->pipeline = 0;                                   // one sub-batch: the exchange works on whole batches
    c->no_probe = true;
    int rc = mbl_upload_batch(c, b);
    c->pipeline = pipeline;
    c->no_probe = false;
    if (rc != MBL_OK) return rc;
    if (c->subs.size() > 1) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        uint64_t slots = 0;
        for (const SubBatch& x : c->subs) slots += x.slots;
        char msg[320];
        snprintf(msg, sizeof msg, "batch too large for one sharded pass (%llu slots, budget %llu slots, %.1f of %.1f GB free, match ratio %.3f, "
                 "pass ratio %.3f): split it on the caller's side", (unsigned long long)slots, (unsigned long long)slots_budget(c), free_b / 1e9,
                 total_b / 1e9, c->match_ratio, c->pass_ratio);
        return fail(c, MBL_E_CAPACITY, msg);
    }
    try {
        cudaStream_t st = c->st;
        float h2d = c->stats.ms[MBL_STAGE_H2D];
        c->stats = mbl_stats{};
        c->stats.ms[MBL_STAGE_H2D] = h2d;
        c->stats.sub_batches = 1;
        c->n_pairs = 0; c->n_pairs_total = 0;
        c->seq_base = seq_base;
        c->results.get<mbl_read_result>(c->n_reads + 1);
        for (uint32_t s = 0; s < n_shards; ++s) send_counts[s] = 0;
        c->sh_n = n_shards; c->sh_perm = nullptr;
        for (uint32_t s = 0; s <= n_shards; ++s) c->sh_begin_h[s] = 0;
        if (c->subs.empty()) return MBL_OK;
        const SubBatch sb = c->subs[0];
        uint64_t S = sb.slots;
        if (c->dir.filter && c->filter_complete) {
            int rc2 = extract_filtered(c, sb, &S);
            if (rc2 != MBL_OK) return rc2;
        } else {
            stage_extract(c, sb, false);
        }
        c->sh_S8 = c->arena_S8;
        const uint64_t* va = (const uint64_t*)c->arena.p;
        const float ms_k1 = c->stats.ms[MBL_STAGE_EXTRACT];
        StageTimer t(c, MBL_STAGE_EXTRACT);
        ShardBounds sbnd{};
        sbnd.n = n_shards;
        for (uint32_t s = 0; s < n_shards; ++s) sbnd.bound[s] = shard_first_value[s];
        uint8_t *ka = c->sh_key_a.get<uint8_t>(S + 16), *kb = c->sh_key_b.get<uint8_t>(S + 16);
        uint32_t *ia = c->sh_idx_a.get<uint32_t>(S + 16), *ib = c->sh_idx_b.get<uint32_t>(S + 16);
        uint64_t* d_begin = c->sh_begin.get<uint64_t>(kMaxShards + 2);
        void* tmp = c->sh_tmp.get<uint8_t>(bucket_sort_temp_bytes(S));
        c->sh_perm = bucket_kmers(tmp, c->sh_tmp.cap, va, S, sbnd, ka, kb, ia, ib, d_begin, st);
        MBL_CUDA(cudaMemcpyAsync(c->sh_begin_h, d_begin, 8 * (size_t)(n_shards + 1), cudaMemcpyDeviceToHost, st));
        c->stats.kernel_launches += 3;
        t.stop();
        c->stats.ms_bucket_kmers = c->stats.ms[MBL_STAGE_EXTRACT] - ms_k1;
        for (uint32_t s = 0; s < n_shards; ++s) send_counts[s] = c->sh_begin_h[s + 1] - c->sh_begin_h[s];
        MBL_CUDA(cudaGetLastError());
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_pack_kmers(mbl_ctx* c, const uint64_t** d_send_value, const uint64_t** d_send_qinfo) {
    if (!c || !d_send_value || !d_send_qinfo) return fail(c, MBL_E_BAD_ARG, "null argument");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        *d_send_value = nullptr; *d_send_qinfo = nullptr;
        const uint64_t n_send = c->sh_begin_h[c->sh_n];
        if (!n_send || c->subs.empty()) return MBL_OK;
        const uint64_t S8 = c->sh_S8;
        const uint64_t* va = (const uint64_t*)c->arena.p;
        const float ms_before = c->stats.ms[MBL_STAGE_EXTRACT];
        StageTimer t(c, MBL_STAGE_EXTRACT);
        uint64_t *sv = c->send_value.get<uint64_t>(n_send + 16), *sq = c->send_qinfo.get<uint64_t>(n_send + 16);
        gather_kmers(c->sh_perm, n_send, va, va + 2 * S8, c->seq_base, sv, sq, c->st);
        c->stats.kernel_launches += 1;
        t.stop();
        c->stats.ms_bucket_kmers += c->stats.ms[MBL_STAGE_EXTRACT] - ms_before;
        MBL_CUDA(cudaGetLastError());
        *d_send_value = sv; *d_send_qinfo = sq;
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_match(mbl_ctx* c, const uint64_t* d_value, const uint64_t* d_qinfo, uint64_t n, uint32_t n_owners,
                    const uint64_t* owner_first_read, uint64_t* send_counts) {
    if (!c || !owner_first_read || !send_counts || (n && (!d_value || !d_qinfo))) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (!c->db_loaded) return fail(c, MBL_E_BAD_ARG, "mbl_load_db[_shard] has not been called");
    if (n_owners < 1 || n_owners > kMaxShards) return fail(c, MBL_E_BAD_ARG, "1..64 owners");
    if (n >= 4000000000ull) return fail(c, MBL_E_CAPACITY, "more than 4e9 received metamers in one pass");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        cudaStream_t st = c->st;
        for (uint32_t o = 0; o < n_owners; ++o) send_counts[o] = 0;
        c->sh_n = n_owners; c->sh_perm = nullptr;
        for (uint32_t o = 0; o <= n_owners; ++o) c->sh_begin_h[o] = 0;
        // the received metamers become the phase-1 arena of this context: keys in `value A`, their positions in `slot idx A`
        const uint64_t S8 = (n + 31) & ~31ull;
        uint64_t* ar = c->arena.get<uint64_t>(4 * S8 + 64);
        uint32_t* ia = reinterpret_cast<uint32_t*>(ar + 3 * S8);
        if (n) MBL_CUDA(cudaMemcpyAsync(ar, d_value, 8 * n, cudaMemcpyDeviceToDevice, st));
        launch_iota(ia, n, st);
        uint64_t reserved = 0, n_match = 0;
        c->arena_S8 = S8;
        int rc = stage_sort_merge(c, n, n, &c->shard_match_ratio, d_qinfo, false, &reserved, &n_match);
        c->stats.n_query_kmers = c->stats.n_merge_queries;
        if (rc != MBL_OK) return rc;
        const float ms_before = c->stats.ms[MBL_STAGE_MSORT];
        StageTimer t(c, MBL_STAGE_MSORT);
        ShardBounds ob{};
        ob.n = n_owners;
        for (uint32_t o = 0; o < n_owners; ++o) ob.bound[o] = owner_first_read[o];
        const uint64_t R = reserved;
        uint8_t *ka = c->sh_key_a.get<uint8_t>(R + 16), *kb = c->sh_key_b.get<uint8_t>(R + 16);
        uint32_t *xa = c->sh_idx_a.get<uint32_t>(R + 16), *xb = c->sh_idx_b.get<uint32_t>(R + 16);
        uint64_t* d_begin = c->sh_begin.get<uint64_t>(kMaxShards + 2);
        void* tmp = c->sh_tmp.get<uint8_t>(bucket_sort_temp_bytes(R));
        c->sh_perm = bucket_matches(tmp, c->sh_tmp.cap, (const mbl_match_rec*)c->m_raw.p, R, ob, ka, kb, xa, xb, d_begin, st);
        MBL_CUDA(cudaMemcpyAsync(c->sh_begin_h, d_begin, 8 * (size_t)(n_owners + 1), cudaMemcpyDeviceToHost, st));
        c->stats.kernel_launches += 4;
        t.stop();
        c->stats.ms_bucket_matches += c->stats.ms[MBL_STAGE_MSORT] - ms_before;
        if (c->sh_begin_h[n_owners] != n_match) return fail(c, MBL_E_CUDA, "internal: bucketed match count differs from the merge count");
        for (uint32_t o = 0; o < n_owners; ++o) send_counts[o] = c->sh_begin_h[o + 1] - c->sh_begin_h[o];
        MBL_CUDA(cudaGetLastError());
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_pack_matches(mbl_ctx* c, const mbl_match_rec** d_send_match) {
    if (!c || !d_send_match) return fail(c, MBL_E_BAD_ARG, "null argument");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        *d_send_match = nullptr;
        const uint64_t n_send = c->sh_begin_h[c->sh_n];
        if (!n_send) return MBL_OK;
        const float ms_before = c->stats.ms[MBL_STAGE_MSORT];
        StageTimer t(c, MBL_STAGE_MSORT);
        mbl_match_rec* sm = c->send_match.get<mbl_match_rec>(n_send + 16);
        gather_matches(c->sh_perm, n_send, (const mbl_match_rec*)c->m_raw.p, sm, c->st);
        c->stats.kernel_launches += 1;
        t.stop();
        c->stats.ms_bucket_matches += c->stats.ms[MBL_STAGE_MSORT] - ms_before;
        MBL_CUDA(cudaGetLastError());
        *d_send_match = sm;
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

// ---- peer-memory transport: the bucket-gather kernels store straight into the receivers' buffers over NVLink ---------------
int mbl_shard_recv_buffers(mbl_ctx* c, uint64_t kmer_rows, uint64_t match_rows, void** d_kmers, void** d_matches,
                           uint8_t* handle_kmers, uint8_t* handle_matches) {
    if (!c || !d_kmers || !d_matches) return fail(c, MBL_E_BAD_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MBL_IPC_HANDLE_BYTES, "handle size");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        // plain cudaMalloc allocations of their own (IPC handles name whole allocations)
        for (void** p : {&c->recv_kmers, &c->recv_matches}) { if (*p) cudaFree(*p); *p = nullptr; }
        {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            if (16 * kmer_rows + 24 * match_rows + 512 > free_b) {
                char msg[320];
                snprintf(msg, sizeof msg, "receive buffers for %llu metamers + %llu matches need %.1f GB, %.1f of %.1f GB free (arena %.1f GB, index %.1f GB)",
                         (unsigned long long)kmer_rows, (unsigned long long)match_rows, (16.0 * kmer_rows + 24.0 * match_rows) / 1e9, free_b / 1e9,
                         total_b / 1e9, c->arena.cap / 1e9, c->db_bytes / 1e9);
                return fail(c, MBL_E_CAPACITY, msg);
            }
        }
        MBL_CUDA(cudaMalloc(&c->recv_kmers, 16 * kmer_rows + 256));
        MBL_CUDA(cudaMalloc(&c->recv_matches, 24 * match_rows + 256));
        c->recv_kmer_rows = kmer_rows; c->recv_match_rows = match_rows;
        if (handle_kmers) MBL_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_kmers), c->recv_kmers));
        if (handle_matches) MBL_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_matches), c->recv_matches));
        *d_kmers = c->recv_kmers; *d_matches = c->recv_matches;
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_attach_peer(mbl_ctx* c, uint32_t peer, const uint8_t* handle_kmers, const uint8_t* handle_matches, void* raw_kmers,
                          void* raw_matches) {
    if (!c || peer >= kMaxShards) return fail(c, MBL_E_BAD_ARG, "bad peer");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        for (int k = 0; k < 2; ++k) {
            void*& slot = k ? c->peer_matches[peer] : c->peer_kmers[peer];
            bool& opened = k ? c->peer_opened[peer][1] : c->peer_opened[peer][0];
            if (opened && slot) cudaIpcCloseMemHandle(slot);
            opened = false; slot = nullptr;
            void* raw = k ? raw_matches : raw_kmers;
            const uint8_t* h = k ? handle_matches : handle_kmers;
            if (raw) { enable_peer_for(c, raw); slot = raw; continue; }   // same process (or this rank itself)
            if (!h) return fail(c, MBL_E_BAD_ARG, "neither a handle nor a pointer for the peer buffer");
            cudaIpcMemHandle_t hh;
            memcpy(&hh, h, sizeof hh);
            MBL_CUDA(cudaIpcOpenMemHandle(&slot, hh, cudaIpcMemLazyEnablePeerAccess));
            opened = true;
        }
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_detach_peers(mbl_ctx* c) {
    if (!c) return MBL_E_BAD_ARG;
    cudaSetDevice(c->cfg.device);
    for (uint32_t p = 0; p < kMaxShards; ++p) {
        if (c->peer_opened[p][0] && c->peer_kmers[p]) cudaIpcCloseMemHandle(c->peer_kmers[p]);
        if (c->peer_opened[p][1] && c->peer_matches[p]) cudaIpcCloseMemHandle(c->peer_matches[p]);
        c->peer_opened[p][0] = c->peer_opened[p][1] = false;
        c->peer_kmers[p] = c->peer_matches[p] = nullptr;
    }
    cudaGetLastError();
    return MBL_OK;
}

int mbl_shard_push_kmers(mbl_ctx* c, const uint64_t* dst_row_offset, const uint64_t* dst_total_rows) {
    if (!c || !dst_row_offset || !dst_total_rows) return fail(c, MBL_E_BAD_ARG, "null argument");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        const uint64_t n_send = c->sh_begin_h[c->sh_n];
        if (!n_send || c->subs.empty()) return MBL_OK;
        PushDst d{};
        d.n = c->sh_n;
        for (uint32_t s = 0; s < c->sh_n; ++s) {
            d.begin[s] = c->sh_begin_h[s];
            d.base[s] = c->peer_kmers[s]; d.row_off[s] = dst_row_offset[s]; d.total[s] = dst_total_rows[s];
            if (c->sh_begin_h[s + 1] > c->sh_begin_h[s] && !d.base[s]) return fail(c, MBL_E_BAD_ARG, "peer buffer not attached");
        }
        d.begin[c->sh_n] = n_send;
        const uint64_t S8 = c->sh_S8;
        const uint64_t* va = (const uint64_t*)c->arena.p;
        const float ms_before = c->stats.ms[MBL_STAGE_EXTRACT];
        StageTimer t(c, MBL_STAGE_EXTRACT);
        push_kmers(c->sh_perm, n_send, va, va + 2 * S8, c->seq_base, d, c->st);
        c->stats.kernel_launches += 1;
        t.stop();
        c->stats.ms_bucket_kmers += c->stats.ms[MBL_STAGE_EXTRACT] - ms_before;
        c->stats.ms_push_kmers += c->stats.ms[MBL_STAGE_EXTRACT] - ms_before;
        MBL_CUDA(cudaGetLastError());
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_push_matches(mbl_ctx* c, const uint64_t* dst_row_offset) {
    if (!c || !dst_row_offset) return fail(c, MBL_E_BAD_ARG, "null argument");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        const uint64_t n_send = c->sh_begin_h[c->sh_n];
        if (!n_send) return MBL_OK;
        PushDst d{};
        d.n = c->sh_n;
        for (uint32_t s = 0; s < c->sh_n; ++s) {
            d.begin[s] = c->sh_begin_h[s];
            d.base[s] = c->peer_matches[s]; d.row_off[s] = dst_row_offset[s]; d.total[s] = 0;
            if (c->sh_begin_h[s + 1] > c->sh_begin_h[s] && !d.base[s]) return fail(c, MBL_E_BAD_ARG, "peer buffer not attached");
        }
        d.begin[c->sh_n] = n_send;
        const float ms_before = c->stats.ms[MBL_STAGE_MSORT];
        StageTimer t(c, MBL_STAGE_MSORT);
        push_matches(c->sh_perm, n_send, (const mbl_match_rec*)c->m_raw.p, d, c->st);
        c->stats.kernel_launches += 1;
        t.stop();
        c->stats.ms_bucket_matches += c->stats.ms[MBL_STAGE_MSORT] - ms_before;
        c->stats.ms_push_matches += c->stats.ms[MBL_STAGE_MSORT] - ms_before;
        MBL_CUDA(cudaGetLastError());
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_shard_score(mbl_ctx* c, const mbl_match_rec* d_match, uint64_t n_match) {
    if (!c || (n_match && !d_match)) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (!c->db_loaded) return fail(c, MBL_E_BAD_ARG, "mbl_load_db[_shard] has not been called");
    if (n_match >= (1ull << 32)) return fail(c, MBL_E_UNSUPPORTED, "more than 2^32 matches in one pass");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        c->pairs_out = c->pairs.p; c->n_pairs_total = 0;
        if (c->subs.empty()) return MBL_OK;
        mbl_match_rec* raw = c->m_raw.get<mbl_match_rec>(n_match + 64);
        localize_matches(d_match, n_match, c->seq_base, raw, c->st);
        c->stats.kernel_launches += 1;
        int rc = stage_sort_score(c, c->subs[0], n_match);
        if (rc != MBL_OK) return rc;
        c->pairs_out = c->pairs.p;
        c->n_pairs_total = c->n_pairs;
        MBL_CUDA(cudaStreamSynchronize(c->st));
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_download_reads(mbl_ctx* c, int mate, char* out, size_t n_bytes) {
    if (!c || !out || (mate != 1 && mate != 2)) return fail(c, MBL_E_BAD_ARG, "null argument or mate not 1 / 2");
    const Buf& src = mate == 1 ? c->bases1 : c->bases2;
    if ((mate == 2 && !c->paired) || n_bytes + 64 > src.cap) return fail(c, MBL_E_BAD_ARG, "no such resident reads");
    try {
        MBL_CUDA(cudaSetDevice(c->cfg.device));
        MBL_CUDA(cudaMemcpyAsync(out, src.p, n_bytes, cudaMemcpyDeviceToHost, c->st));
        MBL_CUDA(cudaStreamSynchronize(c->st));
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return MBL_E_BAD_ARG;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MBL_E_NO_DEVICE : MBL_E_CUDA; }
    return MBL_OK;
}
int mbl_host_unregister(void* ptr) {
    if (!ptr) return MBL_E_BAD_ARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return MBL_E_CUDA; }
    return MBL_OK;
}

int mbl_get_stats(const mbl_ctx* c, mbl_stats* out) {
    if (!c || !out) return MBL_E_BAD_ARG;
    *out = c->stats;
    out->ms_mask = c->ms_mask;
    return MBL_OK;
}

int mbl_get_db_info(const mbl_ctx* c, mbl_db_info* out) {
    if (!c || !out) return MBL_E_BAD_ARG;
    out->n_tiles = c->dir.n_tiles; out->n_jumbo = c->dir.n_jumbo; out->n_kmers = c->n_kmers; out->n_u16 = c->n_u16;
    out->hbm_bytes = c->db_bytes;
    return MBL_OK;
}

// ---- stage-level entry points ------------------------------------------------------------------------
int mbl_extract(mbl_ctx* c, const mbl_batch* b, uint64_t* value, uint64_t* qinfo, size_t cap, size_t* n_out) {
    if (!c || !b || !n_out) return fail(c, MBL_E_BAD_ARG, "null argument");
    int rc = mbl_upload_batch(c, b);
    if (rc != MBL_OK) return rc;
    try {
        uint64_t total = 0;
        for (auto& s : c->subs) total += s.slots;
        *n_out = total;
        if (total > cap || !value || !qinfo) return fail(c, MBL_E_CAPACITY, "k-mer buffers too small");
        cudaStream_t st = c->st;
        const uint32_t n = b->n_reads;
        int32_t *cov1 = c->cov1.get<int32_t>(n), *cov2 = c->cov2.get<int32_t>(n), *w1 = c->w1.get<int32_t>(n), *w2 = c->w2.get<int32_t>(n);
        uint64_t *slots = c->slots.get<uint64_t>(n + 1), *slot_off = c->slot_off.get<uint64_t>(n + 1);
        uint32_t* quot_cnt = c->quot_cnt.get<uint32_t>(n + 1);
        unsigned long long* counters = c->counters.get<unsigned long long>(8);
        MBL_CUDA(cudaMemsetAsync(counters, 0, 64, st));
        MBL_CUDA(cudaMemsetAsync(slots + n, 0, 8, st));
        const uint64_t* off2 = c->paired ? (const uint64_t*)c->off2.p : nullptr;
        launch_read_meta((const uint64_t*)c->off1.p, off2, n, cov1, cov2, w1, w2, slots, quot_cnt, st);
        void* tmp = c->cub_tmp.get<uint8_t>(scan_temp_bytes(n + 1));
        exclusive_sum_u64(tmp, c->cub_tmp.cap, slots, slot_off, n + 1, st);
        uint64_t *va = c->val_a.get<uint64_t>(total), *qa = c->qi_a.get<uint64_t>(total);
        launch_extract(c->cfg.kmer_format, (const uint8_t*)c->bases1.p, (const uint64_t*)c->off1.p,
                       c->paired ? (const uint8_t*)c->bases2.p : nullptr, off2, n, cov1, w1, w2, slot_off, c->d_base_code,
                       c->d_codon, va, qa, nullptr, counters, c->sm_count, st, AaFilter(), nullptr, 0, c->cfg.syncmer ? c->cfg.smer_len : 0);
        MBL_CUDA(cudaMemcpyAsync(value, va, 8 * total, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaMemcpyAsync(qinfo, qa, 8 * total, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        MBL_CUDA(cudaGetLastError());
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_sort_kmers(mbl_ctx* c, uint64_t* value, uint64_t* qinfo, size_t n) {
    if (!c || (n && (!value || !qinfo))) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (!n) return MBL_OK;
    try {
        cudaStream_t st = c->st;
        uint64_t *va = c->val_a.get<uint64_t>(n), *vb = c->val_b.get<uint64_t>(n), *qa = c->qi_a.get<uint64_t>(n), *qb = c->qi_b.get<uint64_t>(n);
        MBL_CUDA(cudaMemcpyAsync(va, value, 8 * n, cudaMemcpyHostToDevice, st));
        MBL_CUDA(cudaMemcpyAsync(qa, qinfo, 8 * n, cudaMemcpyHostToDevice, st));
        void* tmp = c->cub_tmp.get<uint8_t>(sort_kmers_temp_bytes(n));
        int in_b = 0;
        sort_kmers(tmp, c->cub_tmp.cap, va, vb, qa, qb, n, in_b, st);
        MBL_CUDA(cudaMemcpyAsync(value, in_b ? vb : va, 8 * n, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaMemcpyAsync(qinfo, in_b ? qb : qa, 8 * n, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_match(mbl_ctx* c, const uint64_t* value, const uint64_t* qinfo, size_t n, mbl_match_rec* out, size_t cap, size_t* n_match) {
    if (!c || !n_match || (n && (!value || !qinfo))) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (!c->db_loaded) return fail(c, MBL_E_BAD_ARG, "mbl_load_db has not been called");
    try {
        cudaStream_t st = c->st;
        // blanks (UINT64_MAX) sort last; the merge only looks at the non-blank prefix
        size_t nq = n;
        while (nq > 0 && value[nq - 1] == kBlank) --nq;
        uint64_t *va = c->val_a.get<uint64_t>(n + 1), *qa = c->qi_a.get<uint64_t>(n + 1);
        if (n) {
            MBL_CUDA(cudaMemcpyAsync(va, value, 8 * n, cudaMemcpyHostToDevice, st));
            MBL_CUDA(cudaMemcpyAsync(qa, qinfo, 8 * n, cudaMemcpyHostToDevice, st));
        }
        unsigned long long* counters = c->counters.get<unsigned long long>(8);
        MergeArgs ma{};
        ma.diff = c->d_diff; ma.info = c->d_info;
        ma.info_mask = ~((uint32_t)(c->cfg.skip_redundancy == 0) << 31);
        ma.tiles = c->dir.tiles; ma.n_tiles = c->dir.n_tiles; ma.cell_k = c->dir.cell_k; ma.cell_v = c->dir.cell_v;
        ma.jumbo_vals = c->dir.jumbo_vals;
        ma.q_value = va; ma.q_info = qa; ma.n_query = nq;
        ma.taxid2species = c->tax.taxid2species; ma.max_taxid = c->tax.max_taxid;
        ma.ham_pair = c->d_ham_pair; ma.kmer_format = c->cfg.kmer_format;
        ma.ham_single = c->d_ham_single; ma.max_u16 = c->dir.max_u16 + 16; ma.max_kmers = c->dir.max_kmers;
        ma.n_buckets = 2; while (ma.n_buckets <= ma.max_kmers) ma.n_buckets <<= 1;    // hash table load <= 50 %
        ma.out_count = counters + 1;
        ma.error_flag = reinterpret_cast<unsigned int*>(counters + 3);
        ma.item_cursor = reinterpret_cast<unsigned int*>(counters + 3) + 1;
        ma.q_lo = c->q_lo.get<uint64_t>(2 * c->dir.n_tiles + 2);
        ma.cta_threads = c->merge_threads;
        ma.prefix_shift = 24;                // the stage API takes fully ordered queries; any coarser grouping is valid too
        ma.item_cnt = c->item_cnt.get<uint32_t>(c->dir.n_tiles + 2);
        ma.item_off = c->item_off.get<uint32_t>(c->dir.n_tiles + 2);
        ma.items_cap = c->dir.n_tiles + nq / kItemQueries + 2;
        ma.items = c->items.get<MergeItem>(ma.items_cap);
        ma.scan_tmp = c->cub_tmp.get<uint8_t>(scan_temp_bytes(c->dir.n_tiles + 2));
        ma.scan_tmp_bytes = c->cub_tmp.cap;
        uint64_t dcap = cap + out_slack(c);
        unsigned long long h_cnt[4];
        std::vector<mbl_match_rec> host;
        for (int attempt = 0;; ++attempt) {
            ma.out = c->m_raw.get<mbl_match_rec>(dcap);
            ma.out_cap = dcap;
            MBL_CUDA(cudaMemsetAsync(counters, 0, 64, st));
            if (nq && c->dir.n_tiles) { launch_merge_plan(ma, st); launch_merge(ma, c->sm_count, st); }
            MBL_CUDA(cudaMemcpyAsync(h_cnt, counters, 32, cudaMemcpyDeviceToHost, st));
            MBL_CUDA(cudaStreamSynchronize(st));
            MBL_CUDA(cudaGetLastError());
            if ((uint32_t)h_cnt[3] & 1u) return fail(c, MBL_E_BAD_DB, "target k-mer with taxid 0 or unmapped species");
            if (h_cnt[1] <= dcap) break;
            if (attempt > 3) return fail(c, MBL_E_MATCH_OVERFLOW, "match buffer overflow persists");
            dcap = h_cnt[1] + 65536;
        }
        *n_match = h_cnt[2];
        if (h_cnt[2] > cap || !out) return fail(c, MBL_E_MATCH_OVERFLOW, "match buffer too small");
        host.resize(h_cnt[1]);
        if (h_cnt[1]) MBL_CUDA(cudaMemcpy(host.data(), c->m_raw.p, sizeof(mbl_match_rec) * h_cnt[1], cudaMemcpyDeviceToHost));
        size_t w = 0;
        for (const mbl_match_rec& m : host)
            if (qi_seq(m.qinfo) != 0) out[w++] = m;       // drop the blank tails of the warp chunks
        if (w != h_cnt[2]) return fail(c, MBL_E_CUDA, "internal: match count mismatch");
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_sort_matches(mbl_ctx* c, mbl_match_rec* m, size_t n) {
    if (!c || (n && !m)) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (!n) return MBL_OK;
    if (n >= (1ull << 32)) return fail(c, MBL_E_UNSUPPORTED, "too many matches");
    try {
        cudaStream_t st = c->st;
        uint32_t max_seq = 0, max_pos = 0;
        int32_t max_sp = 1;
        for (size_t i = 0; i < n; ++i) {
            max_seq = std::max(max_seq, qi_seq(m[i].qinfo));
            max_pos = std::max(max_pos, qi_pos(m[i].qinfo));
            max_sp = std::max(max_sp, m[i].species_id);
        }
        mbl_match_rec* raw = c->m_raw.get<mbl_match_rec>(n + 1);
        mbl_match_rec* sorted = c->m_sorted.get<mbl_match_rec>(n + 1);
        MBL_CUDA(cudaMemcpyAsync(raw, m, sizeof(mbl_match_rec) * n, cudaMemcpyHostToDevice, st));
        void* tmp = c->cub_tmp.get<uint8_t>(sort_matches_temp_bytes(n));
        // with segment arrays the two-level ordering (seqID radix sort + per-read ordering) runs, as in the pipeline
        sort_matches(tmp, c->cub_tmp.cap, raw, sorted, n, max_seq, max_sp, max_pos, false, c->key_a.get<uint64_t>(n + 1),
                     c->key_b.get<uint64_t>(n + 1), c->idx_a.get<uint32_t>(n + 1), c->idx_b.get<uint32_t>(n + 1), st,
                     c->seg_b.get<uint64_t>((size_t)max_seq + 1), c->seg_e.get<uint64_t>((size_t)max_seq + 1));
        MBL_CUDA(cudaMemcpyAsync(m, sorted, sizeof(mbl_match_rec) * n, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        MBL_CUDA(cudaGetLastError());
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

int mbl_score(mbl_ctx* c, const mbl_match_rec* sorted_h, size_t M, uint32_t n, const int32_t* cov_len1, const int32_t* cov_len2,
              mbl_read_result* out, int32_t* taxcnt_pairs, size_t cap_pairs, size_t* used_pairs) {
    if (!c || !out || !cov_len1 || (M && !sorted_h)) return fail(c, MBL_E_BAD_ARG, "null argument");
    if (!c->db_loaded) return fail(c, MBL_E_BAD_ARG, "mbl_load_db has not been called");
    try {
        cudaStream_t st = c->st;
        std::vector<uint32_t> qcnt(n + 1, 0);
        std::vector<int32_t> zero(n, 0);
        uint64_t quots = 0;
        for (uint32_t r = 0; r < n; ++r) {
            int ql = cov_len1[r] + (cov_len2 ? cov_len2[r] : 0);
            qcnt[r] = ql + 3 > 0 ? (uint32_t)((ql + 3) / 3 + 1) : 1u;
            quots += qcnt[r];
        }
        int32_t *cov1 = c->cov1.get<int32_t>(n), *cov2 = c->cov2.get<int32_t>(n);
        uint32_t *quot_cnt = c->quot_cnt.get<uint32_t>(n + 1), *quot_off = c->quot_off.get<uint32_t>(n + 1);
        MBL_CUDA(cudaMemcpyAsync(cov1, cov_len1, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        MBL_CUDA(cudaMemcpyAsync(cov2, cov_len2 ? cov_len2 : zero.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        MBL_CUDA(cudaMemcpyAsync(quot_cnt, qcnt.data(), 4 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
        void* tmp = c->cub_tmp.get<uint8_t>(scan_temp_bytes(n + 1));
        exclusive_sum_u32(tmp, c->cub_tmp.cap, quot_cnt, quot_off, n + 1, st);
        mbl_match_rec* sorted = c->m_sorted.get<mbl_match_rec>(M + 1);
        if (M) MBL_CUDA(cudaMemcpyAsync(sorted, sorted_h, sizeof(mbl_match_rec) * M, cudaMemcpyHostToDevice, st));
        uint64_t *seg_b = c->seg_b.get<uint64_t>(n + 1), *seg_e = c->seg_e.get<uint64_t>(n + 1);
        launch_segments(sorted, M, n, seg_b, seg_e, st);
        ScoreArgs sa{};
        sa.matches = sorted; sa.n_match = M; sa.n_reads = n; sa.seg_begin = seg_b; sa.seg_end = seg_e;
        sa.cov1 = cov1; sa.cov2 = cov2; sa.quot_off = quot_off; sa.tax = c->tax;
        sa.par.min_score = c->cfg.min_score; sa.par.min_sp_score = c->cfg.min_sp_score; sa.par.tie_ratio = c->cfg.tie_ratio;
        sa.par.min_cons_cnt = c->cfg.min_cons_cnt; sa.par.min_cons_cnt_euk = c->cfg.min_cons_cnt_euk;
        sa.par.accession_level = c->cfg.accession_level;
        sa.par.denominator = (c->cfg.seq_mode == 1 || c->cfg.seq_mode == 2) ? 100 : 1000;
        sa.par.kmer_format = c->cfg.kmer_format;
        sa.par.max_codon_shift = c->cfg.syncmer ? 8 - c->cfg.smer_len : 1;
        sa.par.dna_shift = 3 * sa.par.max_codon_shift;
        const size_t Mp = M + 1;
        sa.l_score = c->l_score.get<float>(Mp); sa.l_start = c->l_start.get<int32_t>(Mp); sa.l_ham = c->l_ham.get<int32_t>(Mp);
        sa.l_depth = c->l_depth.get<int32_t>(Mp); sa.l_smatch = c->l_smatch.get<uint32_t>(Mp); sa.l_conn = c->l_conn.get<uint8_t>(Mp);
        sa.p_start = c->p_start.get<int32_t>(Mp); sa.p_end = c->p_end.get<int32_t>(Mp); sa.p_score = c->p_score.get<float>(Mp);
        sa.p_ham = c->p_ham.get<int32_t>(Mp); sa.p_depth = c->p_depth.get<int32_t>(Mp); sa.p_smatch = c->p_smatch.get<uint32_t>(Mp);
        sa.p_ematch = c->p_ematch.get<uint32_t>(Mp); sa.c_start = c->c_start.get<int32_t>(Mp); sa.c_end = c->c_end.get<int32_t>(Mp);
        sa.s_score = c->s_score.get<float>(Mp);
        sa.q_tax = c->q_tax.get<int32_t>(quots + 1); sa.q_ham = c->q_ham.get<uint8_t>(quots + 1); sa.q_has = c->q_has.get<uint8_t>(quots + 1);
        sa.taxcnt_pairs = c->pairs_raw.get<int32_t>(2 * (quots + 1));
        sa.results = c->res_sub.get<mbl_read_result>(n);
        launch_score(sa, st);
        uint32_t *tl = c->tax_len.get<uint32_t>(n + 1), *to = c->tax_off.get<uint32_t>(n + 1);
        launch_taxcnt_len(sa.results, n, tl, st);
        exclusive_sum_u32(c->cub_tmp.p, c->cub_tmp.cap, tl, to, n + 1, st);
        uint32_t total = 0;
        MBL_CUDA(cudaMemcpyAsync(&total, to + n, 4, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        if (used_pairs) *used_pairs = total;
        if (total > cap_pairs) return fail(c, MBL_E_CAPACITY, "taxcnt_pairs too small");
        int32_t* pc = c->pairs.get<int32_t>(2 * (size_t)total + 2);
        mbl_read_result* rc = c->results.get<mbl_read_result>(n + 1);
        launch_compact_taxcnt(sa.results, n, quot_off, sa.taxcnt_pairs, to, 0u, pc, rc, st);
        MBL_CUDA(cudaMemcpyAsync(out, rc, sizeof(mbl_read_result) * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (total && taxcnt_pairs) MBL_CUDA(cudaMemcpyAsync(taxcnt_pairs, pc, 8 * (size_t)total, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        MBL_CUDA(cudaGetLastError());
    } catch (const CudaError& e) {
        return fail_cuda(c, e);
    } MBL_CATCH_HOST(c)
    return MBL_OK;
}

}  // extern "C"
