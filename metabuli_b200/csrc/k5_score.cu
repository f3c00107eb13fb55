// K5 — per-read taxon scoring kernel (reference rows A10-A12, Classifier.cpp:166-208 +
// Taxonomer.cpp).  One read per thread over the match list sorted in the reference's order; the
// algorithm itself is in score_core.cuh (shared with the CPU unit tests).  Read-level parallelism is
// ample (10^6-10^7 reads per batch); all state is in HBM scratch sized by the match count.
#include <cub/cub.cuh>

#include "score_core.cuh"

namespace mbl {

__global__ void __launch_bounds__(128, 4) score_kernel(ScoreArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_reads) return;
    score_read(a, a.read_perm ? a.read_perm[a.read_begin + i] : a.read_begin + i);
}

// ---- flat pipeline ---------------------------------------------------------------------------------------------
// n_long counts the frame groups with at least min_rows rows — the only ones getMatchPaths can emit a path for (min_group_rows);
// after the length ordering below they are the first n_long tasks, and only those get a thread
__global__ void score_mark_kernel(const mbl_match_rec* __restrict__ m, uint64_t begin, uint64_t end, uint8_t* __restrict__ flag_fg,
                                  uint8_t* __restrict__ flag_sp, uint32_t min_rows, uint32_t* __restrict__ n_long) {
    const uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool is_long = false;
    if (i < end) {
        bool sp = true, fg = true;
        const uint64_t q = m[i].qinfo;
        const int32_t species = m[i].species_id;
        if (i > begin) {
            const uint64_t pq = m[i - 1].qinfo;
            sp = qi_seq(q) != qi_seq(pq) || species != m[i - 1].species_id;
            fg = sp || qi_frame(q) != qi_frame(pq);
        }
        flag_sp[i - begin] = sp;
        flag_fg[i - begin] = fg;
        if (fg && i + min_rows - 1 < end) {              // sorted by (read, species, frame): same group <=> same triple min_rows - 1 rows on
            const uint64_t lq = m[i + min_rows - 1].qinfo;
            is_long = qi_seq(lq) == qi_seq(q) && qi_frame(lq) == qi_frame(q) && m[i + min_rows - 1].species_id == species;
        }
    }
    const int cnt = __syncthreads_count(is_long);          // one atomic per block: ~10^9 rows must not queue up on one address
    if (cnt && threadIdx.x == 0) atomicAdd(n_long, (uint32_t)cnt);
}
// first species task of every read that has matches (read_sp[seqID - 1]): spares score_read two binary searches over sp_list
__global__ void read_first_species_kernel(const mbl_match_rec* __restrict__ m, const uint32_t* __restrict__ sp_list, uint32_t n_sp,
                                          uint32_t* __restrict__ read_sp) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sp) return;
    const uint32_t seq = qi_seq(m[sp_list[s]].qinfo);
    if (s == 0 || qi_seq(m[sp_list[s - 1]].qinfo) != seq) read_sp[seq - 1] = s;
}
// flag of frame-group task g: it is the first one of its species group (both lists come from the same flags)
struct SpeciesStartOfGroup {
    const uint8_t* flag_sp;
    const uint32_t* fg_list;
    uint32_t match_begin;
    __host__ __device__ uint8_t operator()(uint32_t g) const { return flag_sp[fg_list[g] - match_begin]; }
};
// Frame groups differ a lot in length (the true species in the true frame holds tens of matches, chance hits one or
// two), so one group per thread in list order leaves most lanes of a warp idle.  Tasks are therefore ordered by length,
// longest first: key = 255 - min(length, 255), one 8-bit radix pass.
__global__ void fg_len_key_kernel(const uint32_t* __restrict__ fg_list, uint32_t n_fg, uint64_t match_end, uint8_t* __restrict__ key,
                                  uint32_t* __restrict__ idx) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_fg) return;
    const uint64_t len = (g + 1 < n_fg ? (uint64_t)fg_list[g + 1] : match_end) - fg_list[g];
    key[g] = (uint8_t)(255u - (uint32_t)min(len, (uint64_t)255));
    idx[g] = g;
}
__global__ void __launch_bounds__(128, 8) score_fg_kernel(ScoreArgs a, uint32_t n_tasks) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_tasks) score_task_frame_group(a, g);
}
__global__ void __launch_bounds__(128, 4) score_sp_kernel(ScoreArgs a) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < a.n_sp) score_task_species(a, s);
}

size_t score_flat_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, bytes, cub::CountingInputIterator<uint32_t>(0), (const uint8_t*)nullptr, (uint32_t*)nullptr,
                               (uint32_t*)nullptr, (long long)n);
    size_t sb = 0;
    cub::DoubleBuffer<uint8_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, sb, k, v, (long long)n, 0, 8);
    return bytes > sb ? bytes : sb;
}

void launch_score_flat(ScoreArgs a, uint64_t match_begin, const ScoreFlatScratch& s, cudaStream_t st) {
    const uint64_t n = a.match_end - match_begin;
    if (a.n_reads == 0) return;
    uint32_t h_counts[3] = {0, 0, 0};
    if (n) {
        MBL_CUDA(cudaMemsetAsync(s.counts + 2, 0, 4, st));
        score_mark_kernel<<<(unsigned)((n + 1023) / 1024), 1024, 0, st>>>(a.matches, match_begin, a.match_end, s.flags_fg, s.flags_sp,
                                                                      (uint32_t)min(min_group_rows(a.par), 255), s.counts + 2);   // 255: the length key saturates there
        size_t tb = s.cub_tmp_bytes;
        MBL_CUDA(cub::DeviceSelect::Flagged(s.cub_tmp, tb, cub::CountingInputIterator<uint32_t>((uint32_t)match_begin), s.flags_fg, s.fg_list,
                                            s.counts, (long long)n, st));
        tb = s.cub_tmp_bytes;
        MBL_CUDA(cub::DeviceSelect::Flagged(s.cub_tmp, tb, cub::CountingInputIterator<uint32_t>((uint32_t)match_begin), s.flags_sp, s.sp_list,
                                            s.counts + 1, (long long)n, st));
        MBL_CUDA(cudaMemcpyAsync(h_counts, s.counts, 12, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
    }
    a.fg_list = s.fg_list; a.n_fg = h_counts[0]; a.sp_list = s.sp_list; a.n_sp = h_counts[1];
    a.fg_order = nullptr;
    a.sp_fg = nullptr; a.read_sp = nullptr; a.sp_score = nullptr;
    if (a.n_fg && a.n_sp && s.sp_fg && s.read_sp && s.sp_score) {
        // direct maps instead of binary searches: species task -> its first frame-group task (a select over the frame-group tasks
        // whose flag is "starts a species group"; must run before the flags are reused as sort keys), read -> its first species task
        cub::TransformInputIterator<uint8_t, SpeciesStartOfGroup, cub::CountingInputIterator<uint32_t>> flags(
            cub::CountingInputIterator<uint32_t>(0), SpeciesStartOfGroup{s.flags_sp, s.fg_list, (uint32_t)match_begin});
        size_t tb = s.cub_tmp_bytes;
        MBL_CUDA(cub::DeviceSelect::Flagged(s.cub_tmp, tb, cub::CountingInputIterator<uint32_t>(0), flags, s.sp_fg, s.counts + 3, (long long)a.n_fg, st));
        read_first_species_kernel<<<(a.n_sp + 255) / 256, 256, 0, st>>>(a.matches, s.sp_list, a.n_sp, s.read_sp);
        a.sp_fg = s.sp_fg; a.read_sp = s.read_sp; a.sp_score = s.sp_score;
    }
    if (a.n_fg > 1) {
        fg_len_key_kernel<<<(a.n_fg + 255) / 256, 256, 0, st>>>(s.fg_list, a.n_fg, a.match_end, s.flags_fg, s.fg_ord);
        cub::DoubleBuffer<uint8_t> k(s.flags_fg, s.flags_sp);
        cub::DoubleBuffer<uint32_t> v(s.fg_ord, s.fg_ord + n);
        size_t tb = s.cub_tmp_bytes;
        MBL_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_tmp, tb, k, v, (long long)a.n_fg, 0, 8, st));
        a.fg_order = v.Current();
    }
    // tasks are ordered longest group first, so the groups that can emit a path are the first n_long ones (a single group: no order)
    const uint32_t n_tasks = a.fg_order ? (h_counts[2] < a.n_fg ? h_counts[2] : a.n_fg) : a.n_fg;
    if (n_tasks) score_fg_kernel<<<(n_tasks + 127) / 128, 128, 0, st>>>(a, n_tasks);
    if (a.n_sp) score_sp_kernel<<<(a.n_sp + 127) / 128, 128, 0, st>>>(a);
    score_kernel<<<(a.n_reads + 127) / 128, 128, 0, st>>>(a);
}

__global__ void taxcnt_len_kernel(const mbl_read_result* __restrict__ res, uint32_t n, uint32_t* __restrict__ len) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) len[r] = res[r].taxcnt_len;
    if (r == n) len[r] = 0;
}

__global__ void compact_taxcnt_kernel(const mbl_read_result* __restrict__ res, uint32_t n, const uint32_t* __restrict__ quot_off,
                                      const int32_t* __restrict__ pairs_in, const uint32_t* __restrict__ out_off, uint32_t pair_base,
                                      int32_t* __restrict__ pairs_out, mbl_read_result* __restrict__ res_out) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    mbl_read_result x = res[r];
    const int32_t* src = pairs_in + 2ull * quot_off[r];
    int32_t* dst = pairs_out + 2ull * out_off[r];
    for (uint32_t k = 0; k < 2 * x.taxcnt_len; ++k) dst[k] = src[k];
    x.taxcnt_begin = pair_base + out_off[r];         // batch-global: pairs of earlier sub-batches come first
    res_out[r] = x;
}

__global__ void shift_taxcnt_kernel(mbl_read_result* __restrict__ res, uint32_t n, uint32_t delta) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) res[r].taxcnt_begin += delta;
}
void launch_shift_taxcnt(mbl_read_result* results, uint32_t n_reads, uint32_t delta, cudaStream_t st) {
    if (n_reads) shift_taxcnt_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(results, n_reads, delta);
}

void launch_score(const ScoreArgs& a, cudaStream_t st) {
    if (!a.n_reads) return;
    score_kernel<<<(a.n_reads + 127) / 128, 128, 0, st>>>(a);
}
void launch_taxcnt_len(const mbl_read_result* results, uint32_t n_reads, uint32_t* len, cudaStream_t st) {
    taxcnt_len_kernel<<<(n_reads + 1 + 255) / 256, 256, 0, st>>>(results, n_reads, len);
}
void launch_compact_taxcnt(const mbl_read_result* results, uint32_t n_reads, const uint32_t* quot_off, const int32_t* pairs_in,
                           const uint32_t* out_off, uint32_t pair_base, int32_t* pairs_out, mbl_read_result* results_out, cudaStream_t st) {
    if (!n_reads) return;
    compact_taxcnt_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(results, n_reads, quot_off, pairs_in, out_off, pair_base, pairs_out, results_out);
}

}  // namespace mbl
