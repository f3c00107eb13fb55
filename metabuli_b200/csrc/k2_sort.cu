// K2 / K4 — device radix sorts (reference rows A4 and A9: KmerExtractor.cpp:79 with Kmer::compareQueryKmer,
// Kmer.h:89-94; KmerMatcher.cpp:1071-1078 with compareMatches, :1149-1166).  The reference uses ips4o on
// 16-byte / 32-byte structs; here keys and payloads are split (SoA) and run through CUB's onesweep radix
// sort on exactly the bits that carry order:
//   query metamers : the 40-bit amino-acid part only (bits 24..63).  The merge kernel treats an
//                    amino-acid group as a set, so order inside a group is free.
//   matches        : one radix sort of a 32-bit permutation on (seqID, species, frame, pos) packed into the
//                    bits the batch actually uses (48 for 150-bp reads), then the short runs that share that
//                    key are ordered by (hamming, dna) in place — together the reference's total order.
//                    Batches whose packed key would not fit 64 bits fall back to two stable LSD passes.
#include <algorithm>
#include <cstdlib>
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace mbl {

size_t sort_kmers_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (long long)n, 24, 64);
    return bytes;
}

void sort_kmers(void* tmp, size_t tmp_bytes, uint64_t* key_a, uint64_t* key_b, uint64_t* val_a, uint64_t* val_b, size_t n,
                int& result_in_b, cudaStream_t st) {
    cub::DoubleBuffer<uint64_t> k(key_a, key_b), v(val_a, val_b);
    MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, 24, 64, st));
    result_in_b = k.selector;
}

void sort_kmers_qinfo(void* tmp, size_t tmp_bytes, uint64_t* key_a, uint64_t* key_b, uint64_t* val_a, uint64_t* val_b, size_t n, int begin_bit,
                      int& result_in_b, cudaStream_t st) {
    cub::DoubleBuffer<uint64_t> k(key_a, key_b), v(val_a, val_b);
    MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, begin_bit, 64, st));
    result_in_b = k.selector;
}

void sort_kmers_idx(void* tmp, size_t tmp_bytes, uint64_t* key_a, uint64_t* key_b, uint32_t* idx_a, uint32_t* idx_b, size_t n, int begin_bit,
                    int& result_in_b, cudaStream_t st) {
    cub::DoubleBuffer<uint64_t> k(key_a, key_b);
    cub::DoubleBuffer<uint32_t> v(idx_a, idx_b);
    MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, begin_bit, 64, st));
    result_in_b = k.selector;
}

size_t scan_temp_bytes(size_t n) {
    size_t a = 0, b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (long long)n);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (long long)n);
    return a > b ? a : b;
}
void exclusive_sum_u64(void* tmp, size_t tmp_bytes, const uint64_t* in, uint64_t* out, size_t n, cudaStream_t st) {
    MBL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (long long)n, st));
}
void exclusive_sum_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st) {
    MBL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (long long)n, st));
}

// ---- match ordering ---------------------------------------------------------------------------------
namespace {

__device__ __host__ inline int bits_for(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

// low key: frame | pos | hamming(3 bits, <= 7 by construction) | dna(24)
__global__ void match_lowkey_kernel(const mbl_match_rec* __restrict__ m, size_t n, int pos_bits, uint64_t* __restrict__ key,
                                    uint32_t* __restrict__ idx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t q = m[i].qinfo;
    uint64_t k = (uint64_t)qi_frame(q);
    k = (k << pos_bits) | (uint64_t)qi_pos(q);
    k = (k << 3) | (uint64_t)(m[i].hamming & 7u);
    k = (k << 24) | (uint64_t)(m[i].dna_encoding & 0xFFFFFFu);
    key[i] = k;
    idx[i] = (uint32_t)i;
}
// high key: seqID | species, gathered through the permutation of the first pass
__global__ void match_highkey_kernel(const mbl_match_rec* __restrict__ m, const uint32_t* __restrict__ idx, size_t n, int sp_bits,
                                     uint64_t* __restrict__ key) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mbl_match_rec& x = m[idx[i]];
    key[i] = ((uint64_t)qi_seq(x.qinfo) << sp_bits) | (uint64_t)(uint32_t)x.species_id;
}
__global__ void match_gather_kernel(const mbl_match_rec* __restrict__ in, const uint32_t* __restrict__ idx, size_t n,
                                    mbl_match_rec* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t* s = reinterpret_cast<const uint64_t*>(in + idx[i]);
    uint64_t* d = reinterpret_cast<uint64_t*>(out + i);
    uint64_t a = s[0], b = s[1], c = s[2];
    d[0] = a; d[1] = b; d[2] = c;
}
// single-pass key: seqID | species | frame | pos / 3.  Inside one (read, frame) the k-mer positions advance in codon steps
// (and the second mate starts beyond the first), so two distinct positions differ by at least 3 and pos / 3 orders them
// like pos does with ~1.6 fewer key bits — in the benchmark shape that is one radix pass less.
__global__ void match_fullkey_kernel(const mbl_match_rec* __restrict__ m, size_t n, int sp_bits, int pos_bits, uint32_t pos_div,
                                     uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t q = m[i].qinfo;
    uint64_t k = (uint64_t)qi_seq(q);
    k = (k << sp_bits) | (uint64_t)(uint32_t)m[i].species_id;
    k = (k << 3) | (uint64_t)qi_frame(q);
    k = (k << pos_bits) | (uint64_t)(qi_pos(q) / pos_div);
    key[i] = k;
    idx[i] = (uint32_t)i;
}
// Fused tail of the single-key path: gather the 24-byte rows through the sorted permutation, order the short runs that
// share a key by (hamming, dna) and record every read's segment — all decided from the sorted KEYS (coalesced), so the rows
// are touched exactly once (random read, coalesced write).  Equal keys <=> equal (seqID, species, frame, position); key 0 is
// a blank row (seqID 0).  The thread of a run's first element places the whole run.
__global__ void match_gather_fix_kernel(const mbl_match_rec* __restrict__ in, const uint32_t* __restrict__ idx,
                                        const uint64_t* __restrict__ key, size_t n, int seq_shift, uint32_t n_reads,
                                        mbl_match_rec* __restrict__ out, uint64_t* __restrict__ seg_begin, uint64_t* __restrict__ seg_end) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = key[i];
    const uint64_t kp = i > 0 ? key[i - 1] : ~k, kn = i + 1 < n ? key[i + 1] : ~k;
    const uint64_t seq = k >> seq_shift;
    if (seg_begin && seq != 0 && seq <= n_reads) {
        if ((kp >> seq_shift) != seq) seg_begin[seq - 1] = i;
        if ((kn >> seq_shift) != seq) seg_end[seq - 1] = i + 1;
    }
    auto copy = [&](size_t from, size_t to) {
        const uint64_t* s = reinterpret_cast<const uint64_t*>(in + from);
        uint64_t* d = reinterpret_cast<uint64_t*>(out + to);
        const uint64_t a = s[0], b = s[1], c = s[2];
        d[0] = a; d[1] = b; d[2] = c;
    };
    if (k == 0 || (k != kp && k != kn)) { copy(idx[i], i); return; }      // blank row or a run of one
    if (k == kp) return;                                                 // placed by the run's first thread
    size_t e = i + 2;
    while (e < n && key[e] == k) ++e;
    // selection by rank: row r of the run goes to i + (number of rows of the run that order before it)
    for (size_t r = i; r < e; ++r) {
        const mbl_match_rec x = in[idx[r]];
        size_t rank = 0;
        for (size_t q = i; q < e; ++q) {
            if (q == r) continue;
            const mbl_match_rec y = in[idx[q]];
            const bool before = y.hamming != x.hamming ? y.hamming < x.hamming
                              : (y.dna_encoding != x.dna_encoding ? y.dna_encoding < x.dna_encoding : q < r);
            rank += before ? 1 : 0;
        }
        copy(idx[r], i + rank);
    }
}
// ---- two-level ordering (short-read batches) -------------------------------------------------------------------------------------
// compareMatches (KmerMatcher.cpp:1149-1166) orders by (seqID, species, frame, pos, hamming, dna).  Only the seqID needs a global
// sort: three 8-bit radix passes over (32-bit seqID, 32-bit row index) instead of six over (64-bit key, index).  The rest of the
// order is local to a read's ~100-150 rows: one warp per read gathers the rows through the permutation (the one random read of
// the rows that any ordering needs), radix-sorts (species | frame | pos) keys in shared memory — 8-bit digits, histogram by
// shared-memory atomics, stable ranks from match.any ballots, ~250 warp instructions per pass — orders the rare rows that share
// that key by (hamming, dna), and writes the rows to their final places.  (Ranking by counting, n^2 / 32 steps, and a bitonic
// network, ~3 k instructions per read, both lost to the six global passes: 99-106 ms against 74.6 ms.)
constexpr uint32_t kOrderSmallRows = 256;       // the tier most reads with real hits fall into (launch_match_order)
constexpr uint32_t kOrderMaxRows = 2048;        // rows of one read a warp can order in shared memory; beyond that => single-key path

__global__ void match_seqkey_kernel(const mbl_match_rec* __restrict__ m, size_t n, uint32_t* __restrict__ key, uint32_t* __restrict__ idx) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = qi_seq(m[i].qinfo);
    idx[i] = (uint32_t)i;
}
// per-read segments from the sorted seqIDs; blank rows (seqID 0) sort first, n_blank[0] = their count
__global__ void seq_segments_kernel(const uint32_t* __restrict__ key, size_t n, uint32_t n_reads, uint64_t* __restrict__ seg_begin,
                                    uint64_t* __restrict__ seg_end, unsigned long long* __restrict__ n_blank) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = key[i];
    const uint32_t sp = i > 0 ? key[i - 1] : ~s, sn = i + 1 < n ? key[i + 1] : ~s;
    if (s == 0) { if (sn != 0) *n_blank = i + 1; return; }
    if (s > n_reads) return;
    if (sp != s) seg_begin[s - 1] = i;
    if (sn != s) seg_end[s - 1] = i + 1;
}
__global__ void seg_maxlen_kernel(const uint64_t* __restrict__ seg_begin, const uint64_t* __restrict__ seg_end, uint32_t n_reads,
                                  unsigned long long* __restrict__ max_len) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long len = r < n_reads ? seg_end[r] - seg_begin[r] : 0ull;
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len) atomicMax(max_len, len);
}

__device__ __forceinline__ void order_cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// reads with kMinRows < rows <= kMaxRows; KeyT holds species | frame | pos / pos_div (local_bits bits).  A warp takes 32
// consecutive reads at a time (segments loaded coalesced, the reads in range picked by ballot); a read's rows are gathered into
// shared memory with cp.async — every row of the read in flight at once instead of one dependent load chain per 32 rows, which
// made the first version of this kernel latency-bound (45 ms) — sorted there, and written out coalesced.
template <class KeyT, uint32_t kMinRows, uint32_t kMaxRows, int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
match_order_kernel(const mbl_match_rec* __restrict__ in, const uint32_t* __restrict__ idx, const uint64_t* __restrict__ seg_begin,
                   const uint64_t* __restrict__ seg_end, uint32_t n_reads, int pos_bits, uint32_t pos_div, int local_bits,
                   mbl_match_rec* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char order_smem[];
    constexpr size_t kPerWarp = (size_t)kMaxRows * (24 + 2 * sizeof(KeyT) + 2 * 2) + 256 * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned char* base = order_smem + (size_t)warp * kPerWarp;
    uint64_t* rows = reinterpret_cast<uint64_t*>(base);                      // [3 * kMaxRows]
    KeyT* ka = reinterpret_cast<KeyT*>(rows + 3 * (size_t)kMaxRows);
    KeyT* kb = ka + kMaxRows;
    uint32_t* hist = reinterpret_cast<uint32_t*>(kb + kMaxRows);
    uint16_t* oa = reinterpret_cast<uint16_t*>(hist + 256);
    uint16_t* ob = oa + kMaxRows;
    for (uint32_t r0 = (blockIdx.x * kWarps + warp) * 32u; r0 < n_reads; r0 += gridDim.x * kWarps * 32u) {
        uint64_t my_b = 0;
        uint32_t my_n = 0;
        if (r0 + lane < n_reads) { my_b = seg_begin[r0 + lane]; my_n = (uint32_t)(seg_end[r0 + lane] - my_b); }
        uint32_t todo = __ballot_sync(0xffffffffu, my_n > kMinRows && my_n <= kMaxRows);
        while (todo) {
            const int t = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint64_t b = __shfl_sync(0xffffffffu, my_b, t);
            const uint32_t n = __shfl_sync(0xffffffffu, my_n, t);
            if (n == 1) {
                if (lane < 3) reinterpret_cast<uint64_t*>(out + b)[lane] = reinterpret_cast<const uint64_t*>(in + idx[b])[lane];
                continue;
            }
            // gather: the permutation entries of 8 rows per lane at a time, then all their words in flight
            for (uint32_t j0 = 0; j0 < n; j0 += 256) {
                uint32_t ix[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const uint32_t j = j0 + 32u * u + lane; ix[u] = j < n ? idx[b + j] : 0u; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t j = j0 + 32u * u + lane;
                    if (j < n) {
                        const uint64_t* s = reinterpret_cast<const uint64_t*>(in + ix[u]);
                        order_cp_async8(rows + 3 * j, s);
                        order_cp_async8(rows + 3 * j + 1, s + 1);
                        order_cp_async8(rows + 3 * j + 2, s + 2);
                    }
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            for (uint32_t j = lane; j < n; j += 32) {
                const uint64_t q = rows[3 * j], w1 = rows[3 * j + 1];       // qinfo | target, species
                uint64_t k = w1 >> 32;                                        // species
                k = (k << 3) | (q >> 61);                                     // frame
                k = (k << pos_bits) | (uint64_t)((uint32_t)q / pos_div);
                ka[j] = (KeyT)k;
                oa[j] = (uint16_t)j;
            }
            __syncwarp();
            KeyT *src = ka, *dst = kb;
            uint16_t *osrc = oa, *odst = ob;
            for (int shift = 0; shift < local_bits; shift += 8) {
                for (uint32_t k = lane; k < 256; k += 32) hist[k] = 0u;
                __syncwarp();
                for (uint32_t j = lane; j < n; j += 32) atomicAdd(&hist[(uint32_t)(src[j] >> shift) & 255u], 1u);
                __syncwarp();
                {   // exclusive scan of the 256 counts: 8 bins per lane, shuffle scan of the lane sums
                    uint32_t c[8], sum = 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) { c[u] = hist[8 * lane + u]; sum += c[u]; }
                    uint32_t incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                    uint32_t run = incl - sum;
#pragma unroll
                    for (int u = 0; u < 8; ++u) { hist[8 * lane + u] = run; run += c[u]; }
                }
                __syncwarp();
                for (uint32_t j0 = 0; j0 < n; j0 += 32) {            // stable: chunks in order, lanes of a chunk ranked by match.any
                    const uint32_t j = j0 + lane;
                    const bool active = j < n;
                    const KeyT k = active ? src[j] : (KeyT)0;
                    const uint32_t d = active ? ((uint32_t)(k >> shift) & 255u) : (0x10000u | (uint32_t)lane);
                    const uint32_t same = __match_any_sync(0xffffffffu, d);
                    const uint32_t rank = __popc(same & lt);
                    const uint32_t at = active ? hist[d] : 0u;
                    __syncwarp();
                    if (active && rank == 0) hist[d] = at + __popc(same);
                    __syncwarp();
                    if (active) { dst[at + rank] = k; odst[at + rank] = osrc[j]; }
                }
                __syncwarp();
                { KeyT* x = src; src = dst; dst = x; }
                { uint16_t* x = osrc; osrc = odst; odst = x; }
            }
            // rows that share (species, frame, pos): order by (hamming, dna); the first lane of a run does it (runs of 2-3)
            auto tie_of = [&](uint16_t o) { const uint64_t w2 = rows[3 * (uint32_t)o + 2]; return ((uint32_t)((w2 >> 48) & 7ull) << 24) | (uint32_t)(w2 & 0xFFFFFFull); };
            for (uint32_t p = lane; p < n; p += 32) {
                if ((p == 0 || src[p - 1] != src[p]) && p + 1 < n && src[p + 1] == src[p]) {
                    uint32_t e = p + 2;
                    while (e < n && src[e] == src[p]) ++e;
                    for (uint32_t x = p + 1; x < e; ++x) {
                        const uint16_t v = osrc[x];
                        const uint32_t tv = tie_of(v);
                        uint32_t y = x;
                        while (y > p && tie_of(osrc[y - 1]) > tv) { osrc[y] = osrc[y - 1]; --y; }
                        osrc[y] = v;
                    }
                }
            }
            __syncwarp();
            // out: 3 n consecutive 8-byte words, word w of the output = word (w % 3) of row osrc[w / 3]
            uint64_t* d = reinterpret_cast<uint64_t*>(out + b);
            for (uint32_t w = lane; w < 3 * n; w += 32) {
                const uint32_t p = w / 3u;
                d[w] = rows[3 * (uint32_t)osrc[p] + (w - 3 * p)];
            }
            __syncwarp();
        }
    }
}

template <class KeyT>
static void launch_match_order(const mbl_match_rec* in, const uint32_t* idx, const uint64_t* seg_begin, const uint64_t* seg_end, uint32_t n_reads,
                               int pos_bits, uint32_t pos_div, int local_bits, uint64_t max_len, mbl_match_rec* out, cudaStream_t st) {
    // four tiers by rows per read, so that the many reads with a handful of rows (chance hits only) run at full occupancy and only the
    // few long ones pay for a large row buffer: <= 64 rows (8 warps per CTA, 2.8 KB per warp), <= 256 (4 warps, 10 KB), <= 512 (2), <= 2048 (1)
    constexpr uint32_t kTiny = 64, kMid = 512;
    auto bytes = [](uint32_t rows, int warps) { return (size_t)warps * ((size_t)rows * (24 + 2 * sizeof(KeyT) + 2 * 2) + 256 * 4); };
    MBL_CUDA(cudaFuncSetAttribute(match_order_kernel<KeyT, kTiny, kOrderSmallRows, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes(kOrderSmallRows, 4)));
    MBL_CUDA(cudaFuncSetAttribute(match_order_kernel<KeyT, kOrderSmallRows, kMid, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes(kMid, 2)));
    MBL_CUDA(cudaFuncSetAttribute(match_order_kernel<KeyT, kMid, kOrderMaxRows, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes(kOrderMaxRows, 1)));
    const unsigned blocks0 = (unsigned)std::min<uint64_t>((n_reads + 32 * 8 - 1) / (32 * 8), 148ull * 16);
    match_order_kernel<KeyT, 0, kTiny, 8><<<blocks0, 8 * 32, bytes(kTiny, 8), st>>>(in, idx, seg_begin, seg_end, n_reads, pos_bits, pos_div, local_bits, out);
    const unsigned blocks1 = (unsigned)std::min<uint64_t>((n_reads + 32 * 4 - 1) / (32 * 4), 148ull * 32);
    if (max_len > kTiny)
        match_order_kernel<KeyT, kTiny, kOrderSmallRows, 4><<<blocks1, 4 * 32, bytes(kOrderSmallRows, 4), st>>>(in, idx, seg_begin, seg_end, n_reads, pos_bits,
                                                                                                              pos_div, local_bits, out);
    if (max_len > kOrderSmallRows)
        match_order_kernel<KeyT, kOrderSmallRows, kMid, 2><<<148 * 5, 2 * 32, bytes(kMid, 2), st>>>(in, idx, seg_begin, seg_end, n_reads, pos_bits, pos_div,
                                                                                                  local_bits, out);
    if (max_len > kMid)
        match_order_kernel<KeyT, kMid, kOrderMaxRows, 1><<<148 * 3, 1 * 32, bytes(kOrderMaxRows, 1), st>>>(in, idx, seg_begin, seg_end, n_reads, pos_bits,
                                                                                                         pos_div, local_bits, out);
}

// seg_begin/seg_end per read from the sorted match list (Classifier.cpp:174-185 MatchBlocks)
__global__ void segment_kernel(const mbl_match_rec* __restrict__ m, size_t n, uint32_t n_reads, uint64_t* __restrict__ seg_begin,
                               uint64_t* __restrict__ seg_end) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = qi_seq(m[i].qinfo);
    if (s == 0 || s > n_reads) return;
    if (i == 0 || qi_seq(m[i - 1].qinfo) != s) seg_begin[s - 1] = i;
    if (i + 1 == n || qi_seq(m[i + 1].qinfo) != s) seg_end[s - 1] = i + 1;
}

// first match of every chunk of `chunk_reads` reads: bounds[k] = lower_bound(seqID >= k*chunk_reads + 1)
__global__ void seq_bounds_kernel(const mbl_match_rec* __restrict__ m, size_t n, uint32_t chunk_reads, uint32_t n_chunks,
                                  uint64_t* __restrict__ bounds) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_chunks) return;
    if (k == n_chunks) { bounds[k] = n; return; }
    const uint64_t want = (uint64_t)k * chunk_reads + 1;
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if ((uint64_t)qi_seq(m[mid].qinfo) < want) lo = mid + 1; else hi = mid; }
    bounds[k] = lo;
}

// scoring order: inside each chunk of `chunk_reads` reads, reads with similar match counts become neighbours so
// the lanes of a warp (one read per thread in K5) run loops of similar length
__global__ void read_len_key_kernel(const uint64_t* __restrict__ seg_b, const uint64_t* __restrict__ seg_e, uint32_t n, uint32_t chunk_reads,
                                    uint32_t* __restrict__ key, uint32_t* __restrict__ idx) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t len = seg_e[r] - seg_b[r];
    key[r] = ((r / chunk_reads) << 16) | (uint32_t)min(len, (uint64_t)0xFFFF);
    idx[r] = r;
}

}  // namespace

size_t order_reads_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (long long)n, 0, 32);
    return bytes;
}
// -> pointer to the permutation (one of idx_a / idx_b)
const uint32_t* order_reads_by_matches(void* tmp, size_t tmp_bytes, const uint64_t* seg_b, const uint64_t* seg_e, uint32_t n,
                                       uint32_t chunk_reads, uint32_t* key_a, uint32_t* key_b, uint32_t* idx_a, uint32_t* idx_b, cudaStream_t st) {
    if (!n) return idx_a;
    read_len_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(seg_b, seg_e, n, chunk_reads, key_a, idx_a);
    cub::DoubleBuffer<uint32_t> k(key_a, key_b), v(idx_a, idx_b);
    int chunk_bits = 1;
    while ((1u << chunk_bits) <= (n - 1) / chunk_reads) ++chunk_bits;
    MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, 0, 16 + chunk_bits, st));
    return v.Current();
}

void launch_seq_bounds(const mbl_match_rec* sorted, size_t n, uint32_t chunk_reads, uint32_t n_chunks, uint64_t* bounds, cudaStream_t st) {
    seq_bounds_kernel<<<(n_chunks + 1 + 127) / 128, 128, 0, st>>>(sorted, n, chunk_reads, n_chunks, bounds);
}

size_t sort_matches_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (long long)n, 0, 64);
    return bytes;
}

bool sort_matches(void* tmp, size_t tmp_bytes, const mbl_match_rec* in, mbl_match_rec* out, size_t n, uint32_t n_reads,
                  int32_t max_taxid, uint32_t max_pos, bool codon_spaced, uint64_t* key_a, uint64_t* key_b, uint32_t* idx_a,
                  uint32_t* idx_b, cudaStream_t st, uint64_t* seg_begin, uint64_t* seg_end) {
    if (!n) return false;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    const int pos_bits = bits_for(max_pos);
    const int sp_bits = bits_for((uint64_t)(uint32_t)max_taxid);
    const int seq_bits = bits_for(n_reads);
    const uint32_t pos_div = codon_spaced ? 3u : 1u;     // true for matches produced by K3 (see match_fullkey_kernel)
    const int pos3_bits = bits_for(max_pos / pos_div);
    static const bool force_two_pass = getenv("MBL_TEST_TWO_PASS_SORT") && atoi(getenv("MBL_TEST_TWO_PASS_SORT")) != 0;   // tests: keys wider than 64 bits
    static const bool force_fullkey = getenv("MBL_SORT_FULLKEY") && atoi(getenv("MBL_SORT_FULLKEY")) != 0;   // tests / A-B: the single-key path
    const int local_bits = sp_bits + 3 + pos3_bits;
    if (seg_begin && seg_end && !force_fullkey && !force_two_pass && local_bits <= 40) {
        // two-level: global sort by seqID, then per-read ordering fused into the gather
        uint32_t *k32a = reinterpret_cast<uint32_t*>(key_a), *k32b = reinterpret_cast<uint32_t*>(key_b);
        match_seqkey_kernel<<<blocks, 256, 0, st>>>(in, n, k32a, idx_a);
        cub::DoubleBuffer<uint32_t> k(k32a, k32b), v(idx_a, idx_b);
        MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, 0, seq_bits, st));
        MBL_CUDA(cudaMemsetAsync(seg_begin, 0, 8 * (size_t)n_reads, st));
        MBL_CUDA(cudaMemsetAsync(seg_end, 0, 8 * (size_t)n_reads, st));
        unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(tmp);      // the sort is done with its scratch
        MBL_CUDA(cudaMemsetAsync(d_cnt, 0, 16, st));
        seq_segments_kernel<<<blocks, 256, 0, st>>>(k.Current(), n, n_reads, seg_begin, seg_end, d_cnt);
        seg_maxlen_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(seg_begin, seg_end, n_reads, d_cnt + 1);
        unsigned long long h_cnt[2] = {0, 0};
        MBL_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
        MBL_CUDA(cudaStreamSynchronize(st));
        if (h_cnt[1] <= kOrderMaxRows) {
            if (h_cnt[0]) MBL_CUDA(cudaMemsetAsync(out, 0, sizeof(mbl_match_rec) * (size_t)h_cnt[0], st));   // blank rows (seqID 0) come first
            if (local_bits <= 32) launch_match_order<uint32_t>(in, v.Current(), seg_begin, seg_end, n_reads, pos3_bits, pos_div, local_bits, h_cnt[1], out, st);
            else launch_match_order<uint64_t>(in, v.Current(), seg_begin, seg_end, n_reads, pos3_bits, pos_div, local_bits, h_cnt[1], out, st);
            return true;
        }
        // a read with more rows than a warp orders in shared memory (long reads): the single-key path below redoes the order
    }
    if (seq_bits + sp_bits + 3 + pos3_bits <= 64 && !force_two_pass) {
        match_fullkey_kernel<<<blocks, 256, 0, st>>>(in, n, sp_bits, pos3_bits, pos_div, key_a, idx_a);
        cub::DoubleBuffer<uint64_t> k(key_a, key_b);
        cub::DoubleBuffer<uint32_t> v(idx_a, idx_b);
        MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, 0, seq_bits + sp_bits + 3 + pos3_bits, st));
        if (seg_begin && seg_end) {
            MBL_CUDA(cudaMemsetAsync(seg_begin, 0, 8 * (size_t)n_reads, st));
            MBL_CUDA(cudaMemsetAsync(seg_end, 0, 8 * (size_t)n_reads, st));
        }
        match_gather_fix_kernel<<<blocks, 256, 0, st>>>(in, v.Current(), k.Current(), n, sp_bits + 3 + pos3_bits, n_reads, out,
                                                        seg_begin && seg_end ? seg_begin : nullptr, seg_end);
        return seg_begin && seg_end;
    }
    match_lowkey_kernel<<<blocks, 256, 0, st>>>(in, n, pos_bits, key_a, idx_a);
    cub::DoubleBuffer<uint64_t> k(key_a, key_b);
    cub::DoubleBuffer<uint32_t> v(idx_a, idx_b);
    MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, 0, 3 + pos_bits + 27, st));
    uint32_t* perm1 = v.Current();
    uint64_t* kcur = k.Current();
    uint64_t* kalt = k.Alternate();
    match_highkey_kernel<<<blocks, 256, 0, st>>>(in, perm1, n, sp_bits, kalt);
    // second pass: keys live in kalt, permutation in perm1
    cub::DoubleBuffer<uint64_t> k2(kalt, kcur);
    cub::DoubleBuffer<uint32_t> v2(perm1, v.Alternate());
    MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k2, v2, (long long)n, 0, seq_bits + sp_bits, st));
    match_gather_kernel<<<blocks, 256, 0, st>>>(in, v2.Current(), n, out);
    return false;
}

void launch_segments(const mbl_match_rec* sorted, size_t n, uint32_t n_reads, uint64_t* seg_begin, uint64_t* seg_end, cudaStream_t st) {
    MBL_CUDA(cudaMemsetAsync(seg_begin, 0, 8 * (size_t)n_reads, st));
    MBL_CUDA(cudaMemsetAsync(seg_end, 0, 8 * (size_t)n_reads, st));
    if (!n) return;
    segment_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sorted, n, n_reads, seg_begin, seg_end);
}

}  // namespace mbl
