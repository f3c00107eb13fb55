// K3 — delta-decode + merge of the sorted query metamers against the differential index
// (reference rows A5-A8: KmerMatcher.cpp:123-481 matchKmers, :1117-1146 compareDna,
// KmerMatcher.h:282-297 getNextTargetKmer, :348-416 Hamming helpers).
//
// Reference: each OpenMP thread seeks to one of 4096 checkpoints and walks the stream serially,
// collecting for every query the target k-mers with the same 40-bit amino-acid part, then keeps the
// candidates whose codon-level Hamming sum is <= min(2*min, 7).
//
// B200: the index lives in HBM behind a tile directory (k3_index.cu).  A persistent CTA pulls work
// items (tile, query slice); the tile's fragments and taxids are staged into shared memory with two
// 1-D TMA bulk copies (cp.async.bulk + mbarrier), eight warps decode the tile's cells independently
// from their checkpoints (delta_decode.cuh) into a sorted value array in shared memory, and the CTA then
// streams its slice of the sorted queries through: binary search of the amino-acid group, Hamming
// filter via a 4096-entry two-codon table, Match records written through warp-private output chunks.
// HBM traffic per launch = index once + 8 B per query (+ 8 B qinfo per matching query) + 24 B per match.
#include "delta_decode.cuh"
#include "kernels.cuh"

namespace mbl {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kOutChunk = 256;          // match slots a warp reserves at a time
constexpr uint64_t kNone = ~0ull;

struct __align__(16) MergeSmem {
    unsigned long long mbar;
    unsigned int item;
    unsigned int pad;
    uint16_t ham[4096];
    uint16_t frag[kTileMaxU16 + 16];
    int32_t info[kTileMaxKmers + 8];
    uint64_t vals[kTileMaxKmers];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

struct HamOut { uint32_t sum, plain, rev; };
// codon-level Hamming distances of two 24-bit DNA parts via the two-codon table
__device__ __forceinline__ HamOut hamming(const uint16_t* ham, uint32_t q, uint32_t t) {
    uint32_t e0 = ham[((q & 63u) << 6) | (t & 63u)];
    uint32_t e1 = ham[(((q >> 6) & 63u) << 6) | ((t >> 6) & 63u)];
    uint32_t e2 = ham[(((q >> 12) & 63u) << 6) | ((t >> 12) & 63u)];
    uint32_t e3 = ham[(((q >> 18) & 63u) << 6) | ((t >> 18) & 63u)];
    HamOut o;
    o.sum = (e0 & 15u) + (e1 & 15u) + (e2 & 15u) + (e3 & 15u);
    o.plain = ((e0 >> 4) & 15u) | (((e1 >> 4) & 15u) << 4) | (((e2 >> 4) & 15u) << 8) | (((e3 >> 4) & 15u) << 12);
    o.rev = ((e3 >> 8) & 15u) | (((e2 >> 8) & 15u) << 4) | (((e1 >> 8) & 15u) << 8) | (((e0 >> 8) & 15u) << 12);
    return o;
}
__device__ __forceinline__ uint32_t ham_sum_only(const uint16_t* ham, uint32_t q, uint32_t t) {
    return (ham[((q & 63u) << 6) | (t & 63u)] & 15u) + (ham[(((q >> 6) & 63u) << 6) | ((t >> 6) & 63u)] & 15u) +
           (ham[(((q >> 12) & 63u) << 6) | ((t >> 12) & 63u)] & 15u) + (ham[(((q >> 18) & 63u) << 6) | ((t >> 18) & 63u)] & 15u);
}

}  // namespace

// ---- work planning ---------------------------------------------------------------------------------
// q_lo[t] = first query whose amino-acid part is >= the tile's first amino-acid part
__global__ void merge_partition_kernel(const Tile* __restrict__ tiles, uint64_t n_tiles, const uint64_t* __restrict__ q_value,
                                       uint64_t n_query, uint64_t* __restrict__ q_lo) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) { q_lo[t] = n_query; return; }
    const uint64_t key = tiles[t].first_aa;
    uint64_t lo = 0, hi = n_query;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (aa_part(q_value[mid]) < key) lo = mid + 1; else hi = mid;
    }
    q_lo[t] = lo;
}
__global__ void merge_item_count_kernel(const Tile* __restrict__ tiles, uint64_t n_tiles, const uint64_t* __restrict__ q_lo,
                                        uint32_t* __restrict__ item_cnt) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    uint32_t c = 0;
    if (t < n_tiles && tiles[t].n_kmers > 0) {
        uint64_t nq = q_lo[t + 1] - q_lo[t];
        c = (uint32_t)((nq + kItemQueries - 1) / kItemQueries);
    }
    item_cnt[t] = c;
}
__global__ void merge_item_fill_kernel(uint64_t n_tiles, const uint64_t* __restrict__ q_lo, const uint32_t* __restrict__ item_cnt,
                                       const uint32_t* __restrict__ item_off, MergeItem* __restrict__ items, uint64_t items_cap) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint32_t c = item_cnt[t];
    const uint64_t b = q_lo[t], e = q_lo[t + 1];
    for (uint32_t i = 0; i < c; ++i) {
        uint64_t slot = (uint64_t)item_off[t] + i;
        if (slot >= items_cap) return;
        MergeItem it;
        it.tile = (uint32_t)t; it.pad = 0;
        it.q_begin = b + (uint64_t)i * kItemQueries;
        it.q_end = min(e, it.q_begin + kItemQueries);
        items[slot] = it;
    }
}

// ---- the merge kernel ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 3)
merge_kernel(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MergeSmem& sm = *reinterpret_cast<MergeSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 4096; i += kThreads) sm.ham[i] = a.ham_pair[i];
    if (tid == 0) mbar_init(&sm.mbar, 1);
    __syncthreads();
    unsigned parity = 0;
    const uint32_t n_items = a.item_off[a.n_tiles];
    const bool fmt2 = a.kmer_format == 2;

    // warp-private output chunk
    uint64_t chunk_base = 0;
    uint32_t chunk_used = kOutChunk;         // forces a reservation on first use
    unsigned long long my_matches = 0;

    while (true) {
        if (tid == 0) sm.item = atomicAdd(a.item_cursor, 1u);
        __syncthreads();
        const uint32_t item = sm.item;
        if (item >= n_items) break;
        const MergeItem it = a.items[item];
        const Tile tl = a.tiles[it.tile];
        const uint32_t nk = tl.n_kmers;
        const uint64_t* vals;
        const int32_t* infos;
        if (tl.jumbo_off == kNone) {
            // -- stage the tile: fragments + taxids, two bulk copies on one mbarrier
            const uint64_t d0 = tl.diff_begin, d1 = tl.diff_begin + tl.n_u16;
            const uint64_t a0 = d0 & ~7ull, a1 = (d1 + 7ull) & ~7ull;
            const uint64_t i0 = tl.info_begin & ~3ull, i1 = (tl.info_begin + nk + 3ull) & ~3ull;
            if (tid == 0) {
                fence_proxy_async();
                const unsigned fb = (unsigned)((a1 - a0) * 2), ib = (unsigned)((i1 - i0) * 4);
                mbar_expect_tx(&sm.mbar, fb + ib);
                tma_load_1d(sm.frag, a.diff + a0, fb, &sm.mbar);
                if (ib) tma_load_1d(sm.info, a.info + i0, ib, &sm.mbar);
            }
            mbar_wait(&sm.mbar, parity);
            parity ^= 1u;
            // -- decode: one warp per checkpoint cell
            const uint64_t c0 = d0 / kCellU16, c1 = (d1 + kCellU16 - 1) / kCellU16;
            for (uint64_t c = c0 + warp; c < c1; c += kWarps) {
                const uint64_t s_abs = max(c * (uint64_t)kCellU16, d0), e_abs = min((c + 1) * (uint64_t)kCellU16, d1);
                uint64_t v, k;
                if (s_abs == d0) { v = tl.base_value; k = tl.info_begin; }
                else { v = a.cell_v[c]; k = a.cell_k[c]; }
                uint64_t* out_vals = sm.vals;
                const uint64_t kb = tl.info_begin;
                warp_decode(sm.frag, (long long)(d0 - a0), (long long)(s_abs - a0), (long long)(e_abs - a0), v, k,
                            [&](uint64_t kk, uint64_t val, uint64_t, long long) {
                                uint64_t rel = kk - kb;
                                if (rel < nk) out_vals[rel] = val;
                            });
            }
            __syncthreads();
            vals = sm.vals;
            infos = sm.info + (tl.info_begin - i0);
        } else {
            vals = a.jumbo_vals + tl.jumbo_off;
            infos = a.info + tl.info_begin;
        }

        // -- stream the query slice
        for (uint64_t qb = it.q_begin + (uint64_t)warp * 32; qb < it.q_end; qb += kThreads) {
            const uint64_t qi = qb + lane;
            const bool active = qi < it.q_end;
            const uint64_t qv = active ? ld_stream_u64(a.q_value + qi) : kBlank;
            const uint64_t qaa = aa_part(qv);
            uint32_t lo = 0, hi = nk;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (vals[mid] < qaa) lo = mid + 1; else hi = mid;
            }
            uint32_t g0 = lo, g1 = lo, cnt = 0, maxH = 0;
            const uint32_t qd = (uint32_t)(qv & kDnaMask);
            if (active && g0 < nk && aa_part(vals[g0]) == qaa) {
                uint32_t minH = 255;
                for (; g1 < nk && aa_part(vals[g1]) == qaa; ++g1)
                    minH = min(minH, ham_sum_only(sm.ham, qd, (uint32_t)(vals[g1] & kDnaMask)));
                maxH = min(minH * 2u, 7u);                                   // KmerMatcher.cpp:1136
                for (uint32_t j = g0; j < g1; ++j)
                    cnt += ham_sum_only(sm.ham, qd, (uint32_t)(vals[j] & kDnaMask)) <= maxH;
            }
            // warp-aggregated slot assignment inside the warp's private chunk
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;
            const uint32_t rem = kOutChunk - chunk_used;
            uint64_t new_base = 0;
            if (total > rem) {
                const uint32_t need = max(kOutChunk, total - rem);
                if (lane == 0) new_base = atomicAdd(a.out_count, (unsigned long long)need);
                new_base = __shfl_sync(0xffffffffu, new_base, 0);
            }
            if (cnt) {
                const uint64_t qinfo = a.q_info[qi];
                const uint32_t frame = qi_frame(qinfo);
                const bool plain = !((frame < 3) ^ fmt2);                    // KmerMatcher.cpp:1140
                uint32_t w = incl - cnt;
                for (uint32_t j = g0; j < g1; ++j) {
                    const uint64_t tv = vals[j];
                    const uint32_t td = (uint32_t)(tv & kDnaMask);
                    const HamOut h = hamming(sm.ham, qd, td);
                    if (h.sum > maxH) continue;
                    uint32_t field = plain ? h.plain : h.rev;
                    // HAMMING_LUT7 rows 4-5 x columns 6-7 hold 1 (Q3); it serves the codon at value bits 21..23
                    // in the plain orientation and the codon at bits 0..2 in the reversed one
                    const uint32_t qc = plain ? (qd >> 21) & 7u : qd & 7u, tc = plain ? (td >> 21) & 7u : td & 7u;
                    if ((qc & 6u) == 4u && tc >= 6u) field |= 0x4000u;
                    const int32_t taxid = (int32_t)((uint32_t)infos[j] & a.info_mask);
                    const int32_t species = (taxid > 0 && taxid <= a.max_taxid) ? a.taxid2species[taxid] : 0;
                    if (taxid == 0 || species == 0) atomicOr(a.error_flag, 1u);   // Q2
                    const uint64_t slot = (w < rem) ? chunk_base + chunk_used + w : new_base + (w - rem);
                    if (slot < a.out_cap) {
                        uint64_t* o = reinterpret_cast<uint64_t*>(a.out + slot);
                        o[0] = qinfo;
                        o[1] = (uint64_t)(uint32_t)taxid | ((uint64_t)(uint32_t)species << 32);
                        o[2] = (uint64_t)td | ((uint64_t)(field & 0xFFFFu) << 32) | ((uint64_t)(h.sum & 0xFFu) << 48);
                    }
                    ++w;
                }
            }
            my_matches += cnt;
            if (total > rem) {
                chunk_base = new_base;
                chunk_used = total - rem;
                // a reservation larger than one chunk is consumed entirely by this request
                if (total - rem > kOutChunk) chunk_used = kOutChunk;
            } else {
                chunk_used += total;
            }
        }
        __syncthreads();
    }
    // blank out the unused tail of the warp's last chunk (seqID 0 == not a match)
    if (chunk_used < kOutChunk) {
        for (uint32_t w = chunk_used + lane; w < kOutChunk; w += 32) {
            const uint64_t slot = chunk_base + w;
            if (slot < a.out_cap) {
                uint64_t* o = reinterpret_cast<uint64_t*>(a.out + slot);
                o[0] = 0; o[1] = 0; o[2] = 0;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_matches += __shfl_xor_sync(0xffffffffu, my_matches, o);
    if (lane == 0 && my_matches) atomicAdd(a.out_count + 1, my_matches);
}

size_t merge_smem_bytes() { return sizeof(MergeSmem); }

void launch_merge_plan(const MergeArgs& a, cudaStream_t st) {
    const unsigned blocks = (unsigned)((a.n_tiles + 1 + 255) / 256);
    merge_partition_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.n_tiles, a.q_value, a.n_query, a.q_lo);
    merge_item_count_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.n_tiles, a.q_lo, a.item_cnt);
    exclusive_sum_u32(a.scan_tmp, a.scan_tmp_bytes, a.item_cnt, a.item_off, a.n_tiles + 1, st);
    merge_item_fill_kernel<<<blocks, 256, 0, st>>>(a.n_tiles, a.q_lo, a.item_cnt, a.item_off, a.items, a.items_cap);
}

void launch_merge(const MergeArgs& a, int sm_count, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = sizeof(MergeSmem);
    if (!attr_set) {
        MBL_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int per_sm = 0;
    MBL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_kernel, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    merge_kernel<<<(unsigned)(sm_count * per_sm), kThreads, smem, st>>>(a);
}

}  // namespace mbl
