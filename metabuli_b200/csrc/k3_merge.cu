// K3 — delta-decode + merge of the sorted query metamers against the differential index
// (reference rows A5-A8: KmerMatcher.cpp:123-481 matchKmers, :1117-1146 compareDna,
// KmerMatcher.h:282-297 getNextTargetKmer, :348-416 Hamming helpers).
//
// Reference: each OpenMP thread seeks to one of 4096 checkpoints and walks the stream serially,
// collecting for every query the target k-mers with the same 40-bit amino-acid part, then keeps the
// candidates whose codon-level Hamming sum is <= min(2*min, 7).
//
// B200: the index lives in HBM behind a tile directory (k3_index.cu).  A persistent CTA pulls work
// items (tile, query slice):
//   1. stage   two 1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) bring the tile's fragments
//              and taxids into shared memory;
//   2. decode  eight warps decode the tile's cells independently from their checkpoints
//              (delta_decode.cuh) into a sorted value array; every k-mer that starts an amino-acid group is
//              entered into a tagged bucket table (bucket = monotone hash of the 40-bit amino-acid part, entry =
//              {group start, 18-bit tag, collision flag}) so a query misses in ~12 instructions and one LDS;
//   3. match   a warp looks up 32 queries per iteration and queues the hits; 32 queued hits are processed one
//              per lane: a scan of the group finds its size and whether the query's exact value is present
//              (Hamming sum 0 <=> identical DNA part, so the survivors are exactly the equal values and no
//              table lookups are needed); only the remaining hits run the min / select Hamming passes over a
//              8 KiB two-codon table.  Output slots come from one atomicAdd per 32 hits.
// HBM traffic per launch = index once + 8 B per query (+ 8 B qinfo per matching query) + 24 B per match.
#include "delta_decode.cuh"
#include "kernels.cuh"

namespace mbl {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kQueue = 64;              // per-warp hit queue (power of two, >= 63)
constexpr uint32_t kLaneMaxCand = 48;        // groups larger than this are worked on by the whole warp
constexpr uint64_t kNone = ~0ull;
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kEmpty = 0xffffffffu;     // bucket table: no group hashes here

// hash table entry = group start (13 bits) << 19 | tag (19 bits).  Bucket and tag come from disjoint bits of a
// multiplicative hash of the 40-bit amino-acid part; a tag match is verified against the value array.
__device__ __forceinline__ uint32_t aa_hash(uint64_t aa40) { return (uint32_t)((aa40 * 0x9E3779B97F4A7C15ull) >> 32); }

// dynamic shared memory layout (sizes depend on the tile geometry chosen at load time)
struct SmemLayout {
    uint32_t off_ham, off_queue, off_frag0, off_frag1, off_vals, off_tab, total;
};
__host__ __device__ inline SmemLayout smem_layout(uint32_t max_u16, uint32_t max_kmers, uint32_t n_buckets) {
    SmemLayout l;
    uint32_t o = 32;                                   // two mbarriers + two item slots
    l.off_ham = o;    o += 8192;                       // two-codon table: sum | plain nibble | reversed nibble (u16)
    l.off_queue = o;  o += kWarps * kQueue * 12;       // per-warp hit queues {group start, query dna, query offset}
    l.off_frag0 = o;  o += (max_u16 + 16) * 2;         // fragment tile, double buffered: the next item's tile is
    l.off_frag1 = o;  o += (max_u16 + 16) * 2;         // in flight (TMA) while the current one is being matched
    l.off_vals = o;   o += max_kmers * 8;
    l.off_tab = o;    o += n_buckets * 4;              // open-addressing hash table of amino-acid group starts
    l.total = (o + 15) & ~15u;
    return l;
}

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

// codon-level Hamming distances of two 24-bit DNA parts: four lookups in the two-codon table, whose
// entries hold {sum:4, plain nibble:4, reversed nibble:4}; the four entries are kept so the per-codon fields
// of a surviving candidate cost no further lookups
struct HamQuad { uint32_t e0, e1, e2, e3; };
__device__ __forceinline__ HamQuad ham_lookup(const uint16_t* ham, uint32_t q, uint32_t t) {
    HamQuad h;
    h.e0 = ham[((q & 63u) << 6) | (t & 63u)];
    h.e1 = ham[(((q >> 6) & 63u) << 6) | ((t >> 6) & 63u)];
    h.e2 = ham[(((q >> 12) & 63u) << 6) | ((t >> 12) & 63u)];
    h.e3 = ham[((q >> 18) << 6) | (t >> 18)];
    return h;
}
__device__ __forceinline__ uint32_t ham_sum(const HamQuad& h) { return (h.e0 & 15u) + (h.e1 & 15u) + (h.e2 & 15u) + (h.e3 & 15u); }
// per-codon 2-bit fields (KmerMatcher.h:386-416): codon at value bits 3i goes to field bits 2i ("plain",
// getHammings) or 2(7-i) (getHammings_reverse); HAMMING_LUT7 rows 4-5 x columns 6-7 hold 1 instead of 0 (Q3)
__device__ __forceinline__ uint32_t ham_fields(const HamQuad& h, uint32_t q, uint32_t t, bool plain) {
    uint32_t f = plain ? (((h.e0 >> 4) & 15u) | (((h.e1 >> 4) & 15u) << 4) | (((h.e2 >> 4) & 15u) << 8) | (((h.e3 >> 4) & 15u) << 12))
                       : (((h.e3 >> 8) & 15u) | (((h.e2 >> 8) & 15u) << 4) | (((h.e1 >> 8) & 15u) << 8) | (((h.e0 >> 8) & 15u) << 12));
    const uint32_t qc = plain ? (q >> 21) & 7u : q & 7u, tc = plain ? (t >> 21) & 7u : t & 7u;
    if ((qc & 6u) == 4u && tc >= 6u) f |= 0x4000u;
    return f;
}

__device__ __forceinline__ uint64_t load_qinfo(const MergeArgs& a, uint64_t sorted_pos) {
    return a.q_info[a.q_idx ? (uint64_t)a.q_idx[sorted_pos] : sorted_pos];
}

// one 24-byte Match record (Match.h:9-26 without the vptr); Q2: taxid 0 / unmapped species raise the error flag
__device__ __forceinline__ void emit_match(const MergeArgs& a, uint64_t slot, uint64_t qinfo, int32_t raw_taxid, uint32_t td,
                                           uint32_t field, uint32_t sum) {
    const int32_t taxid = (int32_t)((uint32_t)raw_taxid & a.info_mask);
    const int32_t species = (taxid > 0 && taxid <= a.max_taxid) ? a.taxid2species[taxid] : 0;
    if (taxid == 0 || species == 0) atomicOr(a.error_flag, 1u);
    if (slot < a.out_cap) {
        uint64_t* w = reinterpret_cast<uint64_t*>(a.out + slot);
        w[0] = qinfo;
        w[1] = (uint64_t)(uint32_t)taxid | ((uint64_t)(uint32_t)species << 32);
        w[2] = (uint64_t)td | ((uint64_t)(field & 0xFFFFu) << 32) | ((uint64_t)(sum & 0xFFu) << 48);
    }
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
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout L = smem_layout(a.max_u16, a.max_kmers, a.n_buckets);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem);          // [2]
    unsigned int* s_item = reinterpret_cast<unsigned int*>(smem + 16);                 // [2]
    uint16_t* s_ham = reinterpret_cast<uint16_t*>(smem + L.off_ham);
    uint32_t* s_queue = reinterpret_cast<uint32_t*>(smem + L.off_queue);
    uint16_t* s_frag0 = reinterpret_cast<uint16_t*>(smem + L.off_frag0);
    const uint32_t frag_stride = (L.off_frag1 - L.off_frag0) / 2;     // in u16
    uint64_t* s_vals = reinterpret_cast<uint64_t*>(smem + L.off_vals);
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + L.off_tab);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4096; i += kThreads) s_ham[i] = a.ham_pair[i];
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); }
    __syncthreads();
    unsigned parity_bits = 0u;                        // bit b = phase parity of mbarrier b
    const uint32_t n_items = a.item_off[a.n_tiles];
    const bool fmt2 = a.kmer_format == 2;
    const uint32_t tab_mask = a.n_buckets - 1;
    const int hash_shift = 32 - (31 - __clz(a.n_buckets));
    uint32_t* my_queue = s_queue + warp * kQueue * 3;
    unsigned long long my_matches = 0;

    // thread 0: start the TMA copy of an item's fragment tile into buffer `buf`
    auto stage = [&](uint32_t item, int buf) {
        if (item >= n_items) return;
        const Tile t = a.tiles[a.items[item].tile];
        if (t.jumbo_off != kNone) return;
        const uint64_t d0 = t.diff_begin, d1 = t.diff_begin + t.n_u16;
        const uint64_t a0 = d0 & ~7ull, a1 = (d1 + 7ull) & ~7ull;
        fence_proxy_async();
        mbar_expect_tx(mbar + buf, (unsigned)((a1 - a0) * 2));
        tma_load_1d(s_frag0 + (uint32_t)buf * frag_stride, a.diff + a0, (unsigned)((a1 - a0) * 2), mbar + buf);
    };

    if (tid == 0) { s_item[0] = atomicAdd(a.item_cursor, 1u); stage(s_item[0], 0); }
    __syncthreads();
    uint32_t item = s_item[0];
    int buf = 0;

    while (item < n_items) {
        const MergeItem it = a.items[item];
        const Tile tl = a.tiles[it.tile];
        const uint32_t nk = tl.n_kmers;
        const bool jumbo = tl.jumbo_off != kNone;
        const uint64_t* vals;
        const int32_t* infos = a.info + tl.info_begin;
        // claim the next item and clear the hash table (the previous item's lookups ended at the barrier below)
        if (tid == 0) s_item[buf ^ 1] = atomicAdd(a.item_cursor, 1u);
        for (uint32_t x = tid; x < a.n_buckets; x += kThreads) s_tab[x] = kEmpty;
        __syncthreads();
        const uint32_t next_item = s_item[buf ^ 1];
        if (tid == 0) stage(next_item, buf ^ 1);             // its tile streams in while this item is decoded and matched
        if (!jumbo) {
            // -- 1. the tile's fragments were requested one item ago
            const uint64_t d0 = tl.diff_begin, d1 = tl.diff_begin + tl.n_u16;
            const uint64_t a0 = d0 & ~7ull;
            mbar_wait(mbar + buf, (parity_bits >> buf) & 1u);
            parity_bits ^= 1u << buf;
            // -- 2. decode: one warp per checkpoint cell; group starts go into the hash table
            const uint16_t* frag = s_frag0 + (uint32_t)buf * frag_stride;
            const uint64_t c0 = d0 / kCellU16, c1 = (d1 + kCellU16 - 1) / kCellU16;
            for (uint64_t c = c0 + warp; c < c1; c += kWarps) {
                const uint64_t s_abs = max(c * (uint64_t)kCellU16, d0), e_abs = min((c + 1) * (uint64_t)kCellU16, d1);
                uint64_t v, k;
                if (s_abs == d0) { v = tl.base_value; k = tl.info_begin; }
                else { v = a.cell_v[c]; k = a.cell_k[c]; }
                const uint64_t kb = tl.info_begin;
                warp_decode(frag, (long long)(d0 - a0), (long long)(s_abs - a0), (long long)(e_abs - a0), v, k,
                            [&](uint64_t kk, uint64_t val, uint64_t delta, long long) {
                                const uint64_t rel = kk - kb;
                                if (rel >= nk) return;
                                s_vals[rel] = val;
                                const uint64_t aa = val >> 24;
                                if (rel == 0 || ((val - delta) >> 24) != aa) {
                                    const uint32_t h = aa_hash(aa);
                                    const uint32_t entry = ((uint32_t)rel << 19) | (h & 0x7FFFFu);
                                    uint32_t slot = h >> hash_shift;
                                    while (atomicCAS(&s_tab[slot], kEmpty, entry) != kEmpty) slot = (slot + 1) & tab_mask;
                                }
                            });
            }
            __syncthreads();
            vals = s_vals;
        } else {
            vals = a.jumbo_vals + tl.jumbo_off;
        }

        // -- 3. stream the query slice.  A warp looks up 32 queries per iteration and appends the ones that found
        //       their amino-acid group ("hits") to its private queue; whenever 32 hits are queued they are processed
        //       one hit per lane, so every lane of the candidate loops has work.
        uint32_t q_head = 0, q_count = 0;
        auto process_hits = [&](uint32_t m) {
            const bool valid = (uint32_t)lane < m;
            const uint32_t* rec = my_queue + 3u * ((q_head + lane) & (kQueue - 1));
            const uint32_t g0 = valid ? rec[0] : 0u;                                // group start inside the tile
            const uint32_t qd = valid ? rec[1] : 0u;                                // query DNA part
            const uint32_t qoff = valid ? rec[2] : 0u;                              // query index relative to the item
            q_head = (q_head + m) & (kQueue - 1);
            q_count -= m;
            const uint64_t aa = valid ? vals[g0] >> 24 : 0ull;
            // pass A: size of the group (capped) and the run of candidates equal to the query value.  Candidates are
            // sorted by value, Hamming sum 0 <=> identical DNA part (the distance table is 0 only on its diagonal),
            // so when such a run exists min = 0, maxHamming = 0 and the survivors are exactly that run.
            uint32_t n = 0, ex0 = 0, exn = 0;
            bool open_end = valid;
            for (uint32_t c = 0; c <= kLaneMaxCand; ++c) {
                if (open_end) {
                    const uint32_t j = g0 + c;
                    const uint64_t v = j < nk ? vals[j] : ~0ull;
                    if ((v >> 24) != aa) { open_end = false; }
                    else {
                        n = c + 1;
                        if (((uint32_t)v & 0xFFFFFFu) == qd) { if (!exn) ex0 = j; ++exn; }
                    }
                }
                if (!__any_sync(kFull, open_end)) break;
            }
            // (a) oversized groups (still open after the cap): the whole warp works on one hit at a time
            uint32_t big = __ballot_sync(kFull, open_end);
            while (big) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                const uint32_t bg0 = __shfl_sync(kFull, g0, src);
                const uint32_t bqd = __shfl_sync(kFull, qd, src), bqoff = __shfl_sync(kFull, qoff, src);
                const uint64_t baa = __shfl_sync(kFull, aa, src);
                uint32_t bn = 0;                                  // group size, found cooperatively
                for (uint32_t cb = 0;; cb += 32) {
                    const uint32_t j = bg0 + cb + lane;
                    const bool in = j < nk && (vals[j] >> 24) == baa;
                    const uint32_t bal = __ballot_sync(kFull, in);
                    bn += __popc(bal);
                    if (bal != kFull) break;
                }
                uint32_t mn = 255u;
                for (uint32_t c = lane; c < bn; c += 32) mn = min(mn, ham_sum(ham_lookup(s_ham, bqd, (uint32_t)vals[bg0 + c] & 0xFFFFFFu)));
                mn = __reduce_min_sync(kFull, mn);
                const uint32_t maxH = min(mn * 2u, 7u);
                const uint64_t qinfo = load_qinfo(a, it.q_begin + bqoff);
                const bool plain = !((qi_frame(qinfo) < 3) ^ fmt2);
                for (uint32_t cb = 0; cb < bn; cb += 32) {
                    const uint32_t c = cb + lane;
                    const uint32_t td = c < bn ? (uint32_t)vals[bg0 + c] & 0xFFFFFFu : 0u;
                    const HamQuad hq = ham_lookup(s_ham, bqd, td);
                    const uint32_t sum = ham_sum(hq);
                    const bool sel = c < bn && sum <= maxH;
                    const uint32_t bal = __ballot_sync(kFull, sel);
                    if (!bal) continue;
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(a.out_count, (unsigned long long)__popc(bal));
                    base = __shfl_sync(kFull, base, 0);
                    if (sel) emit_match(a, base + __popc(bal & ((1u << lane) - 1)), qinfo, infos[bg0 + c], td, ham_fields(hq, bqd, td, plain), sum);
                    my_matches += __popc(bal);
                }
                if (lane == src) { n = 0; exn = 0; }
            }
            // (b) hits without an exact run: minimum Hamming sum, then the survivors (KmerMatcher.cpp:1117-1146)
            const uint32_t ne = exn ? 0u : n;                     // candidates this lane still has to evaluate
            uint32_t nmax = ne;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, o));
            uint32_t mn = 255u;
            uint64_t nib = 0;                                     // sums of the first 16 candidates, 4 bits each
            for (uint32_t c = 0; c < nmax; ++c) {
                if (c < ne) {
                    const uint32_t sm = ham_sum(ham_lookup(s_ham, qd, (uint32_t)vals[g0 + c] & 0xFFFFFFu));
                    mn = min(mn, sm);
                    if (c < 16) nib |= (uint64_t)min(sm, 15u) << (4 * c);
                }
            }
            const uint32_t maxH = min(mn * 2u, 7u);                                     // KmerMatcher.cpp:1136
            uint32_t cnt = exn;
            for (uint32_t c = 0; c < nmax; ++c) {
                if (c < ne) {
                    const uint32_t sm = c < 16 ? (uint32_t)(nib >> (4 * c)) & 15u
                                               : ham_sum(ham_lookup(s_ham, qd, (uint32_t)vals[g0 + c] & 0xFFFFFFu));
                    cnt += sm <= maxH;
                }
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            if (total == 0) return;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.out_count, (unsigned long long)total);
            base = __shfl_sync(kFull, base, 0) + (incl - cnt);
            my_matches += total;
            uint64_t qinfo = 0;
            bool plain = true;
            if (cnt) { qinfo = load_qinfo(a, it.q_begin + qoff); plain = !((qi_frame(qinfo) < 3) ^ fmt2); }   // KmerMatcher.cpp:1140
            // exact runs: Hamming 0, all per-codon fields 0
            uint32_t xmax = exn;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) xmax = max(xmax, __shfl_xor_sync(kFull, xmax, o));
            for (uint32_t c = 0; c < xmax; ++c)
                if (c < exn) { emit_match(a, base, qinfo, infos[ex0 + c], qd, 0u, 0u); ++base; }
            for (uint32_t c = 0; c < nmax; ++c) {
                if (c < ne && cnt) {
                    const uint32_t td = (uint32_t)vals[g0 + c] & 0xFFFFFFu;
                    const HamQuad hq = ham_lookup(s_ham, qd, td);
                    const uint32_t sum = ham_sum(hq);
                    if (sum <= maxH) { emit_match(a, base, qinfo, infos[g0 + c], td, ham_fields(hq, qd, td, plain), sum); ++base; }
                }
            }
        };

        for (uint64_t qb = it.q_begin + (uint64_t)warp * 32; qb < it.q_end; qb += kThreads) {
            const uint64_t qi = qb + lane;
            const bool active = qi < it.q_end;
            const uint64_t qv = active ? ld_stream_u64(a.q_value + qi) : kBlank;
            const uint64_t q40 = qv >> 24;
            uint32_t g0 = 0;
            bool hit = false;
            if (!jumbo) {
                if (active) {
                    const uint32_t h = aa_hash(q40);
                    const uint32_t tag = h & 0x7FFFFu;
                    uint32_t slot = h >> hash_shift;
                    for (uint32_t e = s_tab[slot]; e != kEmpty; e = s_tab[slot = (slot + 1) & tab_mask]) {
                        if ((e & 0x7FFFFu) == tag && (vals[e >> 19] >> 24) == q40) { g0 = e >> 19; hit = true; break; }
                    }
                }
            } else if (active) {
                uint32_t lo = 0, hi = nk;
                while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if ((vals[mid] >> 24) < q40) lo = mid + 1; else hi = mid; }
                if (lo < nk && (vals[lo] >> 24) == q40) { g0 = lo; hit = true; }
            }
            const uint32_t bal = __ballot_sync(kFull, hit);
            if (bal) {
                if (hit) {
                    uint32_t* rec = my_queue + 3u * ((q_head + q_count + __popc(bal & ((1u << lane) - 1))) & (kQueue - 1));
                    rec[0] = g0; rec[1] = (uint32_t)qv & 0xFFFFFFu; rec[2] = (uint32_t)(qi - it.q_begin);
                }
                q_count += __popc(bal);
                __syncwarp();
                if (q_count >= 32) { process_hits(32); __syncwarp(); }
            }
        }
        if (q_count) { process_hits(q_count); __syncwarp(); }
        __syncthreads();
        item = next_item;
        buf ^= 1;
    }
    if (lane == 0 && my_matches) atomicAdd(a.out_count + 1, my_matches);
}

size_t merge_smem_bytes(uint32_t max_u16, uint32_t max_kmers, uint32_t n_buckets) { return smem_layout(max_u16, max_kmers, n_buckets).total; }

void launch_merge_plan(const MergeArgs& a, cudaStream_t st) {
    const unsigned blocks = (unsigned)((a.n_tiles + 1 + 255) / 256);
    merge_partition_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.n_tiles, a.q_value, a.n_query, a.q_lo);
    merge_item_count_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.n_tiles, a.q_lo, a.item_cnt);
    exclusive_sum_u32(a.scan_tmp, a.scan_tmp_bytes, a.item_cnt, a.item_off, a.n_tiles + 1, st);
    merge_item_fill_kernel<<<blocks, 256, 0, st>>>(a.n_tiles, a.q_lo, a.item_cnt, a.item_off, a.items, a.items_cap);
}

void launch_merge(const MergeArgs& a, int sm_count, cudaStream_t st) {
    static size_t attr_smem = 0;
    const size_t smem = smem_layout(a.max_u16, a.max_kmers, a.n_buckets).total;
    if (attr_smem != smem) {
        MBL_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    int per_sm = 0;
    MBL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_kernel, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    merge_kernel<<<(unsigned)(sm_count * per_sm), kThreads, smem, st>>>(a);
}

}  // namespace mbl
