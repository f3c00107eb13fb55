// K3 — delta-decode + merge of the sorted query metamers against the differential index
// (reference rows A5-A8: KmerMatcher.cpp:123-481 matchKmers, :1117-1146 compareDna,
// KmerMatcher.h:282-297 getNextTargetKmer, :348-416 Hamming helpers).
//
// Reference: each OpenMP thread seeks to one of 4096 checkpoints and walks the stream serially,
// collecting for every query the target k-mers with the same 40-bit amino-acid part, then keeps the
// candidates whose codon-level Hamming sum is <= min(2*min, 7).
//
// B200: the index lives in HBM behind a tile directory (k3_index.cu).  A persistent CTA pulls work items (a tile and
// a slice of its queries; the item carries the tile geometry, fetched into shared memory with cp.async one item ahead):
//   1. stage   two 1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) bring the tile's fragments and taxids
//              into shared memory; the fragments of the next tile are requested as soon as this one is decoded;
//   2. decode  the warps that hold fragments decode the tile together (delta_decode.cuh: one 16-byte octet per thread,
//              shuffle scan of (count, sum), one barrier per sweep) into a value array; a second, k-mer-parallel pass
//              marks the first k-mer of every amino-acid group in a bitmap (one ballot per 32 k-mers) and enters it
//              into a bucketed hash table keyed by the 40-bit amino-acid part;
//   3. match   CTA-wide: probe (hit list + pair owners) -> pass 1 (Hamming sums, group minima) -> pass 2 (survivors,
//              ballot-compacted, staged per warp, written as contiguous 8-byte words); see merge_kernel_v2 below.
//              Output slots come from warp-private chunks of 1024 (one global atomic per chunk).
// HBM traffic per launch = index once + 16 B per query (value and qinfo, both sorted streams) + 24 B per match: ncu measures
// 1.01x these algorithmic bytes (profiles/r02_v3_merge_ncu.md).
#include "delta_decode.cuh"
#include "kernels.cuh"

namespace mbl {

namespace {

// CTA shapes: 512 threads x 2 CTAs per SM (default) or 256 threads x up to 4 CTAs per SM, 64 registers either way;
// MergeArgs::cta_threads selects one at launch
constexpr uint32_t kOutChunk = 1024;         // match slots a warp reserves from the global cursor at a time
constexpr uint64_t kNone = ~0ull;
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kEmpty = 0xffffffffu;     // hash table: free slot

// hash table of amino-acid group starts: buckets of two entries, entry = group start (13 bits) << 19 | tag (19 bits).
// Bucket and tag come from disjoint bits of a multiplicative hash of the 40-bit amino-acid part; a tag match is
// verified against the value array.
__device__ __forceinline__ uint32_t aa_hash(uint64_t aa40) { return (uint32_t)((aa40 * 0x9E3779B97F4A7C15ull) >> 32); }

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

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// Ampere-style asynchronous copies of single words: the gathers a hit needs (slot index, qinfo) land in shared memory
// without a register round trip, so nothing waits on them until the data is used
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

// Output slots: a single global cursor would serialise ~10^7 atomics per launch, so every warp reserves kOutChunk
// slots at a time and hands them out locally; a request that does not fit is split across the old and the new chunk.
// What is left of the last chunk at kernel end is filled with blank records (seqID 0), which sort first and are
// ignored downstream.
struct OutChunk { uint64_t base = 0; uint32_t used = kOutChunk; };
struct Reservation { uint64_t old_base, new_base; uint32_t old_used, rem; };
__device__ __forceinline__ Reservation reserve(OutChunk& c, uint32_t cnt, unsigned long long* cursor, int lane) {
    Reservation r;
    r.old_base = c.base; r.old_used = c.used; r.rem = kOutChunk - c.used; r.new_base = 0;
    if (cnt > r.rem) {
        const uint32_t need = max(kOutChunk, cnt - r.rem);
        unsigned long long nb = 0;
        if (lane == 0) nb = atomicAdd(cursor, (unsigned long long)need);
        r.new_base = __shfl_sync(kFull, nb, 0);
        c.base = r.new_base;
        c.used = (cnt - r.rem > kOutChunk) ? kOutChunk : cnt - r.rem;
    } else {
        c.used += cnt;
    }
    return r;
}
__device__ __forceinline__ uint64_t slot_of(const Reservation& r, uint32_t i) {
    return i < r.rem ? r.old_base + r.old_used + i : r.new_base + (i - r.rem);
}

}  // namespace

// ---- work planning ---------------------------------------------------------------------------------
// q_lo[t] = first query whose amino-acid part is >= the tile's first amino-acid part
// The queries are radix-sorted on value >> prefix_shift only (k2_sort.cu), which is all the merge needs: a tile takes
// every query whose prefix lies between the prefixes of its first and last k-mer and the per-tile hash table rejects
// the ones that belong to a neighbour (tiles are amino-acid-group aligned, so a query can hit in one tile only).
__global__ void merge_partition_kernel(const Tile* __restrict__ tiles, uint64_t n_tiles, const uint64_t* __restrict__ q_value,
                                       uint64_t n_query, int prefix_shift, uint64_t* __restrict__ q_lo) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint64_t key_lo = tiles[t].first_aa >> prefix_shift;          // first_aa = value with the DNA bits cleared
    const uint64_t key_hi = tiles[t].last_value >> prefix_shift;
    uint64_t lo = 0, hi = n_query;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((q_value[mid] >> prefix_shift) < key_lo) lo = mid + 1; else hi = mid;
    }
    q_lo[2 * t] = lo;
    hi = n_query;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((q_value[mid] >> prefix_shift) <= key_hi) lo = mid + 1; else hi = mid;
    }
    q_lo[2 * t + 1] = lo;
}
__global__ void merge_item_count_kernel(const Tile* __restrict__ tiles, uint64_t n_tiles, const uint64_t* __restrict__ q_lo,
                                        uint32_t* __restrict__ item_cnt) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    uint32_t c = 0;
    if (t < n_tiles && tiles[t].n_kmers > 0) {
        uint64_t nq = q_lo[2 * t + 1] - q_lo[2 * t];
        c = (uint32_t)((nq + kItemQueries - 1) / kItemQueries);
    }
    item_cnt[t] = c;
}
__global__ void merge_item_fill_kernel(const Tile* __restrict__ tiles, uint64_t n_tiles, const uint64_t* __restrict__ q_lo,
                                       const uint32_t* __restrict__ item_cnt, const uint32_t* __restrict__ item_off,
                                       MergeItem* __restrict__ items, uint64_t items_cap) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint32_t c = item_cnt[t];
    if (!c) return;
    const uint64_t b = q_lo[2 * t], e = q_lo[2 * t + 1];
    const Tile tl = tiles[t];
    MergeItem it;
    it.diff_begin = tl.diff_begin; it.info_begin = tl.info_begin; it.base_value = tl.base_value; it.jumbo_off = tl.jumbo_off;
    it.n_u16 = tl.n_u16; it.n_kmers = tl.n_kmers; it.tile = (uint32_t)t; it.pad = 0;
    for (uint32_t i = 0; i < c; ++i) {
        uint64_t slot = (uint64_t)item_off[t] + i;
        if (slot >= items_cap) return;
        it.q_begin = b + (uint64_t)i * kItemQueries;
        it.q_end = min(e, it.q_begin + kItemQueries);
        items[slot] = it;
    }
}

// ---- the merge kernel ------------------------------------------------------------------------------------------------------
// The match stage is balanced over the whole CTA.  (Round 1 kept hits in warp-private queues and found their pairs with
// shuffle searches: 17 % of the stall samples sat at the closing barrier of an item, waiting for the warp with the fullest
// queue, and 32 (query, candidate) pairs cost ~365 warp instructions in two sweeps; measured 48.9 ms against 38.0 ms for this
// kernel on the benchmark, profiles/r02_v2_bench_*.json.)
//   probe   a thread looks up one query per step; a hit appends {group start | length, pair offset, minimum | query DNA, qinfo}
//           to a CTA-wide hit list (one shared-memory atomic per warp and step), its qinfo word is loaded right there (it travels
//           through the K2 sort next to the value, so this is a coalesced stream and not a gather), and the hit's thread writes
//           its index into the owner slot of every (query, candidate) pair the hit expands to;
//   pass 1  pairs are spread over all threads: Hamming sum over the 8 KiB two-codon table, one byte per pair kept, atomicMin
//           into the hit's record;
//   pass 2  pairs again: keep those with sum <= min(2 * minimum, 7) (KmerMatcher.cpp:1136), ballot-compact the warp's
//           survivors into its staging buffer and write the rows out as contiguous 8-byte words.
// A round takes up to 2 * kThreads queries and 8 * kThreads pairs; a round with more pairs (a hot amino-acid group) walks the
// pair range in windows and recomputes the sums in pass 2.  Jumbo items (an amino-acid group larger than a shared-memory tile:
// values pre-decoded in HBM) run the same passes over HBM-resident values, their queries finding the group by binary search.
struct SmemLayout2 {
    uint32_t off_rec, off_scan, off_ham, off_hg, off_hl, off_hp, off_hm, off_hq, off_owner, off_psum, off_stage, off_bits, off_frag, off_info, off_vals,
        off_tab, total;
};
__host__ __device__ inline SmemLayout2 smem_layout2(uint32_t max_u16, uint32_t max_kmers, uint32_t n_buckets, uint32_t kThreads) {
    const uint32_t kWarps = kThreads / 32, kQR = 2 * kThreads, kPC = 8 * kThreads;
    SmemLayout2 l;
    uint32_t o = 64;                                   // two mbarriers, two item slots, hit / pair counters
    l.off_rec = o;    o += 2 * 64;                     // work item records: current / next
    l.off_scan = o;   o += 2 * 2 * kWarps * 8;         // block_decode cross-warp scan
    l.off_ham = o;    o += 8192;                       // two-codon table
    l.off_hq = o;     o += kQR * 8;                    // hit list: qinfo
    l.off_hg = o;     o += kQR * 4;                    //           group start
    l.off_hl = o;     o += kQR * 4;                    //           group length
    l.off_hp = o;     o += kQR * 4;                    //           first pair of the hit
    l.off_hm = o;     o += kQR * 4;                    //           minimum Hamming sum << 24 | query DNA part
    l.off_owner = o;  o += kPC * 2;                    // pair -> hit
    l.off_psum = o;   o += kPC;                        // pair -> Hamming sum
    o = (o + 15) & ~15u;
    l.off_stage = o;  o += kWarps * 32 * 24;           // per-warp staging of up to 32 Match rows
    l.off_bits = o;   o += ((max_kmers + 63) / 32) * 4;
    o = (o + 15) & ~15u;
    l.off_frag = o;   o += (max_u16 + 16) * 2;
    l.off_info = o;   o += (max_kmers + 8) * 4;
    o = (o + 15) & ~15u;
    l.off_vals = o;   o += max_kmers * 8;
    l.off_tab = o;    o += n_buckets * 4;
    l.total = (o + 15) & ~15u;
    return l;
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 4 : 2)      // 64 registers either way: 32 warps per SM when shared memory allows
merge_kernel_v2(MergeArgs a) {
    constexpr int kWarps = kThreads / 32;
    constexpr uint32_t kQR = 2 * kThreads;            // queries per round
    constexpr uint32_t kPC = 8 * kThreads;            // pairs per window
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout2 L = smem_layout2(a.max_u16, a.max_kmers, a.n_buckets, kThreads);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem);          // [0] fragments, [1] taxids
    unsigned int* s_item = reinterpret_cast<unsigned int*>(smem + 32);                 // [2] claimed item numbers
    unsigned int* s_nhit = reinterpret_cast<unsigned int*>(smem + 40);
    unsigned int* s_npair = reinterpret_cast<unsigned int*>(smem + 44);
    MergeItem* s_rec = reinterpret_cast<MergeItem*>(smem + L.off_rec);
    uint64_t* s_scan = reinterpret_cast<uint64_t*>(smem + L.off_scan);
    uint16_t* s_ham = reinterpret_cast<uint16_t*>(smem + L.off_ham);
    uint64_t* s_hq = reinterpret_cast<uint64_t*>(smem + L.off_hq);
    uint32_t* s_hg = reinterpret_cast<uint32_t*>(smem + L.off_hg);
    uint32_t* s_hl = reinterpret_cast<uint32_t*>(smem + L.off_hl);
    uint32_t* s_hp = reinterpret_cast<uint32_t*>(smem + L.off_hp);
    uint32_t* s_hm = reinterpret_cast<uint32_t*>(smem + L.off_hm);
    uint16_t* s_owner = reinterpret_cast<uint16_t*>(smem + L.off_owner);
    uint8_t* s_psum = smem + L.off_psum;
    uint64_t* s_stage = reinterpret_cast<uint64_t*>(smem + L.off_stage);
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem + L.off_bits);
    uint16_t* s_frag = reinterpret_cast<uint16_t*>(smem + L.off_frag);
    int32_t* s_info = reinterpret_cast<int32_t*>(smem + L.off_info);
    uint64_t* s_vals = reinterpret_cast<uint64_t*>(smem + L.off_vals);
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + L.off_tab);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = tid; i < 4096; i += kThreads) s_ham[i] = a.ham_pair[i];
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); *s_nhit = 0u; *s_npair = 0u; }
    unsigned parity_bits = 0u;
    const uint32_t n_items_all = a.item_off[a.n_tiles];
    const uint32_t n_items = (uint32_t)min((uint64_t)n_items_all, a.items_cap);
    if (n_items_all > n_items && blockIdx.x == 0 && tid == 0) atomicOr(a.error_flag, 2u);   // work list too short: host retries
    const bool fmt2 = a.kmer_format == 2;
    const bool via_idx = a.q_idx != nullptr;
    const uint32_t bucket_mask = a.n_buckets / 2 - 1;
    const int hash_shift = 32 - (31 - __clz(a.n_buckets / 2));
    uint64_t* my_stage = s_stage + warp * 96;
    unsigned long long my_matches = 0;
    OutChunk chunk;

    auto stage_frag = [&](const MergeItem& r) {
        if (r.jumbo_off != kNone) return;
        const uint64_t d0 = r.diff_begin, d1 = r.diff_begin + r.n_u16;
        const uint64_t a0 = d0 & ~7ull, a1 = (d1 + 7ull) & ~7ull;
        fence_proxy_async();
        mbar_expect_tx(mbar, (unsigned)((a1 - a0) * 2));
        tma_load_1d(s_frag, a.diff + a0, (unsigned)((a1 - a0) * 2), mbar);
    };
    // warp-collective: ballot-compact the selected lanes' Match rows (Match.h:9-26 without the vptr) into the warp's staging
    // buffer and write them as contiguous 8-byte words; Q2: taxid 0 / unmapped species raise the error flag
    auto emit = [&](bool sel, uint64_t qinfo, int32_t taxid_raw, uint32_t qd, uint32_t td, uint32_t sum) {
        const uint32_t bal = __ballot_sync(kFull, sel);
        if (!bal) return;
        const uint32_t cnt = __popc(bal);
        const Reservation rs = reserve(chunk, cnt, a.out_count, lane);
        if (sel) {
            const int32_t taxid = (int32_t)((uint32_t)taxid_raw & a.info_mask);
            const int32_t species = (taxid > 0 && taxid <= a.max_taxid) ? __ldg(a.taxid2species + taxid) : 0;
            uint32_t field = 0u;
            if (sum) field = ham_fields(ham_lookup(s_ham, qd, td), qd, td, !((qi_frame(qinfo) < 3) ^ fmt2));   // KmerMatcher.cpp:1140
            if (taxid == 0 || species == 0) atomicOr(a.error_flag, 1u);
            uint64_t* w = my_stage + 3u * (uint32_t)__popc(bal & lt);
            w[0] = qinfo;
            w[1] = (uint64_t)(uint32_t)taxid | ((uint64_t)(uint32_t)species << 32);
            w[2] = (uint64_t)td | ((uint64_t)(field & 0xFFFFu) << 32) | ((uint64_t)(sum & 0xFFu) << 48);
        }
        __syncwarp();
        if (cnt <= rs.rem) {                                                    // one contiguous run of slots (the usual case)
            const uint64_t s0 = rs.old_base + rs.old_used;
            if (s0 + cnt <= a.out_cap) {
                uint64_t* dst = reinterpret_cast<uint64_t*>(a.out + s0);
                const uint32_t nw3 = 3u * cnt;                                  // <= 96 words: three predicated copies
                if ((uint32_t)lane < nw3) dst[lane] = my_stage[lane];
                if ((uint32_t)lane + 32u < nw3) dst[lane + 32] = my_stage[lane + 32];
                if ((uint32_t)lane + 64u < nw3) dst[lane + 64] = my_stage[lane + 64];
            }
        } else {
            for (uint32_t w = lane; w < 3u * cnt; w += 32) {
                const uint32_t r = w / 3u;
                const uint64_t sl = slot_of(rs, r);
                if (sl < a.out_cap) reinterpret_cast<uint64_t*>(a.out + sl)[w - 3u * r] = my_stage[w];
            }
        }
        __syncwarp();
        my_matches += cnt;
    };

    if (tid == 0) {
        s_item[0] = atomicAdd(a.item_cursor, 1u);
        s_item[1] = atomicAdd(a.item_cursor, 1u);
        if (s_item[0] < n_items) { s_rec[0] = a.items[s_item[0]]; }
    }
    __syncthreads();
    uint32_t item = s_item[0];
    int slot = 0;
    if (tid == 0 && item < n_items) stage_frag(s_rec[0]);

    while (item < n_items) {
        const MergeItem it = s_rec[slot];
        const uint32_t nk = it.n_kmers;
        const uint32_t nw = (nk + 31) >> 5;
        const bool jumbo = it.jumbo_off != kNone;
        const uint32_t next_item = s_item[slot ^ 1];
        if (tid == 0) {
            s_item[slot] = atomicAdd(a.item_cursor, 1u);      // the item after next (this slot was read one iteration ago)
            if (!jumbo) {                                     // this tile's taxids (buffer idle since the last barrier)
                const uint64_t i0 = it.info_begin & ~3ull, i1 = (it.info_begin + nk + 3ull) & ~3ull;
                fence_proxy_async();
                mbar_expect_tx(mbar + 1, (unsigned)((i1 - i0) * 4));
                tma_load_1d(s_info, a.info + i0, (unsigned)((i1 - i0) * 4), mbar + 1);
            }
        }
        if (warp == 1 && lane < 4 && next_item < n_items)     // next item's record, four 16-byte words
            cp_async16(reinterpret_cast<unsigned char*>(s_rec + (slot ^ 1)) + 16 * lane,
                       reinterpret_cast<const unsigned char*>(a.items + next_item) + 16 * lane);
        const uint64_t* vals;
        const int32_t* infos;
        uint64_t qv_first = kBlank;
        if (jumbo) {
            // an amino-acid group larger than a shared-memory tile: values pre-decoded in HBM at load, taxids read from HBM; no
            // hash table — a query finds its group with two binary searches — but the same balanced pair passes as a tile, so a
            // group of 10^5 candidates (a universally conserved 8-mer of a large index) is spread over all threads, window by window
            if (warp == 1 && lane < 4) cp_async_wait_all();
            __syncthreads();
            if (tid == 0 && next_item < n_items) stage_frag(s_rec[slot ^ 1]);
            vals = a.jumbo_vals + it.jumbo_off;
            infos = a.info + it.info_begin;
        } else {
        for (uint32_t x = tid; x < a.n_buckets / 4; x += kThreads) reinterpret_cast<uint4*>(s_tab)[x] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
        // the first round's queries: pull this thread's second query (and both qinfo words) towards L2 now, load the first one
        // into a register after the decode — the probe then starts without a cold DRAM miss
        {
            const uint64_t q1 = it.q_begin + (uint64_t)tid;
            if (q1 < it.q_end) { prefetch_l2(a.q_value + q1); if (!via_idx) prefetch_l2(a.q_info + q1); }
            if (q1 + kThreads < it.q_end) { prefetch_l2(a.q_value + q1 + kThreads); if (!via_idx) prefetch_l2(a.q_info + q1 + kThreads); }
        }
        // -- 1. the tile's fragments were requested one item ago
        const uint64_t d0 = it.diff_begin, d1 = it.diff_begin + it.n_u16;
        const uint64_t a0 = d0 & ~7ull;
        mbar_wait(mbar, parity_bits & 1u);
        parity_bits ^= 1u;
        // -- 2. block-wide decode into the value array
        {
            uint64_t v = it.base_value, k = it.info_begin;
            const uint64_t kb = it.info_begin;
            block_decode<kThreads>(s_frag, (int)(d0 - a0), (int)(d0 - a0), (int)(d1 - a0), v, k, s_scan,
                                   [&](uint64_t kk, uint64_t val, uint64_t) {
                const uint64_t rel = kk - kb;
                if (rel < nk) s_vals[rel] = val;
            });
        }
        if (warp == 1 && lane < 4) cp_async_wait_all();
        __syncthreads();
        if (tid == 0 && next_item < n_items) stage_frag(s_rec[slot ^ 1]);      // streams in during the match phase
        if (it.q_begin + (uint64_t)tid < it.q_end) qv_first = ld_stream_u64(a.q_value + it.q_begin + tid);
        // -- 2b. amino-acid group starts: bitmap + hash table (the table was cleared before the decode barriers)
        for (uint32_t base = (uint32_t)warp * 32; base < nw * 32; base += kThreads) {
            const uint32_t rel = base + lane;
            bool start = false;
            uint64_t aa = 0;
            if (rel < nk) {
                aa = s_vals[rel] >> 24;
                start = rel == 0 || (s_vals[rel - 1] >> 24) != aa;
            }
            const uint32_t b = __ballot_sync(kFull, start);
            if (lane == 0) s_bits[base >> 5] = b;
            if (start) {
                const uint32_t h = aa_hash(aa);
                const uint32_t entry = (rel << 19) | (h & 0x7FFFFu);
                uint32_t bs = 2u * (h >> hash_shift);
                while (true) {
                    if (atomicCAS(&s_tab[bs], kEmpty, entry) == kEmpty) break;
                    if (atomicCAS(&s_tab[bs + 1], kEmpty, entry) == kEmpty) break;
                    bs = (bs + 2) & (2u * bucket_mask + 1u);
                }
            }
        }
        mbar_wait(mbar + 1, (parity_bits >> 1) & 1u);
        parity_bits ^= 2u;
        __syncthreads();
        vals = s_vals;
        infos = s_info + (it.info_begin & 3ull);
        }

        // -- 3. rounds of up to kQR queries (a jumbo item: few enough that hits x group length stays below 2^31 pairs)
        const uint32_t round_q = jumbo ? max(1u, min(kQR, 0x7FFFFFFFu / max(nk, 1u))) : kQR;
        for (uint64_t r0 = it.q_begin; r0 < it.q_end; r0 += round_q) {
            const uint32_t nq = (uint32_t)min((uint64_t)round_q, it.q_end - r0);
            // probe: hit list + owner slots of the first pair window
            for (uint32_t qb = (uint32_t)warp * 32; qb < nq; qb += kThreads) {
                const uint32_t q = qb + lane;
                const bool active = q < nq;
                const uint64_t qi = r0 + q;
                const uint64_t qv = (!jumbo && r0 == it.q_begin && qb == (uint32_t)warp * 32) ? qv_first : (active ? ld_stream_u64(a.q_value + qi) : kBlank);
                const uint64_t q40 = qv >> 24;
                uint32_t g0 = 0, len = 0;
                bool hit = false;
                if (active && jumbo) {
                    uint32_t lo = 0, hi = nk;
                    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((vals[mid] >> 24) < q40) lo = mid + 1; else hi = mid; }
                    if (lo < nk && (vals[lo] >> 24) == q40) {
                        g0 = lo; hit = true;
                        hi = nk;
                        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((vals[mid] >> 24) <= q40) lo = mid + 1; else hi = mid; }
                        len = lo - g0;
                    }
                } else if (active) {
                    const uint32_t h = aa_hash(q40);
                    const uint32_t tag = h & 0x7FFFFu;
                    uint32_t bkt = h >> hash_shift;
                    while (true) {
                        const uint2 e = *reinterpret_cast<const uint2*>(s_tab + 2u * bkt);
                        if (e.x == kEmpty) break;
                        if ((e.x & 0x7FFFFu) == tag && (vals[e.x >> 19] >> 24) == q40) { g0 = e.x >> 19; hit = true; break; }
                        if (e.y == kEmpty) break;
                        if ((e.y & 0x7FFFFu) == tag && (vals[e.y >> 19] >> 24) == q40) { g0 = e.y >> 19; hit = true; break; }
                        bkt = (bkt + 1) & bucket_mask;
                    }
                }
                const uint32_t bal = __ballot_sync(kFull, hit);
                if (!bal) continue;
                uint64_t qinfo = 0;
                if (hit) {
                    qinfo = via_idx ? a.q_info[a.q_idx[qi]] : ld_stream_u64(a.q_info + qi);
                    if (!jumbo) {                                   // group length = distance to the next group start
                        uint32_t w = (g0 + 1) >> 5;
                        uint32_t bits = w < nw ? s_bits[w] & (0xffffffffu << ((g0 + 1) & 31)) : 0u;
                        while (!bits && ++w < nw) bits = s_bits[w];
                        const uint32_t nxt = bits ? (w << 5) + (uint32_t)__ffs(bits) - 1u : nk;
                        len = min(nxt, nk) - g0;
                    }
                }
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
                const uint32_t total = __shfl_sync(kFull, incl, 31);
                uint32_t hb = 0, pb = 0;
                if (lane == 0) { hb = atomicAdd(s_nhit, (uint32_t)__popc(bal)); pb = atomicAdd(s_npair, total); }
                hb = __shfl_sync(kFull, hb, 0);
                pb = __shfl_sync(kFull, pb, 0);
                if (hit) {
                    const uint32_t h = hb + (uint32_t)__popc(bal & lt);
                    const uint32_t pbase = pb + incl - len;
                    s_hq[h] = qinfo;
                    s_hg[h] = g0;
                    s_hl[h] = len;
                    s_hp[h] = pbase;
                    s_hm[h] = 0xFF000000u | ((uint32_t)qv & 0xFFFFFFu);
                    for (uint32_t p = pbase, e = min(pbase + len, kPC); p < e; ++p) s_owner[p] = (uint16_t)h;
                }
            }
            __syncthreads();
            const uint32_t nh = *s_nhit, np = *s_npair;
            // pair p of the window starting at w0 -> (hit o, candidate j)
            auto fill = [&](uint32_t w0) {
                for (uint32_t h = tid; h < nh; h += kThreads) {
                    const uint32_t pbase = s_hp[h], len = s_hl[h];
                    const uint32_t b = max(pbase, w0), e = min(pbase + len, w0 + kPC);
                    for (uint32_t p = b; p < e; ++p) s_owner[p - w0] = (uint16_t)h;
                }
            };
            auto pass1 = [&](uint32_t w0, bool keep) {
                const uint32_t wend = min(np, w0 + kPC);
                for (uint32_t p = w0 + tid; p < wend; p += kThreads) {
                    const uint32_t o = s_owner[p - w0];
                    const uint32_t j = s_hg[o] + (p - s_hp[o]);
                    const uint32_t qd = s_hm[o] & 0xFFFFFFu;
                    const uint32_t td = (uint32_t)vals[j] & 0xFFFFFFu;
                    const uint32_t sum = td == qd ? 0u : ham_sum(ham_lookup(s_ham, qd, td));
                    if (keep) s_psum[p - w0] = (uint8_t)sum;
                    atomicMin(&s_hm[o], (sum << 24) | qd);
                }
            };
            auto pass2 = [&](uint32_t w0, bool kept) {
                const uint32_t wend = min(np, w0 + kPC);
                for (uint32_t p0 = w0 + (uint32_t)warp * 32; p0 < wend; p0 += kThreads) {
                    const uint32_t p = p0 + lane;
                    uint32_t o = 0, j = 0, qd = 0, td = 0, sum = 255u;
                    bool sel = false;
                    if (p < wend) {
                        o = s_owner[p - w0];
                        j = s_hg[o] + (p - s_hp[o]);
                        const uint32_t mq = s_hm[o];
                        qd = mq & 0xFFFFFFu;
                        td = (uint32_t)vals[j] & 0xFFFFFFu;
                        sum = kept ? (uint32_t)s_psum[p - w0] : (td == qd ? 0u : ham_sum(ham_lookup(s_ham, qd, td)));
                        sel = sum <= min((mq >> 24) * 2u, 7u);                  // KmerMatcher.cpp:1136
                    }
                    emit(sel, sel ? s_hq[o] : 0ull, sel ? infos[j] : 0, qd, td, sum);
                }
            };
            if (np != 0 && np <= kPC) {
                pass1(0, true);
                __syncthreads();
                pass2(0, true);
            } else if (np != 0) {
                for (uint32_t w0 = 0; w0 < np; w0 += kPC) {
                    __syncthreads();
                    fill(w0);
                    __syncthreads();
                    pass1(w0, false);
                }
                for (uint32_t w0 = 0; w0 < np; w0 += kPC) {
                    __syncthreads();
                    fill(w0);
                    __syncthreads();
                    pass2(w0, false);
                }
            }
            if (tid == 0) { *s_nhit = 0u; *s_npair = 0u; }
            __syncthreads();
        }
        item = next_item;
        slot ^= 1;
    }
    // blank out the unused tail of the warp's last chunk (seqID 0 == not a match)
    if (chunk.used < kOutChunk) {
        for (uint32_t w = chunk.used + lane; w < kOutChunk; w += 32) {
            const uint64_t sl = chunk.base + w;
            if (sl < a.out_cap) {
                uint64_t* o = reinterpret_cast<uint64_t*>(a.out + sl);
                o[0] = 0; o[1] = 0; o[2] = 0;
            }
        }
    }
    if (lane == 0 && my_matches) atomicAdd(a.out_count + 1, my_matches);
}

size_t merge_smem_bytes(uint32_t max_u16, uint32_t max_kmers, uint32_t n_buckets, int cta_threads) {
    return smem_layout2(max_u16, max_kmers, n_buckets, (uint32_t)cta_threads).total;
}

void launch_merge_plan(const MergeArgs& a, cudaStream_t st) {
    const unsigned blocks = (unsigned)((a.n_tiles + 1 + 255) / 256);
    merge_partition_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.n_tiles, a.q_value, a.n_query, a.prefix_shift, a.q_lo);
    merge_item_count_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.n_tiles, a.q_lo, a.item_cnt);
    exclusive_sum_u32(a.scan_tmp, a.scan_tmp_bytes, a.item_cnt, a.item_off, a.n_tiles + 1, st);
    merge_item_fill_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.n_tiles, a.q_lo, a.item_cnt, a.item_off, a.items, a.items_cap);
}

template <int kThreads>
static void launch_merge_v2(const MergeArgs& a, int sm_count, cudaStream_t st) {
    const size_t smem = smem_layout2(a.max_u16, a.max_kmers, a.n_buckets, kThreads).total;
    MBL_CUDA(cudaFuncSetAttribute(merge_kernel_v2<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MBL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_kernel_v2<kThreads>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    merge_kernel_v2<kThreads><<<(unsigned)(sm_count * per_sm), kThreads, smem, st>>>(a);
}

void launch_merge(const MergeArgs& a, int sm_count, cudaStream_t st) {
    if (a.cta_threads == 256) launch_merge_v2<256>(a, sm_count, st);
    else launch_merge_v2<512>(a, sm_count, st);
}

}  // namespace mbl
