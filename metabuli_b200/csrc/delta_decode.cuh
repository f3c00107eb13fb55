// Warp-parallel decoder of the differential index (reference delta codec: KmerMatcher.h:282-297,
// writer IndexCreator.cpp:874-892).
//
// Stream: u16 fragments, 15 payload bits each, most-significant group first, bit 15 marks the last
// fragment of a k-mer; value_i = value_{i-1} + delta_i.  k-mer boundaries are locally decidable (the end
// flag sits on the fragment itself), so the stream decodes as a scan: a lane takes 8 consecutive
// fragments (one 16-byte load), rebuilds the deltas of the k-mers that END in its octet (looking back at
// most 4 fragments for a k-mer that started earlier), and a warp shuffle scan turns per-lane
// (count, sum) into k-mer indices and absolute values.  256 fragments per warp iteration.
#pragma once
#include "mbl_common.cuh"

namespace mbl {

// Decode the fragments [s, e) of `frag` (indices are relative to the pointer; the pointer itself and
// index 0 must be 16-byte aligned, fragments [lo, s) must be readable for look-back, lo <= s).  Fragments
// below `lo` are treated as cuts.  For every k-mer whose LAST fragment lies in [s, e), in stream order,
// calls emit(k, value, delta, first_frag) with k counted from k0 and value accumulated from v0.
// Returns through k0 / v0 the running totals after the range.  All 32 lanes must call.
template <class Emit>
__device__ __forceinline__ void warp_decode(const uint16_t* frag, long long lo, long long s, long long e,
                                            uint64_t& v0, uint64_t& k0, Emit emit) {
    const int lane = threadIdx.x & 31;
    for (long long p0 = s & ~7ll; p0 < e; p0 += 256) {
        const long long p = p0 + 8ll * lane;
        uint16_t f[8];
        uint16_t lb[4];
        if (p < e) {
            uint4 v = *reinterpret_cast<const uint4*>(frag + p);
            f[0] = (uint16_t)v.x; f[1] = (uint16_t)(v.x >> 16); f[2] = (uint16_t)v.y; f[3] = (uint16_t)(v.y >> 16);
            f[4] = (uint16_t)v.z; f[5] = (uint16_t)(v.z >> 16); f[6] = (uint16_t)v.w; f[7] = (uint16_t)(v.w >> 16);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0;
        }
        if (p < e && p - 4 >= (lo & ~3ll) && p >= 4) {
            uint2 v = *reinterpret_cast<const uint2*>(frag + p - 4);
            lb[0] = (uint16_t)v.x; lb[1] = (uint16_t)(v.x >> 16); lb[2] = (uint16_t)v.y; lb[3] = (uint16_t)(v.y >> 16);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) lb[j] = 0x8000u;
        }
        // incoming partial delta and first-fragment index from the look-back quartet
        uint64_t acc = 0;
        long long kstart = p - 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long idx = p - 4 + j;
            const bool cut = (lb[j] & 0x8000u) || idx < lo;
            acc = cut ? 0ull : ((acc << 15) | lb[j]);
            if (cut) kstart = idx + 1;
        }
        // pass 1: count and sum of the deltas ending in this octet
        uint32_t cnt = 0;
        uint64_t sum = 0;
        {
            uint64_t a = acc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long idx = p + j;
                const bool dead = idx < lo;
                a = dead ? 0ull : ((a << 15) | (uint64_t)(f[j] & 0x7FFFu));
                if (!dead && (f[j] & 0x8000u)) {
                    if (idx >= s && idx < e) { ++cnt; sum += a; }
                    a = 0;
                }
            }
        }
        // warp inclusive scan of (cnt, sum)
        uint32_t icnt = cnt;
        uint64_t isum = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t c2 = __shfl_up_sync(0xffffffffu, icnt, o);
            uint64_t s2 = __shfl_up_sync(0xffffffffu, isum, o);
            if (lane >= o) { icnt += c2; isum += s2; }
        }
        uint64_t k = k0 + (icnt - cnt);
        uint64_t val = v0 + (isum - sum);
        // pass 2: re-walk and emit
        {
            uint64_t a = acc;
            long long ks = kstart;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long idx = p + j;
                const bool dead = idx < lo;
                a = dead ? 0ull : ((a << 15) | (uint64_t)(f[j] & 0x7FFFu));
                if (dead) ks = idx + 1;
                if (!dead && (f[j] & 0x8000u)) {
                    if (idx >= s && idx < e) {
                        val += a;
                        emit(k, val, a, ks);
                        ++k;
                    }
                    a = 0;
                    ks = idx + 1;
                }
            }
        }
        k0 += __shfl_sync(0xffffffffu, icnt, 31);
        v0 += __shfl_sync(0xffffffffu, isum, 31);
    }
}


// Block-wide variant for the merge kernel: all kThreadsDecode threads take one octet each (2048 fragments per
// sweep), so a tile of a few KiB keeps every warp busy and needs no per-cell checkpoints.  The (count, sum) pairs
// are scanned inside each warp with shuffles and across warps through a tiny double-buffered shared array (one
// __syncthreads per sweep).  `scan` must hold 2 * 2 * (threads/32) uint64_t.  Indices are 32-bit (a tile is a
// few thousand fragments); octets that lie completely inside [s, e) with their look-back inside [lo, ...) take a
// path without per-element range tests.  All threads of the block must call; v0 / k0 are block-uniform totals.
// emit(k, value, delta) — no first-fragment index here.
template <int kThreadsDecode, class Emit>
__device__ __forceinline__ void block_decode(const uint16_t* frag, int lo, int s, int e, uint64_t& v0, uint64_t& k0, uint64_t* scan, Emit emit) {
    constexpr int kW = kThreadsDecode / 32;
    static_assert(kW <= 32, "one lane per warp in the cross-warp scan");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int flip = 0;
    for (int p0 = s & ~7; p0 < e; p0 += 8 * kThreadsDecode, flip ^= 1) {
        const int p = p0 + 8 * (int)threadIdx.x;
        const bool warp_has_work = p0 + 256 * warp < e;     // warp-uniform: a 4 KiB tile keeps only the first 8 warps busy
        uint32_t w[4] = {0u, 0u, 0u, 0u};          // the octet as four pairs of fragments
        uint32_t l0 = 0x80008000u, l1 = 0x80008000u;   // look-back quartet (cuts when unavailable)
        const bool inner = p >= s && p + 8 <= e && p - 4 >= lo;
        uint64_t acc = 0;
        uint32_t cnt = 0, icnt = 0;
        uint64_t sum = 0, isum = 0;
        if (warp_has_work) {
            if (p < e) {
                const uint4 v = *reinterpret_cast<const uint4*>(frag + p);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                if (p >= 4 && p - 4 >= (lo & ~3)) { const uint2 b = *reinterpret_cast<const uint2*>(frag + p - 4); l0 = b.x; l1 = b.y; }
            }
            if (inner) {
                // incoming partial delta: fragments after the last end flag of the look-back quartet
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t fr = ((j & 2 ? l1 : l0) >> (16 * (j & 1))) & 0xFFFFu;
                    acc = (fr & 0x8000u) ? 0ull : ((acc << 15) | fr);
                }
                uint64_t a = acc;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t fr = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                    a = (a << 15) | (uint64_t)(fr & 0x7FFFu);
                    if (fr & 0x8000u) { ++cnt; sum += a; a = 0; }
                }
            } else if (p < e) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = p - 4 + j;
                    const uint32_t fr = ((j & 2 ? l1 : l0) >> (16 * (j & 1))) & 0xFFFFu;
                    const bool cut = (fr & 0x8000u) || idx < lo;
                    acc = cut ? 0ull : ((acc << 15) | fr);
                }
                uint64_t a = acc;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int idx = p + j;
                    const uint32_t fr = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                    const bool dead = idx < lo;
                    a = dead ? 0ull : ((a << 15) | (uint64_t)(fr & 0x7FFFu));
                    if (!dead && (fr & 0x8000u)) {
                        if (idx >= s && idx < e) { ++cnt; sum += a; }
                        a = 0;
                    }
                }
            }
            icnt = cnt;
            isum = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t c2 = __shfl_up_sync(0xffffffffu, icnt, o);
                const uint64_t s2 = __shfl_up_sync(0xffffffffu, isum, o);
                if (lane >= o) { icnt += c2; isum += s2; }
            }
        }
        uint64_t* sc = scan + flip * 2 * kW;
        if (lane == 31) { sc[warp] = icnt; sc[kW + warp] = isum; }
        __syncthreads();
        // cross-warp scan: lane x holds the totals of warp x, a shuffle scan gives this warp's offset and the block totals
        uint32_t c = 0;
        uint64_t v = 0;
        if (lane < kW) { c = (uint32_t)sc[lane]; v = sc[kW + lane]; }
        uint32_t ic = c;
        uint64_t iv = v;
#pragma unroll
        for (int o = 1; o < kW; o <<= 1) {
            const uint32_t c2 = __shfl_up_sync(0xffffffffu, ic, o);
            const uint64_t s2 = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) { ic += c2; iv += s2; }
        }
        const uint32_t wk = __shfl_sync(0xffffffffu, ic - c, warp);
        const uint64_t wv = __shfl_sync(0xffffffffu, iv - v, warp);
        const uint32_t tk = __shfl_sync(0xffffffffu, ic, kW - 1);
        const uint64_t tv = __shfl_sync(0xffffffffu, iv, kW - 1);
        if (cnt) {
            uint64_t k = k0 + wk + (icnt - cnt);
            uint64_t val = v0 + wv + (isum - sum);
            uint64_t a = acc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int idx = p + j;
                const uint32_t fr = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                const bool dead = !inner && idx < lo;
                a = dead ? 0ull : ((a << 15) | (uint64_t)(fr & 0x7FFFu));
                if (!dead && (fr & 0x8000u)) {
                    if (inner || (idx >= s && idx < e)) {
                        val += a;
                        emit(k, val, a);
                        ++k;
                    }
                    a = 0;
                }
            }
        }
        k0 += tk;
        v0 += tv;
    }
}

}  // namespace mbl
