// Shared definitions of the B200 classify path: PODs, bit layouts, error plumbing.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/metabuli_b200.h"

#define MBL_HD __host__ __device__ __forceinline__

namespace mbl {

// ---- k-mer / qinfo bit layout (reference Kmer.h:11-16, KmerScanner.h:110-114) --------------------
constexpr uint64_t kDnaMask = 0xFFFFFFull;          // low 24 bits: 8 x 3-bit synonymous-codon ids
constexpr uint64_t kAaMask = ~kDnaMask;             // high 40 bits: 8 x 5-bit amino acids
constexpr uint64_t kBlank = 0xFFFFFFFFFFFFFFFFull;  // blank slot marker (sorts last; never matches)

MBL_HD uint64_t aa_part(uint64_t v) { return v & kAaMask; }
MBL_HD uint64_t pack_qinfo(uint32_t seqId, uint32_t pos, uint32_t frame) {
    return (uint64_t)pos | ((uint64_t)(seqId & 0x1FFFFFFFu) << 32) | ((uint64_t)(frame & 7u) << 61);
}
MBL_HD uint32_t qi_pos(uint64_t q) { return (uint32_t)q; }
MBL_HD uint32_t qi_seq(uint64_t q) { return (uint32_t)((q >> 32) & 0x1FFFFFFFu); }
MBL_HD uint32_t qi_frame(uint64_t q) { return (uint32_t)(q >> 61); }

// ---- amino-acid presence filter (k3_index.cu builds it, K1 probes it) ---------------------------------
// Blocked Bloom filter over the 40-bit amino-acid parts of the index: one 128-byte line per key (what one DRAM access brings
// in anyway — ncu shows 126 bytes of DRAM traffic per random 32-byte probe), two bits inside it.  A query whose amino-acid part
// is not in the filter cannot have a candidate (KmerMatcher.cpp:363-375 would skip it), so K1 drops it before the sort; false
// positives only cost work.
// Line choice (format-2 k-mers: eight 5-bit residues): by the MINIMIZER of the 8-mer — the smallest hash among its three
// 6-residue sub-words.  Consecutive windows of a frame share seven residues and, about half of the time, the minimizer, so
// the ~10 windows per frame a K1 warp handles at once fall into ~5 lines instead of 10: fewer DRAM lines per read.  Format-1
// k-mers (base-21 amino-acid part) hash the whole part.
struct AaFilter {
    const uint32_t* words = nullptr;   // n_lines x 32 words
    uint32_t n_lines = 0;
    int minimizer = 0;                 // 1: line from the 6-residue minimizer (kmer_format 2)
};
MBL_HD uint32_t aa_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
MBL_HD uint64_t aa_filter_hash(uint64_t value) {
    uint64_t h = (value >> 24) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
    return h * 0xD6E8FEB86659FD93ull;
}
// line of a k-mer value; h = aa_filter_hash(value)
MBL_HD uint32_t aa_filter_line(uint64_t value, uint64_t h, uint32_t n_lines, int minimizer) {
    uint32_t key = (uint32_t)(h >> 32);
    if (minimizer) {
        const uint64_t aa = value >> 24;
        const uint32_t m0 = aa_mix32((uint32_t)(aa >> 10) & 0x3FFFFFFFu), m1 = aa_mix32((uint32_t)(aa >> 5) & 0x3FFFFFFFu),
                       m2 = aa_mix32((uint32_t)aa & 0x3FFFFFFFu);
        key = m0 < m1 ? (m0 < m2 ? m0 : m2) : (m1 < m2 ? m1 : m2);
    }
    return (uint32_t)(((uint64_t)key * (uint64_t)n_lines) >> 32);
}
// both bits sit in the same 32-byte sector of the line (bits 8-9 pick the sector), so a probe is one sector request
MBL_HD uint32_t aa_filter_bit1(uint64_t h) { return ((uint32_t)(h >> 8) & 0x300u) | ((uint32_t)(h >> 12) & 255u); }
MBL_HD uint32_t aa_filter_bit2(uint64_t h) { return ((uint32_t)(h >> 8) & 0x300u) | ((uint32_t)(h >> 22) & 255u); }

// LocalUtil.h:46-60
MBL_HD int max_covered_length(int len) {
    int r = len % 3;
    return r == 2 ? len - 2 : (r == 1 ? len - 4 : len - 3);
}
// windows per frame = cov/3 - 7; the reference's kmerCnt is 6x this (LocalUtil.h:46-49)
MBL_HD int windows_per_frame(int len) { return max_covered_length(len) / 3 - 7; }

// ---- in-HBM tile directory of the differential index -------------------------------------------
// A tile is an amino-acid-group-aligned run of k-mers; the decoder additionally restarts at fixed
// cells of kCellU16 fragments, each with its own (k-mer index, running value) checkpoint.
constexpr int kCellU16 = 1024;            // decode checkpoint grid (u16 fragments)

struct Tile {
    uint64_t diff_begin;     // u16 index of the first fragment of the tile's first k-mer
    uint64_t info_begin;     // k-mer index of the tile's first k-mer
    uint64_t base_value;     // value of the k-mer preceding the tile (0 at the stream start)
    uint64_t first_aa;       // amino-acid part of the tile's first k-mer, in place (value with the DNA bits cleared)
    uint32_t n_u16;          // fragments in the tile
    uint32_t n_kmers;        // k-mers in the tile (after the Q1 trim of the very last k-mer)
    uint64_t jumbo_off;      // offset into the pre-decoded value array, or ~0 when not jumbo
    uint64_t last_value;     // value of the tile's last k-mer (bucket geometry of the merge kernel)
};

struct alignas(16) MergeItem {   // one unit of merge work: a slice of the queries of one tile, with the tile's
    uint64_t q_begin, q_end;     // geometry copied in so the merge kernel needs a single 64-byte fetch per item
    uint64_t diff_begin, info_begin;
    uint64_t base_value, jumbo_off;
    uint32_t n_u16, n_kmers;
    uint32_t tile, pad;
};
static_assert(sizeof(MergeItem) == 64, "MergeItem is fetched as four 16-byte words");

// ---- error handling ------------------------------------------------------------------------------
struct CudaError {
    cudaError_t code;
    const char* file;
    int line;
};
#define MBL_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) throw ::mbl::CudaError{_e, __FILE__, __LINE__};                 \
    } while (0)

}  // namespace mbl
