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
//                    (Round 2 tried a two-level order — three radix passes over (32-bit seqID, index) and a warp per read that
//                    orders its ~150 rows in shared memory while gathering them: bit-identical, but 99-106 ms against 74.6 ms
//                    for this single-key path on the benchmark, as a rank-by-counting and as a bitonic network alike: a
//                    256-key network costs ~3 k warp instructions per read, more than the three 8-byte radix passes it saves.)
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
    if (seq_bits + sp_bits + 3 + pos3_bits <= 64) {
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
