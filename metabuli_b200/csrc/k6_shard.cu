// K6 — exchange staging of the index-sharded mode (no counterpart in the single-node reference, whose
// OpenMP threads all see the whole diffIdx: KmerMatcher.cpp:156-217).
//
// The diffIdx/info arrays are range-partitioned over the GPUs at amino-acid-group boundaries (mbl_plan_shards), so a query
// metamer can only match on the shard whose value range holds its amino-acid part.  Per batch:
//   read owner   K1 extract  -> bucket the metamers by owning shard (this file)        -> all-to-all #1 (value, qinfo)
//   shard owner  K2 sort + K3 merge over its shard -> bucket the matches by read owner -> all-to-all #2 (24-byte rows)
//   read owner   K4 sort + K5 score
// Bucketing = one stable radix pass over (bucket id : 8 bit, element index : 32 bit) and a gather into a contiguous send
// buffer, so the collective moves each element exactly once and the buckets need no host-side packing.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace mbl {

namespace {

// bucket of a query metamer: the last shard whose first amino-acid part is <= the query's; blanks go to bucket n (dropped)
__global__ void kmer_bucket_kernel(const uint64_t* __restrict__ value, uint64_t n, ShardBounds b, uint8_t* __restrict__ key,
                                   uint32_t* __restrict__ idx) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t v = value[i];
    uint32_t s = b.n;
    if (v != kBlank) {
        const uint64_t aa = v & kAaMask;
        s = 0;
        for (uint32_t k = 1; k < b.n; ++k) s += (b.bound[k] & kAaMask) <= aa ? 1u : 0u;     // bounds ascend
    }
    key[i] = (uint8_t)s;
    idx[i] = (uint32_t)i;
}

// bucket of a match row: the rank that owns the read (global 1-based seqID); blank rows (seqID 0) go to bucket n
__global__ void match_bucket_kernel(const mbl_match_rec* __restrict__ m, uint64_t n, ShardBounds b, uint8_t* __restrict__ key,
                                    uint32_t* __restrict__ idx) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t seq = qi_seq(m[i].qinfo);
    uint32_t s = b.n;
    if (seq != 0) {
        const uint64_t g = (uint64_t)seq - 1;
        s = 0;
        for (uint32_t k = 1; k < b.n; ++k) s += b.bound[k] <= g ? 1u : 0u;
    }
    key[i] = (uint8_t)s;
    idx[i] = (uint32_t)i;
}

// begin[k] = first position of the sorted bucket ids that is >= k, k = 0..n_buckets (begin[n_buckets] = elements to send)
__global__ void bucket_begin_kernel(const uint8_t* __restrict__ key, uint64_t n, uint32_t n_buckets, uint64_t* __restrict__ begin) {
    const uint32_t k = threadIdx.x;
    if (k > n_buckets) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (key[mid] < k) lo = mid + 1; else hi = mid; }
    begin[k] = lo;
}

// send buffers of all-to-all #1; seqIDs become global (seq_add = index of the batch's first read among all ranks' reads)
__global__ void kmer_gather_kernel(const uint32_t* __restrict__ idx, uint64_t n_send, const uint64_t* __restrict__ value,
                                   const uint64_t* __restrict__ qinfo, uint64_t seq_add, uint64_t* __restrict__ out_value,
                                   uint64_t* __restrict__ out_qinfo) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_send) return;
    const uint32_t s = idx[i];
    out_value[i] = value[s];
    out_qinfo[i] = qinfo[s] + (seq_add << 32);
}

// send buffer of all-to-all #2: one thread per 8-byte word of a 24-byte row
__global__ void match_gather3_kernel(const uint32_t* __restrict__ idx, uint64_t n_send, const uint64_t* __restrict__ in,
                                     uint64_t* __restrict__ out) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 3 * n_send) return;
    const uint64_t r = w / 3;
    out[w] = in[3ull * idx[r] + (w - 3 * r)];
}

// received rows -> the owner's local seqIDs (seq_sub = index of the batch's first read among all ranks' reads)
__global__ void match_localize_kernel(const uint64_t* __restrict__ in, uint64_t n, uint64_t seq_sub, uint64_t* __restrict__ out) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 3 * n) return;
    uint64_t x = in[w];
    if (w % 3 == 0) x -= seq_sub << 32;
    out[w] = x;
}

// fused bucket-gather + all-to-all: every row goes straight into the receive buffer of the rank that owns its bucket (a
// cudaMalloc'd buffer mapped through CUDA IPC, written over NVLink; the local bucket is an ordinary store).  Lanes of a warp
// write consecutive 8-byte words of one destination except at a bucket boundary, so the peer traffic is full 32-byte sectors.
__device__ __forceinline__ uint32_t bucket_of(const PushDst& d, uint64_t i) {
    uint32_t b = 0;
    while (b + 1 < d.n && d.begin[b + 1] <= i) ++b;
    return b;
}
__global__ void kmer_push_kernel(const uint32_t* __restrict__ idx, uint64_t n_send, const uint64_t* __restrict__ value,
                                 const uint64_t* __restrict__ qinfo, uint64_t seq_add, const __grid_constant__ PushDst d) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_send) return;
    const uint32_t b = bucket_of(d, i);
    const uint32_t s = idx[i];
    uint64_t* dst = reinterpret_cast<uint64_t*>(d.base[b]);
    const uint64_t row = d.row_off[b] + (i - d.begin[b]);
    dst[row] = value[s];                                        // receiver layout: values [0, total) | qinfo [total, 2 total)
    dst[d.total[b] + row] = qinfo[s] + (seq_add << 32);
}
__global__ void match_push_kernel(const uint32_t* __restrict__ idx, uint64_t n_send, const uint64_t* __restrict__ in,
                                  const __grid_constant__ PushDst d) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 3 * n_send) return;
    const uint64_t i = w / 3;
    const uint32_t b = bucket_of(d, i);
    uint64_t* dst = reinterpret_cast<uint64_t*>(d.base[b]);
    dst[3 * (d.row_off[b] + (i - d.begin[b])) + (w - 3 * i)] = in[3ull * idx[i] + (w - 3 * i)];
}

__global__ void iota_kernel(uint32_t* __restrict__ idx, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}

unsigned blocks_for(uint64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

size_t bucket_sort_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint8_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (long long)n, 0, 8);
    return bytes;
}

// stable partition of the element indices by bucket id; -> sorted indices (in idx_a or idx_b), begin[0..n_buckets] on the device
static const uint32_t* bucket_partition(void* tmp, size_t tmp_bytes, uint64_t n, uint32_t n_buckets, uint8_t* key_a, uint8_t* key_b,
                                        uint32_t* idx_a, uint32_t* idx_b, uint64_t* d_begin, cudaStream_t st) {
    int bits = 1;
    while ((1u << bits) <= n_buckets) ++bits;
    cub::DoubleBuffer<uint8_t> k(key_a, key_b);
    cub::DoubleBuffer<uint32_t> v(idx_a, idx_b);
    if (n) MBL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, 0, bits, st));
    bucket_begin_kernel<<<1, 128, 0, st>>>(k.Current(), n, n_buckets, d_begin);
    return v.Current();
}

const uint32_t* bucket_kmers(void* tmp, size_t tmp_bytes, const uint64_t* value, uint64_t n, const ShardBounds& b, uint8_t* key_a, uint8_t* key_b,
                             uint32_t* idx_a, uint32_t* idx_b, uint64_t* d_begin, cudaStream_t st) {
    if (n) kmer_bucket_kernel<<<blocks_for(n), 256, 0, st>>>(value, n, b, key_a, idx_a);
    return bucket_partition(tmp, tmp_bytes, n, b.n, key_a, key_b, idx_a, idx_b, d_begin, st);
}

const uint32_t* bucket_matches(void* tmp, size_t tmp_bytes, const mbl_match_rec* m, uint64_t n, const ShardBounds& b, uint8_t* key_a,
                               uint8_t* key_b, uint32_t* idx_a, uint32_t* idx_b, uint64_t* d_begin, cudaStream_t st) {
    if (n) match_bucket_kernel<<<blocks_for(n), 256, 0, st>>>(m, n, b, key_a, idx_a);
    return bucket_partition(tmp, tmp_bytes, n, b.n, key_a, key_b, idx_a, idx_b, d_begin, st);
}

void gather_kmers(const uint32_t* idx, uint64_t n_send, const uint64_t* value, const uint64_t* qinfo, uint64_t seq_add, uint64_t* out_value,
                  uint64_t* out_qinfo, cudaStream_t st) {
    if (n_send) kmer_gather_kernel<<<blocks_for(n_send), 256, 0, st>>>(idx, n_send, value, qinfo, seq_add, out_value, out_qinfo);
}

void gather_matches(const uint32_t* idx, uint64_t n_send, const mbl_match_rec* in, mbl_match_rec* out, cudaStream_t st) {
    if (n_send) match_gather3_kernel<<<blocks_for(3 * n_send), 256, 0, st>>>(idx, n_send, reinterpret_cast<const uint64_t*>(in),
                                                                           reinterpret_cast<uint64_t*>(out));
}

void localize_matches(const mbl_match_rec* in, uint64_t n, uint64_t seq_sub, mbl_match_rec* out, cudaStream_t st) {
    if (n) match_localize_kernel<<<blocks_for(3 * n), 256, 0, st>>>(reinterpret_cast<const uint64_t*>(in), n, seq_sub, reinterpret_cast<uint64_t*>(out));
}

void push_kmers(const uint32_t* idx, uint64_t n_send, const uint64_t* value, const uint64_t* qinfo, uint64_t seq_add, const PushDst& d,
                cudaStream_t st) {
    if (n_send) kmer_push_kernel<<<blocks_for(n_send), 256, 0, st>>>(idx, n_send, value, qinfo, seq_add, d);
}

void push_matches(const uint32_t* idx, uint64_t n_send, const mbl_match_rec* in, const PushDst& d, cudaStream_t st) {
    if (n_send) match_push_kernel<<<blocks_for(3 * n_send), 256, 0, st>>>(idx, n_send, reinterpret_cast<const uint64_t*>(in), d);
}

void launch_iota(uint32_t* idx, uint64_t n, cudaStream_t st) {
    if (n) iota_kernel<<<blocks_for(n), 256, 0, st>>>(idx, n);
}

}  // namespace mbl
