// K5 — per-read taxon scoring kernel (reference rows A10-A12, Classifier.cpp:166-208 +
// Taxonomer.cpp).  One read per thread over the match list sorted in the reference's order; the
// algorithm itself is in score_core.cuh (shared with the CPU unit tests).  Read-level parallelism is
// ample (10^6-10^7 reads per batch); all state is in HBM scratch sized by the match count.
#include "score_core.cuh"

namespace mbl {

__global__ void __launch_bounds__(128, 4) score_kernel(ScoreArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_reads) return;
    score_read(a, a.read_perm ? a.read_perm[a.read_begin + i] : a.read_begin + i);
}

__global__ void taxcnt_len_kernel(const mbl_read_result* __restrict__ res, uint32_t n, uint32_t* __restrict__ len) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) len[r] = res[r].taxcnt_len;
    if (r == n) len[r] = 0;
}

__global__ void compact_taxcnt_kernel(const mbl_read_result* __restrict__ res, uint32_t n, const uint32_t* __restrict__ quot_off,
                                      const int32_t* __restrict__ pairs_in, const uint32_t* __restrict__ out_off,
                                      int32_t* __restrict__ pairs_out, mbl_read_result* __restrict__ res_out) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    mbl_read_result x = res[r];
    const int32_t* src = pairs_in + 2ull * quot_off[r];
    int32_t* dst = pairs_out + 2ull * out_off[r];
    for (uint32_t k = 0; k < 2 * x.taxcnt_len; ++k) dst[k] = src[k];
    x.taxcnt_begin = out_off[r];
    res_out[r] = x;
}

void launch_score(const ScoreArgs& a, cudaStream_t st) {
    if (!a.n_reads) return;
    score_kernel<<<(a.n_reads + 127) / 128, 128, 0, st>>>(a);
}
void launch_taxcnt_len(const mbl_read_result* results, uint32_t n_reads, uint32_t* len, cudaStream_t st) {
    taxcnt_len_kernel<<<(n_reads + 1 + 255) / 256, 256, 0, st>>>(results, n_reads, len);
}
void launch_compact_taxcnt(const mbl_read_result* results, uint32_t n_reads, const uint32_t* quot_off, const int32_t* pairs_in,
                           const uint32_t* out_off, int32_t* pairs_out, mbl_read_result* results_out, cudaStream_t st) {
    if (!n_reads) return;
    compact_taxcnt_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(results, n_reads, quot_off, pairs_in, out_off, pairs_out, results_out);
}

}  // namespace mbl
