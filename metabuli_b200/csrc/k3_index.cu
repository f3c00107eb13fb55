// Load-time construction of the in-HBM tile directory over the differential index
// (reference: the 4096-entry `split` checkpoints, Kmer.h:111-119, IndexCreator.cpp:849-857, and their
// use in KmerMatcher.cpp:156-194).  The reference seeks to one of 4096 checkpoints and decodes serially;
// on the GPU the stream is cut every kCellU16 fragments into cells with their own (k-mer index, value)
// checkpoint, and into amino-acid-group-aligned tiles of ~kTileCells cells that the merge kernel stages
// in shared memory.  Built once per mbl_load_db, entirely on the device.
#include <algorithm>
#include <cub/cub.cuh>

#include "delta_decode.cuh"
#include "kernels.cuh"

namespace mbl {

namespace {
constexpr int kWarps = 8;
constexpr uint64_t kNone = ~0ull;
}

// pass 1: per cell, number of k-mers ending in it and the sum of their deltas
__global__ void __launch_bounds__(kWarps * 32)
cell_stats_kernel(const uint16_t* __restrict__ diff, uint64_t n_u16, uint64_t n_cells,
                  uint64_t* __restrict__ cell_cnt, uint64_t* __restrict__ cell_sum) {
    const int lane = threadIdx.x & 31;
    for (uint64_t c = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); c < n_cells; c += (uint64_t)gridDim.x * kWarps) {
        long long s = (long long)(c * kCellU16), e = (long long)min((c + 1) * (uint64_t)kCellU16, n_u16);
        uint64_t v = 0, k = 0;
        warp_decode(diff, 0, s, e, v, k, [](uint64_t, uint64_t, uint64_t, long long) {});
        if (lane == 0) { cell_cnt[c] = k; cell_sum[c] = v; }
    }
}

// pass 2 (after exclusive scans): first amino-acid-group start among the k-mers ending in each cell
__global__ void __launch_bounds__(kWarps * 32)
cell_boundary_kernel(const uint16_t* __restrict__ diff, uint64_t n_u16, uint64_t n_cells,
                     const uint64_t* __restrict__ cell_k, const uint64_t* __restrict__ cell_v,
                     uint64_t* __restrict__ b_kidx, uint64_t* __restrict__ b_off, uint64_t* __restrict__ b_base,
                     uint64_t* __restrict__ b_aa) {
    const int lane = threadIdx.x & 31;
    for (uint64_t c = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); c < n_cells; c += (uint64_t)gridDim.x * kWarps) {
        long long s = (long long)(c * kCellU16), e = (long long)min((c + 1) * (uint64_t)kCellU16, n_u16);
        uint64_t v = cell_v[c], k = cell_k[c];
        uint64_t bestK = kNone, bestOff = 0, bestBase = 0, bestAa = 0;
        warp_decode(diff, 0, s, e, v, k, [&](uint64_t kk, uint64_t val, uint64_t delta, long long first) {
            uint64_t prev = val - delta;
            bool start = (kk == 0) || (aa_part(val) != aa_part(prev));
            if (start && bestK == kNone) { bestK = kk; bestOff = (uint64_t)first; bestBase = prev; bestAa = aa_part(val); }
        });
        // the smallest k-mer index among lanes (lanes hold increasing index ranges)
        uint64_t m = bestK;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (m == kNone) { if (lane == 0) b_kidx[c] = kNone; }
        else if (bestK == m) { b_kidx[c] = bestK; b_off[c] = bestOff; b_base[c] = bestBase; b_aa[c] = bestAa; }
    }
}

// amino-acid presence filter: two bits per amino-acid group in the 1024-bit line it maps to (mbl_common.cuh)
__global__ void __launch_bounds__(kWarps * 32)
filter_build_kernel(const uint16_t* __restrict__ diff, uint64_t n_u16, uint64_t n_cells, const uint64_t* __restrict__ cell_k,
                    const uint64_t* __restrict__ cell_v, uint32_t* __restrict__ words, uint32_t n_lines, int minimizer) {
    for (uint64_t c = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); c < n_cells; c += (uint64_t)gridDim.x * kWarps) {
        long long s = (long long)(c * kCellU16), e = (long long)min((c + 1) * (uint64_t)kCellU16, n_u16);
        uint64_t v = cell_v[c], k = cell_k[c];
        warp_decode(diff, 0, s, e, v, k, [&](uint64_t kk, uint64_t val, uint64_t delta, long long) {
            if (kk != 0 && aa_part(val) == aa_part(val - delta)) return;           // same group as its predecessor: bits are set
            const uint64_t h = aa_filter_hash(val);
            uint32_t* blk = words + (size_t)aa_filter_line(val, h, n_lines, minimizer) * 32;
            const uint32_t b1 = aa_filter_bit1(h), b2 = aa_filter_bit2(h);
            atomicOr(blk + (b1 >> 5), 1u << (b1 & 31));
            atomicOr(blk + (b2 >> 5), 1u << (b2 & 31));
        });
    }
}

// one thread per nominal tile grid point: the cell that holds the first group start at or after it
__global__ void tile_candidate_kernel(const uint64_t* __restrict__ b_kidx, uint64_t n_cells, uint64_t n_grid, uint32_t tile_cells,
                                      uint64_t* __restrict__ cand) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_grid) return;
    uint64_t c = m * tile_cells;
    while (c < n_cells && b_kidx[c] == kNone) ++c;
    cand[m] = c < n_cells ? c : kNone;
}

__global__ void tile_flag_kernel(const uint64_t* __restrict__ cand, uint64_t n_grid, uint32_t* __restrict__ flag) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_grid) return;
    flag[m] = (cand[m] != kNone && (m == 0 || cand[m] != cand[m - 1])) ? 1u : 0u;
}

__global__ void tile_fill_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ flag,
                                 const uint32_t* __restrict__ rank, uint64_t n_grid, const uint64_t* __restrict__ b_kidx,
                                 const uint64_t* __restrict__ b_off, const uint64_t* __restrict__ b_base,
                                 const uint64_t* __restrict__ b_aa, Tile* __restrict__ tiles) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_grid || !flag[m]) return;
    uint64_t c = cand[m];
    Tile t;
    t.diff_begin = b_off[c];
    t.info_begin = b_kidx[c];
    t.base_value = b_base[c];
    t.first_aa = b_aa[c];
    t.n_u16 = 0; t.n_kmers = 0; t.jumbo_off = kNone; t.last_value = 0;
    tiles[rank[m]] = t;
}

// extents from the successor; Q1: the numerically last k-mer of the DB is never a candidate
// (KmerMatcher.cpp:378-380), so the directory simply does not contain it.
__global__ void tile_extent_kernel(Tile* __restrict__ tiles, uint64_t n_tiles, uint64_t n_u16, uint64_t n_kmers_eff, uint64_t final_value,
                                   uint32_t max_u16, uint32_t max_kmers, unsigned long long* __restrict__ jumbo_kmers,
                                   unsigned long long* __restrict__ jumbo_tiles) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    uint64_t dend = t + 1 < n_tiles ? tiles[t + 1].diff_begin : n_u16;
    uint64_t kend = t + 1 < n_tiles ? tiles[t + 1].info_begin : n_kmers_eff;
    if (kend > n_kmers_eff) kend = n_kmers_eff;
    uint64_t nk = kend > tiles[t].info_begin ? kend - tiles[t].info_begin : 0;
    uint64_t nu = dend - tiles[t].diff_begin;
    // the k-mer before the next tile is this tile's last one (for the final tile: the last k-mer of the stream)
    tiles[t].last_value = t + 1 < n_tiles ? tiles[t + 1].base_value : final_value;
    tiles[t].n_u16 = (uint32_t)min(nu, (uint64_t)0xFFFFFFFFu);
    tiles[t].n_kmers = (uint32_t)min(nk, (uint64_t)0xFFFFFFFFu);
    if (nu > max_u16 || nk > max_kmers) {
        tiles[t].jumbo_off = atomicAdd(jumbo_kmers, (unsigned long long)nk);
        atomicAdd(jumbo_tiles, 1ull);
    }
}

// jumbo tiles (one amino-acid group larger than a shared-memory tile) are decoded once into HBM
__global__ void __launch_bounds__(kWarps * 32)
jumbo_decode_kernel(const uint16_t* __restrict__ diff, const Tile* __restrict__ tiles, uint64_t n_tiles,
                    const uint64_t* __restrict__ cell_k, const uint64_t* __restrict__ cell_v,
                    uint64_t* __restrict__ jumbo_vals) {
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const Tile tl = tiles[t];
        if (tl.jumbo_off == kNone) continue;
        const uint64_t d0 = tl.diff_begin, d1 = tl.diff_begin + tl.n_u16;
        const uint64_t c0 = d0 / kCellU16, c1 = (d1 + kCellU16 - 1) / kCellU16;
        for (uint64_t c = c0 + (threadIdx.x >> 5); c < c1; c += kWarps) {
            long long s = (long long)max(c * (uint64_t)kCellU16, d0), e = (long long)min((c + 1) * (uint64_t)kCellU16, d1);
            uint64_t v, k;
            if ((uint64_t)s == d0) { v = tl.base_value; k = tl.info_begin; }
            else { v = cell_v[c]; k = cell_k[c]; }
            warp_decode(diff, (long long)d0, s, e, v, k, [&](uint64_t kk, uint64_t val, uint64_t, long long) {
                uint64_t rel = kk - tl.info_begin;
                if (rel < tl.n_kmers) jumbo_vals[tl.jumbo_off + rel] = val;
            });
        }
    }
}

// How coarse may the query sort be?  With queries ordered on value >> b only, a tile receives every query whose prefix lies
// in [prefix(first k-mer), prefix(last k-mer)]; that is (nearly) free when a tile spans many prefixes and wasteful when whole
// tiles sit inside one prefix.  cnt[0] / cnt[1] = tiles whose first and last k-mer share the 24-bit / 32-bit prefix.
__global__ void tile_prefix_kernel(const Tile* __restrict__ tiles, uint64_t n_tiles, unsigned long long* __restrict__ cnt) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles || tiles[t].n_kmers == 0) return;
    const uint64_t lo = tiles[t].first_aa, hi = tiles[t].last_value;     // first_aa = value with the DNA bits cleared
    if ((lo >> 40) == (hi >> 40)) atomicAdd(cnt, 1ull);
    if ((lo >> 32) == (hi >> 32)) atomicAdd(cnt + 1, 1ull);
}

// -------------------------------------------------------------------------------------------------
namespace { struct AddU64 { __host__ __device__ uint64_t operator()(uint64_t a, uint64_t b) const { return a + b; } }; }

// base_value: value the first delta of the stream is relative to (0 for a whole diffIdx file, the preceding k-mer's value for a
// shard that starts inside the file, mbl_plan_shards); holds_db_tail: the stream ends with the numerically last k-mer of the DB (Q1)
void build_tile_directory(const uint16_t* d_diff, uint64_t n_u16, uint64_t n_kmers, int sm_count, uint32_t tile_cells,
                          cudaStream_t st, TileDirectory& dir, uint64_t base_value, bool holds_db_tail, int filter_bits_per_kmer,
                          uint64_t filter_total_kmers, int filter_minimizer) {
    dir = TileDirectory();
    if (tile_cells < 1) tile_cells = 1;
    dir.tile_cells = tile_cells;
    dir.max_u16 = 2 * tile_cells * kCellU16 - 16;
    dir.max_kmers = tile_cells * kCellU16;
    if (n_u16 == 0 || n_kmers == 0) return;
    const uint64_t n_cells = (n_u16 + kCellU16 - 1) / kCellU16;
    const uint64_t n_grid = (n_cells + tile_cells - 1) / tile_cells;
    const uint64_t n_kmers_eff = n_kmers - (holds_db_tail ? 1 : 0);   // Q1
    uint64_t *cell_cnt, *cell_sum, *b_kidx, *b_off, *b_base, *b_aa, *cand;
    uint32_t *flag, *rank;
    MBL_CUDA(cudaMalloc(&cell_cnt, 8 * (n_cells + 1)));
    MBL_CUDA(cudaMalloc(&cell_sum, 8 * (n_cells + 1)));
    MBL_CUDA(cudaMalloc(&dir.cell_k, 8 * (n_cells + 1)));
    MBL_CUDA(cudaMalloc(&dir.cell_v, 8 * (n_cells + 1)));
    MBL_CUDA(cudaMalloc(&b_kidx, 8 * n_cells));
    MBL_CUDA(cudaMalloc(&b_off, 8 * n_cells));
    MBL_CUDA(cudaMalloc(&b_base, 8 * n_cells));
    MBL_CUDA(cudaMalloc(&b_aa, 8 * n_cells));
    MBL_CUDA(cudaMalloc(&cand, 8 * n_grid));
    MBL_CUDA(cudaMalloc(&flag, 4 * (n_grid + 1)));
    MBL_CUDA(cudaMalloc(&rank, 4 * (n_grid + 1)));
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_cells + kWarps - 1) / kWarps, (uint64_t)sm_count * 32);
    cell_stats_kernel<<<blocks, kWarps * 32, 0, st>>>(d_diff, n_u16, n_cells, cell_cnt, cell_sum);
    size_t tmp_bytes = 0, tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cell_cnt, dir.cell_k, n_cells, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, flag, rank, n_grid + 1, st);
    tmp_bytes = std::max(tmp_bytes, tb2);
    cub::DeviceScan::ExclusiveScan(nullptr, tb2, cell_sum, dir.cell_v, AddU64(), base_value, n_cells, st);
    tmp_bytes = std::max(tmp_bytes, tb2);
    void* tmp;
    MBL_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cell_cnt, dir.cell_k, n_cells, st);
    cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, cell_sum, dir.cell_v, AddU64(), base_value, n_cells, st);
    cell_boundary_kernel<<<blocks, kWarps * 32, 0, st>>>(d_diff, n_u16, n_cells, dir.cell_k, dir.cell_v, b_kidx, b_off, b_base, b_aa);
    if (filter_bits_per_kmer > 0) {
        // a shard sizes its filter for the whole index: the ranks OR their filters together (mbl_shard_filter_or)
        const uint64_t want = (std::max(n_kmers, filter_total_kmers) * (uint64_t)filter_bits_per_kmer + 1023) / 1024 + 1;
        dir.filter_lines = (uint32_t)std::min<uint64_t>(want, 0xFFFFFFF0ull);
        dir.filter_minimizer = filter_minimizer;
        MBL_CUDA(cudaMalloc(&dir.filter, 128 * (size_t)dir.filter_lines));
        MBL_CUDA(cudaMemsetAsync(dir.filter, 0, 128 * (size_t)dir.filter_lines, st));
        filter_build_kernel<<<blocks, kWarps * 32, 0, st>>>(d_diff, n_u16, n_cells, dir.cell_k, dir.cell_v, dir.filter, dir.filter_lines,
                                                            filter_minimizer);
    }
    tile_candidate_kernel<<<(unsigned)((n_grid + 255) / 256), 256, 0, st>>>(b_kidx, n_cells, n_grid, tile_cells, cand);
    MBL_CUDA(cudaMemsetAsync(flag, 0, 4 * (n_grid + 1), st));
    tile_flag_kernel<<<(unsigned)((n_grid + 255) / 256), 256, 0, st>>>(cand, n_grid, flag);
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, rank, n_grid + 1, st);
    uint32_t n_tiles32 = 0;
    uint64_t last_k = 0, last_c = 0, last_v = 0, last_s = 0;
    MBL_CUDA(cudaMemcpyAsync(&n_tiles32, rank + n_grid, 4, cudaMemcpyDeviceToHost, st));
    MBL_CUDA(cudaMemcpyAsync(&last_k, dir.cell_k + (n_cells - 1), 8, cudaMemcpyDeviceToHost, st));
    MBL_CUDA(cudaMemcpyAsync(&last_c, cell_cnt + (n_cells - 1), 8, cudaMemcpyDeviceToHost, st));
    MBL_CUDA(cudaMemcpyAsync(&last_v, dir.cell_v + (n_cells - 1), 8, cudaMemcpyDeviceToHost, st));
    MBL_CUDA(cudaMemcpyAsync(&last_s, cell_sum + (n_cells - 1), 8, cudaMemcpyDeviceToHost, st));
    MBL_CUDA(cudaStreamSynchronize(st));
    dir.n_kmers_decoded = last_k + last_c;
    dir.n_tiles = n_tiles32;
    dir.n_cells = n_cells;
    MBL_CUDA(cudaMalloc(&dir.tiles, sizeof(Tile) * (dir.n_tiles + 1)));
    tile_fill_kernel<<<(unsigned)((n_grid + 255) / 256), 256, 0, st>>>(cand, flag, rank, n_grid, b_kidx, b_off, b_base, b_aa, dir.tiles);
    unsigned long long* d_cnt;
    MBL_CUDA(cudaMalloc(&d_cnt, 16));
    MBL_CUDA(cudaMemsetAsync(d_cnt, 0, 16, st));
    tile_extent_kernel<<<(unsigned)((dir.n_tiles + 255) / 256), 256, 0, st>>>(dir.tiles, dir.n_tiles, n_u16, n_kmers_eff, last_v + last_s,
                                                                            dir.max_u16, dir.max_kmers, d_cnt, d_cnt + 1);
    unsigned long long h_cnt[2];
    MBL_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
    MBL_CUDA(cudaStreamSynchronize(st));
    dir.n_jumbo = h_cnt[1];
    dir.n_jumbo_kmers = h_cnt[0];
    if (dir.n_jumbo) {
        MBL_CUDA(cudaMalloc(&dir.jumbo_vals, 8 * (dir.n_jumbo_kmers + 1)));
        jumbo_decode_kernel<<<(unsigned)std::min<uint64_t>(dir.n_tiles, (uint64_t)sm_count * 8), kWarps * 32, 0, st>>>(
            d_diff, dir.tiles, dir.n_tiles, dir.cell_k, dir.cell_v, dir.jumbo_vals);
    }
    MBL_CUDA(cudaMemsetAsync(d_cnt, 0, 16, st));
    tile_prefix_kernel<<<(unsigned)((dir.n_tiles + 255) / 256), 256, 0, st>>>(dir.tiles, dir.n_tiles, d_cnt);
    MBL_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
    MBL_CUDA(cudaStreamSynchronize(st));
    MBL_CUDA(cudaGetLastError());
    dir.sort_begin_bit = h_cnt[0] * 32 <= dir.n_tiles ? 40 : h_cnt[1] * 32 <= dir.n_tiles ? 32 : 24;
    cudaFree(cell_cnt); cudaFree(cell_sum); cudaFree(b_kidx); cudaFree(b_off); cudaFree(b_base); cudaFree(b_aa);
    cudaFree(cand); cudaFree(flag); cudaFree(rank); cudaFree(tmp); cudaFree(d_cnt);
}

__global__ void filter_or_kernel(uint4* __restrict__ w, const uint4* __restrict__ o, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 a = w[i];
        const uint4 b = o[i];
        a.x |= b.x; a.y |= b.y; a.z |= b.z; a.w |= b.w;
        w[i] = a;
    }
}
void launch_filter_or(uint32_t* words, const uint32_t* other, uint64_t n_words, cudaStream_t st) {
    if (n_words) filter_or_kernel<<<148 * 16, 256, 0, st>>>(reinterpret_cast<uint4*>(words), reinterpret_cast<const uint4*>(other), n_words / 4);
}

void free_tile_directory(TileDirectory& dir) {
    cudaFree(dir.tiles); cudaFree(dir.cell_k); cudaFree(dir.cell_v); cudaFree(dir.jumbo_vals); cudaFree(dir.filter);
    dir = TileDirectory();
}

}  // namespace mbl
