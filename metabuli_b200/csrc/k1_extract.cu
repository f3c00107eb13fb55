// K1 — six-frame translation + metamer packing (reference rows A0-A3': LocalUtil.h:46-60,
// KmerScanner.h:74-117, KmerExtractor.cpp:292-373, 429-481).
//
// B200 formulation: every 24-nt window of a read is independent, so instead of the
// reference's six rolling scanners a thread owns one leftmost position x and emits the forward and the
// reverse-strand metamer that cover [x, x+23].  Valid x are the contiguous range [0, 3W) with
// W = cov/3 - 7 windows per frame; the reference's kmerCnt = 6W slots are laid out as
//   slot = read_base + strand * 3W + x          (order inside a read is irrelevant downstream).
// A warp owns a read: bases are pulled with 16-byte loads into shared memory, turned into per-position
// codon bytes ((aa << 3) | codon_id, 0xFF when a base is not ACGT) once, and each window is 8 byte
// lookups per strand.  Windows with an invalid codon become blank slots (value = UINT64_MAX).
#include "kernels.cuh"

namespace mbl {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kChunk = 480;              // leftmost positions per pass of a warp
constexpr int kCodonBuf = kChunk + 32;   // positions x0 .. x0+kChunk+21 (+ slack)
constexpr int kRawBuf = kChunk + 64;     // raw bases incl. alignment slack

struct WarpScratch {
    alignas(16) uint8_t raw[kRawBuf];
    uint8_t fwd[kCodonBuf];
    uint8_t rev[kCodonBuf];
};

__device__ __forceinline__ uint4 ld_stream_16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

}  // namespace

// thread per read: covered lengths, windows per frame per mate, slot count (KmerExtractor.cpp:429-481)
__global__ void read_meta_kernel(const uint64_t* __restrict__ off1, const uint64_t* __restrict__ off2,
                                 uint32_t n_reads, int32_t* __restrict__ cov1, int32_t* __restrict__ cov2,
                                 int32_t* __restrict__ w1, int32_t* __restrict__ w2,
                                 uint64_t* __restrict__ slots, uint32_t* __restrict__ quot_cnt) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int l1 = (int)(off1[r + 1] - off1[r]);
    int c1 = max_covered_length(l1);
    int a = windows_per_frame(l1);
    int c2 = 0, b = 0;
    bool paired = off2 != nullptr;
    if (paired) {
        int l2 = (int)(off2[r + 1] - off2[r]);
        c2 = max_covered_length(l2);
        b = windows_per_frame(l2);
    }
    // a read (pair) contributes only when every mate has at least one k-mer (KmerExtractor.cpp:436-471)
    bool empty = a < 1 || (paired && b < 1);
    if (empty) { a = 0; b = 0; }
    cov1[r] = c1;
    cov2[r] = c2;
    w1[r] = a;
    w2[r] = b;
    slots[r] = 6ull * (uint64_t)(a + b);
    // entries of the per-read quotient table used by scoring: (queryLength + 3) / 3 + 1 (Taxonomer.cpp:210)
    int ql = c1 + c2;
    quot_cnt[r] = ql + 3 > 0 ? (uint32_t)((ql + 3) / 3 + 1) : 1u;
}

// probe of the amino-acid presence filter (mbl_common.cuh): both bits live in one 128-byte line
__device__ __forceinline__ bool aa_filter_pass(const AaFilter& f, uint64_t value) {
    const uint64_t h = aa_filter_hash(value);
    const uint32_t* line = f.words + (size_t)aa_filter_line(value, h, f.n_lines, f.minimizer) * 32;
    const uint32_t b1 = aa_filter_bit1(h), b2 = aa_filter_bit2(h);
    const uint32_t w1 = __ldg(line + (b1 >> 5)), w2 = __ldg(line + (b2 >> 5));
    return ((w1 >> (b1 & 31)) & (w2 >> (b2 & 31)) & 1u) != 0u;
}

// closed-syncmer test of a format-2 amino-acid part (SyncmerScanner.h:36-74): of the 9 - s s-mers of the 8-residue window the
// smallest one (leftmost on ties: the reference's deque only evicts strictly larger values) is the first or the last.  The
// window is decided by its own residues alone, so it fits the one-thread-per-window shape of this kernel.
__device__ __forceinline__ bool is_closed_syncmer(uint64_t aa40, int s) {
    const int n = 9 - s;
    const uint64_t mask = (1ull << (5 * s)) - 1;
    const uint64_t first = (aa40 >> (5 * (n - 1))) & mask, last = aa40 & mask;
    bool first_min = first <= last, last_min = last < first;
    for (int j = 1; j < n - 1; ++j) {
        const uint64_t m = (aa40 >> (5 * (n - 1 - j))) & mask;
        first_min = first_min && first <= m;
        last_min = last_min && last < m;
    }
    return first_min || last_min;
}

template <int FORMAT, bool FILTER>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
extract_kernel(const uint8_t* __restrict__ bases1, const uint64_t* __restrict__ off1,
               const uint8_t* __restrict__ bases2, const uint64_t* __restrict__ off2, uint32_t n_reads,
               const int32_t* __restrict__ cov1, const int32_t* __restrict__ w1, const int32_t* __restrict__ w2,
               const uint64_t* __restrict__ slot_off, const uint8_t* __restrict__ g_base_code,
               const uint8_t* __restrict__ g_codon, uint64_t* __restrict__ value, uint64_t* __restrict__ qinfo,
               uint32_t* __restrict__ slot_idx, unsigned long long* __restrict__ n_valid, const AaFilter filter,
               unsigned long long* __restrict__ out_cursor, const uint64_t out_cap, const int smer_len) {
    __shared__ uint8_t s_code[256];
    __shared__ uint8_t s_codon[512];
    __shared__ WarpScratch s_warp[kWarpsPerBlock];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = g_base_code[i];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_codon[i] = g_codon[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = s_warp[warp];
    unsigned valid_cnt = 0;
    // FILTER: survivors are packed from slot 0 upwards; a warp owns a chunk of kExtractChunk slots at a time
    uint64_t chunk_base = 0;
    uint32_t chunk_used = kExtractChunk;
    unsigned pass_cnt = 0;
    for (uint32_t r = blockIdx.x * kWarpsPerBlock + warp; r < n_reads; r += gridDim.x * kWarpsPerBlock) {
        const int wa = w1[r], wb = w2[r];
        if (wa + wb == 0) continue;
        uint64_t slot_base = slot_off[r];
        for (int mate = 0; mate < 2; ++mate) {
            const int W = mate ? wb : wa;
            if (W <= 0) continue;
            const uint8_t* bases = mate ? bases2 : bases1;
            const uint64_t b0 = mate ? off2[r] : off1[r];
            const int L = (int)((mate ? off2[r + 1] : off1[r + 1]) - b0);
            const int npos = 3 * W;
            const uint32_t pos_off = mate ? (uint32_t)(cov1[r] + 3) : 0u;     // KmerExtractor.cpp:322-329
            const int lmod = L % 3;
            for (int x0 = 0; x0 < npos; x0 += kChunk) {
                const int cnt = min(kChunk, npos - x0);
                // 1. raw bases [x0, x0+cnt+23] via aligned 16-byte loads
                const uint64_t gbeg = b0 + (uint64_t)x0;
                const uint64_t abeg = gbeg & ~15ull;
                const int shift = (int)(gbeg - abeg);
                const int nbytes = shift + cnt + 23;
                for (int v = lane; v * 16 < nbytes; v += 32)
                    *reinterpret_cast<uint4*>(ws.raw + v * 16) = ld_stream_16(bases + abeg + (uint64_t)v * 16);
                __syncwarp();
                // 2. codon bytes of both strands for positions x0 .. x0+cnt+20
                for (int i = lane; i < cnt + 21; i += 32) {
                    unsigned c0 = s_code[ws.raw[shift + i]], c1 = s_code[ws.raw[shift + i + 1]], c2 = s_code[ws.raw[shift + i + 2]];
                    bool ok = (c0 | c1 | c2) < 4;
                    // reverse strand: complement = code ^ 2, read right to left (KmerScanner.h:95-97)
                    ws.fwd[i] = ok ? s_codon[c0 * 64 + c1 * 8 + c2] : (uint8_t)0xFF;
                    ws.rev[i] = ok ? s_codon[(c2 ^ 2) * 64 + (c1 ^ 2) * 8 + (c0 ^ 2)] : (uint8_t)0xFF;
                }
                __syncwarp();
                // 3. one forward and one reverse metamer per leftmost position
                for (int i0 = 0; i0 < cnt; i0 += 32) {
                    const int i = i0 + lane;
                    const bool active = i < cnt;
                    if (!FILTER && !active) continue;
                    const int x = x0 + (active ? i : 0);
                    uint64_t aaF = 0, aaR = 0;
                    uint32_t dnaF = 0, dnaR = 0;
                    unsigned badF = 0, badR = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        unsigned f, rv;
                        const int ii = active ? i : 0;
                        if (FORMAT == 2) { f = ws.fwd[ii + 3 * k]; rv = ws.rev[ii + 21 - 3 * k]; }
                        else             { f = ws.fwd[ii + 21 - 3 * k]; rv = ws.rev[ii + 3 * k]; }
                        badF |= (f == 0xFFu); badR |= (rv == 0xFFu);
                        if (FORMAT == 2) { aaF = (aaF << 5) | (f >> 3); aaR = (aaR << 5) | (rv >> 3); }
                        else             { aaF = aaF * 21 + (f >> 3); aaR = aaR * 21 + (rv >> 3); }
                        dnaF = (dnaF << 3) | (f & 7); dnaR = (dnaR << 3) | (rv & 7);
                    }
                    bool okF = active && !badF, okR = active && !badR;
                    if (FORMAT == 2 && smer_len > 0) {                 // syncmer databases: only closed syncmers are queries
                        okF = okF && is_closed_syncmer(aaF, smer_len);
                        okR = okR && is_closed_syncmer(aaR, smer_len);
                    }
                    const uint32_t res = (uint32_t)(x % 3);
                    const uint32_t frameF = res;
                    const uint32_t frameR = 3u + (uint32_t)((lmod - (int)res + 3) % 3);
                    const uint64_t sF = slot_base + (uint64_t)x;
                    const uint64_t sR = slot_base + (uint64_t)npos + (uint64_t)x;
                    const uint32_t pos = (uint32_t)x + pos_off;
                    if (FILTER) {
                        const uint64_t vF = (aaF << 24) | (dnaF & 0xFFFFFFu), vR = (aaR << 24) | (dnaR & 0xFFFFFFu);
                        const bool pF = okF && aa_filter_pass(filter, vF), pR = okR && aa_filter_pass(filter, vR);
                        const uint32_t bF = __ballot_sync(0xffffffffu, pF), bR = __ballot_sync(0xffffffffu, pR);
                        const uint32_t nF = __popc(bF), nR = __popc(bR), need = nF + nR;
                        valid_cnt += (unsigned)okF + (unsigned)okR;
                        if (need) {
                            if (chunk_used + need > kExtractChunk) {          // blank out the rest of the old chunk, take a new one
                                if (chunk_base + kExtractChunk <= out_cap)
                                    for (uint32_t w = chunk_used + lane; w < kExtractChunk; w += 32) { value[chunk_base + w] = kBlank; qinfo[chunk_base + w] = 0ull; }
                                unsigned long long nb = 0;
                                if (lane == 0) nb = atomicAdd(out_cursor, (unsigned long long)kExtractChunk);
                                chunk_base = __shfl_sync(0xffffffffu, nb, 0);
                                chunk_used = 0;
                            }
                            const uint32_t lt = (1u << lane) - 1u;
                            const bool room = chunk_base + kExtractChunk <= out_cap;   // else: capacity guess too small, the host redoes K1
                            if (!room) { chunk_used += need; continue; }
                            if (pF) { const uint64_t s = chunk_base + chunk_used + __popc(bF & lt); value[s] = vF; qinfo[s] = pack_qinfo(r + 1, pos, frameF); }
                            if (pR) { const uint64_t s = chunk_base + chunk_used + nF + __popc(bR & lt); value[s] = vR; qinfo[s] = pack_qinfo(r + 1, pos, frameR); }
                            chunk_used += need;
                            pass_cnt += (lane == 0) ? need : 0u;
                        }
                        continue;
                    }
                    value[sF] = okF ? ((aaF << 24) | (dnaF & 0xFFFFFFu)) : kBlank;
                    qinfo[sF] = okF ? pack_qinfo(r + 1, pos, frameF) : 0ull;
                    value[sR] = okR ? ((aaR << 24) | (dnaR & 0xFFFFFFu)) : kBlank;
                    qinfo[sR] = okR ? pack_qinfo(r + 1, pos, frameR) : 0ull;
                    if (slot_idx) { slot_idx[sF] = (uint32_t)sF; slot_idx[sR] = (uint32_t)sR; }   // sort payload (see k2_sort.cu)
                    valid_cnt += (unsigned)okF + (unsigned)okR;
                }
                __syncwarp();
            }
            slot_base += 6ull * (uint64_t)W;
        }
    }
    for (int o = 16; o > 0; o >>= 1) valid_cnt += __shfl_xor_sync(0xffffffffu, valid_cnt, o);
    if (FILTER) {
        // n_valid[0] = metamers that passed the filter (what the sort and the merge see), n_valid[5] = valid metamers extracted
        if (chunk_used < kExtractChunk && chunk_used > 0 && chunk_base + kExtractChunk <= out_cap)
            for (uint32_t w = chunk_used + lane; w < kExtractChunk; w += 32) { value[chunk_base + w] = kBlank; qinfo[chunk_base + w] = 0ull; }
        if (lane == 0 && pass_cnt) atomicAdd(n_valid, (unsigned long long)pass_cnt);
        if (lane == 0 && valid_cnt) atomicAdd(n_valid + 5, (unsigned long long)valid_cnt);
        return;
    }
    if (lane == 0 && valid_cnt) atomicAdd(n_valid, (unsigned long long)valid_cnt);
}

void launch_read_meta(const uint64_t* off1, const uint64_t* off2, uint32_t n_reads, int32_t* cov1, int32_t* cov2,
                      int32_t* w1, int32_t* w2, uint64_t* slots, uint32_t* quot_cnt, cudaStream_t st) {
    if (!n_reads) return;
    read_meta_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(off1, off2, n_reads, cov1, cov2, w1, w2, slots, quot_cnt);
}

uint64_t extract_filtered_capacity(uint64_t slots, int sm_count) {
    // every warp can leave one chunk partly used, and a chunk switch can strand up to 63 slots
    const uint64_t warps = (uint64_t)sm_count * 64u * kWarpsPerBlock;
    return slots + slots / 7 + warps * kExtractChunk + 4096;
}

void launch_extract(int format, const uint8_t* bases1, const uint64_t* off1, const uint8_t* bases2, const uint64_t* off2,
                    uint32_t n_reads, const int32_t* cov1, const int32_t* w1, const int32_t* w2, const uint64_t* slot_off,
                    const uint8_t* base_code, const uint8_t* codon, uint64_t* value, uint64_t* qinfo, uint32_t* slot_idx,
                    unsigned long long* n_valid, int sm_count, cudaStream_t st, AaFilter filter, unsigned long long* out_cursor,
                    uint64_t out_cap, int smer_len) {
    if (!n_reads) return;
    unsigned blocks = (n_reads + kWarpsPerBlock - 1) / kWarpsPerBlock;
    unsigned cap = (unsigned)sm_count * 64u;          // grid-stride beyond a few waves
    if (blocks > cap) blocks = cap;
    const bool filtered = filter.words != nullptr && out_cursor != nullptr;
#define MBL_LAUNCH_EXTRACT(F, B)                                                                                             \
    extract_kernel<F, B><<<blocks, kWarpsPerBlock * 32, 0, st>>>(bases1, off1, bases2, off2, n_reads, cov1, w1, w2, slot_off, \
                                                                 base_code, codon, value, qinfo, slot_idx, n_valid, filter, out_cursor, out_cap, smer_len)
    if (format == 2) { if (filtered) MBL_LAUNCH_EXTRACT(2, true); else MBL_LAUNCH_EXTRACT(2, false); }
    else             { if (filtered) MBL_LAUNCH_EXTRACT(1, true); else MBL_LAUNCH_EXTRACT(1, false); }
#undef MBL_LAUNCH_EXTRACT
}

}  // namespace mbl
