// K0 — low-complexity masking of the resident reads (`metabuli classify --mask 1`): tantan's repeat probability per letter
// (KmerExtractor.cpp:308-314 -> SeqIterator::maskLowComplexityRegions, SeqIterator.cpp:154-175 -> tantan::maskSequences,
// lib/mmseqs/lib/tantan/tantan.cpp) computed on the device, in place, before K1 reads the letters.
//
// The model is a forward/backward HMM with one background state and one repeat state per offset 1..50; both sweeps are strictly
// sequential along a read, so the parallelism is ACROSS reads (one warp per read, reads handed out by an atomic counter) and, inside
// a step, across the 50 offsets (lane l owns offsets l and l + 32).  The masking decision is a threshold on a float, so every
// rounding has to be the reference's: products, sums and fused multiply-adds are spelled with the _rn intrinsics in exactly the
// places the reference build rounds (see host/tantan_mask.hpp, which this kernel follows statement by statement), and the sum over
// the offsets keeps the reference's shape — four interleaved chains (lanes 0..3 walk them from a shared-memory copy of the
// offsets' values), folded (0+2)+(1+3), then the tail left to right.
//
// Per warp: two 52-double staging rows (alternating steps, so one __syncwarp per step), a 320-letter window of letter codes, and in
// global scratch one float per letter (the forward background probabilities) plus one double per 16 letters (the rescalings).
//
// tests/host/k0_emulation.cpp compiles THIS file for the CPU (MBL_K0_EMULATION: a 128-thread lock-step emulation of one block with
// the shuffles and barriers the kernel uses) so that the kernel's logic is checked against the goldens in the CPU suite as well.
#include <algorithm>

#include "host/tantan_model.hpp"
#ifndef MBL_K0_EMULATION
#include "kernels.cuh"
#endif

namespace mbl {

namespace {

constexpr int kMaskWarps = 4;          // warps per block
constexpr int kOff = mblhost::TantanModel::kOffsets;
constexpr int kWin = 320;              // letters of the code window
constexpr int kBack = 64;              // letters kept behind the current one when the window moves forward (>= kOff)
static_assert(kOff == 50 && kBack >= kOff && kWin >= 4 * kBack, "window geometry");

struct MaskTables {
    double lr[25];
    double b2f[kOff];
    double b2b, f2b, f2f, min_mask;
    uint8_t code[256];
};

__device__ __forceinline__ double shfl_d(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}
__device__ __forceinline__ double shfl_down_d(double v, int d) {
    return __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(v), d), __shfl_down_sync(0xffffffffu, __double2loint(v), d));
}

__global__ void __launch_bounds__(kMaskWarps * 32)
tantan_mask_kernel(uint8_t* __restrict__ bases, const uint64_t* __restrict__ off, uint32_t n_reads, const __grid_constant__ MaskTables tb,
                   float* __restrict__ prob_scratch, double* __restrict__ scale_scratch, uint64_t stride, unsigned long long* next_read) {
    __shared__ double s_lr[25];
    __shared__ double s_b2f[kOff + 2];
    __shared__ uint8_t s_codetab[256];
    __shared__ double s_f[kMaskWarps][2][kOff + 2];
    __shared__ uint8_t s_code[kMaskWarps][kWin];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 25; i += blockDim.x) s_lr[i] = tb.lr[i];
    for (int i = threadIdx.x; i < kOff; i += blockDim.x) s_b2f[i] = tb.b2f[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_codetab[i] = tb.code[i];
    __syncthreads();
    const double b2b = tb.b2b, f2b = tb.f2b, f2f = tb.f2f, min_mask = tb.min_mask;
    const int k0 = lane, k1 = lane + 32;
    const bool has1 = k1 < kOff;
    const double b2f0 = s_b2f[k0], b2f1 = has1 ? s_b2f[k1] : 0.0;
    const uint64_t gw = (uint64_t)blockIdx.x * kMaskWarps + w;
    float* prob = prob_scratch + gw * stride;
    double* scale = scale_scratch + gw * (stride / 16 + 1);
    uint8_t* code = s_code[w];

    for (;;) {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(next_read, 1ull);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= n_reads) break;
        const uint64_t b0 = off[r];
        const int64_t n = (int64_t)(off[r + 1] - b0);
        if (n <= 0) continue;
        uint8_t* seq = bases + b0;
        int64_t w0 = 0;                                   // s_code[i] = code of letter w0 + i
        auto fill = [&](int64_t start) {
            __syncwarp();
            w0 = start;
            for (int i = lane; i < kWin; i += 32) {
                const int64_t p = start + i;
                code[i] = p < n ? s_codetab[seq[p]] : (uint8_t)4;
            }
            __syncwarp();
        };
        // ---------------- forward ----------------
        fill(0);
        double bg = 1.0, fg0 = 0.0, fg1 = 0.0;
        for (int64_t t = 0; t < n; ++t) {
            if (t >= w0 + kWin) fill(t - kBack);
            const int ct = code[t - w0];
            const int max_off = t < kOff ? (int)t : kOff;
            double* sf = s_f[w][t & 1];
            // the offsets' old values go to shared memory for the sum; the new ones stay in registers
            if (k0 < max_off) {
                sf[k0] = fg0;
                fg0 = __dmul_rn(__fma_rn(bg, b2f0, __dmul_rn(fg0, f2f)), s_lr[ct * 5 + code[t - k0 - 1 - w0]]);
            }
            if (has1 && k1 < max_off) {
                sf[k1] = fg1;
                fg1 = __dmul_rn(__fma_rn(bg, b2f1, __dmul_rn(fg1, f2f)), s_lr[ct * 5 + code[t - k1 - 1 - w0]]);
            }
            __syncwarp();
            const int nq = max_off >> 2;                  // full groups of four offsets
            double s = 0.0;
            if (lane < 4)
                for (int j = 0; j < nq; ++j) s = __dadd_rn(s, sf[4 * j + lane]);
            s = __dadd_rn(s, shfl_down_d(s, 2));          // lane 0: 0+2, lane 1: 1+3
            s = __dadd_rn(s, shfl_down_d(s, 1));          // lane 0: (0+2)+(1+3)
            if (lane == 0) {
                for (int k = 4 * nq; k < max_off; ++k) s = __dadd_rn(s, sf[k]);
                bg = __fma_rn(bg, b2b, __dmul_rn(s, f2b));
            }
            bg = shfl_d(bg, 0);
            if ((t & 15) == 15) {
                const double sc = __ddiv_rn(1.0, bg);
                if (lane == 0) scale[t >> 4] = sc;
                bg = __dmul_rn(bg, sc);
                fg0 = __dmul_rn(fg0, sc);
                fg1 = __dmul_rn(fg1, sc);
            }
            if (lane == 0) prob[t] = __double2float_rn(bg);
        }
        // z = sum of all offsets (left to right) * f2b + bg * b2b
        double z;
        {
            double* sf = s_f[w][0];
            __syncwarp();
            sf[k0] = fg0;
            if (has1) sf[k1] = fg1;
            __syncwarp();
            double total = 0.0;
            if (lane == 0) {
                for (int k = 0; k < kOff; ++k) total = __dadd_rn(total, sf[k]);
                total = __fma_rn(total, f2b, __dmul_rn(bg, b2b));
            }
            z = shfl_d(total, 0);
            __syncwarp();
        }
        // ---------------- backward ----------------
        bg = b2b; fg0 = f2b; fg1 = f2b;
        fill(n > kWin ? n - kWin : 0);
        for (int64_t t = n - 1; t >= 0; --t) {
            if (w0 > 0 && t - kOff < w0) fill(t + 1 > kWin ? t + 1 - kWin : 0);
            const int ct = code[t - w0];
            // this letter's verdict: repeat probability 1 - float(forward background * backward background / z)
            if (lane == 0) {
                const double non_repeat = __ddiv_rn(__dmul_rn((double)prob[t], bg), z);
                const float p = __fsub_rn(1.0f, __double2float_rn(non_repeat));
                if ((double)p >= min_mask || ct == 4) seq[t] = 'N';
            }
            if ((t & 15) == 15) {
                const double sc = scale[t >> 4];
                bg = __dmul_rn(bg, sc);
                fg0 = __dmul_rn(fg0, sc);
                fg1 = __dmul_rn(fg1, sc);
            }
            const double to_bg = __dmul_rn(f2b, bg);
            const int max_off = t < kOff ? (int)t : kOff;
            double* sf = s_f[w][t & 1];
            if (k0 < max_off) {
                const double f = __dmul_rn(fg0, s_lr[ct * 5 + code[t - k0 - 1 - w0]]);
                sf[k0] = f;
                fg0 = __fma_rn(f, f2f, to_bg);
            }
            if (has1 && k1 < max_off) {
                const double f = __dmul_rn(fg1, s_lr[ct * 5 + code[t - k1 - 1 - w0]]);
                sf[k1] = f;
                fg1 = __fma_rn(f, f2f, to_bg);
            }
            __syncwarp();
            const int nq = max_off >> 2;
            double s = 0.0;
            if (lane < 4)
                for (int j = 0; j < nq; ++j) s = __fma_rn(s_b2f[4 * j + lane], sf[4 * j + lane], s);
            s = __dadd_rn(s, shfl_down_d(s, 2));
            s = __dadd_rn(s, shfl_down_d(s, 1));
            if (lane == 0) {
                for (int k = 4 * nq; k < max_off; ++k) s = __fma_rn(s_b2f[k], sf[k], s);
                bg = __fma_rn(bg, b2b, s);
            }
            bg = shfl_d(bg, 0);
        }
        __syncwarp();
    }
}

MaskTables make_tables(float mask_prob) {
    static const mblhost::TantanModel model;
    MaskTables tb;
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) tb.lr[i * 5 + j] = model.lr[i][j];
    for (int k = 0; k < kOff; ++k) tb.b2f[k] = model.b2f[k];
    tb.b2b = model.b2b; tb.f2b = model.f2b; tb.f2f = model.f2f;
    tb.min_mask = (double)mask_prob;            // the float --mask-prob promoted, as tantan::maskSequences receives it
    for (int c = 0; c < 256; ++c) tb.code[c] = model.code[c];
    return tb;
}

}  // namespace

#ifndef MBL_K0_EMULATION
// warps the launch will use and the scratch they need: every warp owns `stride` floats and stride / 16 + 1 doubles
MaskPlan plan_mask(uint32_t n_reads, uint64_t max_len, int sm_count) {
    MaskPlan p{};
    if (!n_reads || !max_len) return p;
    p.stride = (max_len + 15) / 16 * 16;
    const uint64_t per_warp = 4 * p.stride + 8 * (p.stride / 16 + 1);
    uint64_t warps = (uint64_t)sm_count * 64u;                               // 16 blocks of 4 warps per SM
    warps = std::min<uint64_t>(warps, ((uint64_t)n_reads + kMaskWarps - 1) / kMaskWarps * kMaskWarps);
    const uint64_t budget = 2ull << 30;                                      // long reads: fewer warps in flight, 2 GiB of scratch at most
    if (warps * per_warp > budget) warps = std::max<uint64_t>(kMaskWarps, budget / per_warp / kMaskWarps * kMaskWarps);
    p.blocks = (uint32_t)(warps / kMaskWarps);
    p.prob_floats = warps * p.stride;
    p.scale_doubles = warps * (p.stride / 16 + 1);
    return p;
}

void launch_mask(uint8_t* bases, const uint64_t* off, uint32_t n_reads, float mask_prob, const MaskPlan& plan, float* prob_scratch,
                 double* scale_scratch, unsigned long long* counter, cudaStream_t st) {
    if (!n_reads || !plan.blocks) return;
    const MaskTables tb = make_tables(mask_prob);
    cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
    tantan_mask_kernel<<<plan.blocks, kMaskWarps * 32, 0, st>>>(bases, off, n_reads, tb, prob_scratch, scale_scratch, plan.stride, counter);
}

#endif  // MBL_K0_EMULATION

}  // namespace mbl
