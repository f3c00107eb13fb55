// Test infrastructure: runs the SOURCE of the device masking kernel (metabuli_b200/csrc/k0_mask.cu, tantan_mask_kernel) on the CPU.
// One block of the kernel = 128 host threads in lock step: __syncwarp / __syncthreads are pthread barriers, the shuffles go
// through a per-warp exchange row, the _rn intrinsics are the plain IEEE operations (this file is compiled with
// -ffp-contract=off, __fma_rn is std::fma).  It proves the kernel's LOGIC (offset ownership, the four summation chains, the moving
// code window, rescaling, read hand-out) against the reference's goldens where no GPU is present; the GPU tests prove the rest.
//   usage: k0_emulation <mask-prob> < reads (one per line)  > masked reads
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <pthread.h>
#include <string>
#include <vector>

#define MBL_K0_EMULATION 1
#define __device__
#define __forceinline__ inline
#define __global__
#define __shared__ static
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(x)

namespace {
struct Dim { int x; };
thread_local Dim threadIdx;
Dim blockIdx{0}, blockDim{128};
pthread_barrier_t g_block_bar, g_warp_bar[4];
int g_x32[4][32];
unsigned long long g_x64[4][32];
inline int warp_of() { return threadIdx.x >> 5; }
inline int lane_of() { return threadIdx.x & 31; }
inline void warp_wait() { pthread_barrier_wait(&g_warp_bar[warp_of()]); }
}  // namespace

inline void __syncwarp() { warp_wait(); }
inline void __syncthreads() { pthread_barrier_wait(&g_block_bar); }
inline int __shfl_sync(unsigned, int v, int src) {
    g_x32[warp_of()][lane_of()] = v; warp_wait();
    const int r = g_x32[warp_of()][src]; warp_wait();
    return r;
}
inline unsigned long long __shfl_sync(unsigned, unsigned long long v, int src) {
    g_x64[warp_of()][lane_of()] = v; warp_wait();
    const unsigned long long r = g_x64[warp_of()][src]; warp_wait();
    return r;
}
inline int __shfl_down_sync(unsigned, int v, int d) {
    g_x32[warp_of()][lane_of()] = v; warp_wait();
    const int s = lane_of() + d;
    const int r = s < 32 ? g_x32[warp_of()][s] : v; warp_wait();
    return r;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo) { const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __double2float_rn(double a) { return (float)a; }
inline float __fsub_rn(float a, float b) { return a - b; }

#include "../../metabuli_b200/csrc/k0_mask.cu"

namespace {
struct Args {
    uint8_t* bases; const uint64_t* off; uint32_t n; mbl::MaskTables tb; float* prob; double* scale; uint64_t stride;
    unsigned long long* counter; int tid;
};
void* run(void* p) {
    Args* a = (Args*)p;
    threadIdx.x = a->tid;
    mbl::tantan_mask_kernel(a->bases, a->off, a->n, a->tb, a->prob, a->scale, a->stride, a->counter);
    return nullptr;
}
}  // namespace

int main(int argc, char** argv) {
    const float mask_prob = argc > 1 ? (float)atof(argv[1]) : 0.9f;
    std::vector<std::string> seqs;
    std::string line, all;
    while (std::getline(std::cin, line)) seqs.push_back(line);
    std::vector<uint64_t> off{0};
    size_t longest = 0;
    for (auto& s : seqs) { all += s; off.push_back(all.size()); longest = std::max(longest, s.size()); }
    const uint64_t stride = (longest + 15) / 16 * 16;
    const int warps = mbl::kMaskWarps, threads = warps * 32;
    if (warps > 4) { fprintf(stderr, "the emulation holds four warps\n"); return 2; }
    blockDim.x = threads;
    std::vector<float> prob((size_t)warps * stride + 16);
    std::vector<double> scale((size_t)warps * (stride / 16 + 1) + 2);
    unsigned long long counter = 0;
    pthread_barrier_init(&g_block_bar, nullptr, threads);
    for (int w = 0; w < warps; ++w) pthread_barrier_init(&g_warp_bar[w], nullptr, 32);
    std::vector<Args> args(threads);
    std::vector<pthread_t> th(threads);
    const mbl::MaskTables tb = mbl::make_tables(mask_prob);
    for (int t = 0; t < threads; ++t) {
        args[t] = Args{(uint8_t*)&all[0], off.data(), (uint32_t)seqs.size(), tb, prob.data(), scale.data(), stride, &counter, t};
        pthread_create(&th[t], nullptr, run, &args[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], nullptr);
    for (size_t i = 0; i < seqs.size(); ++i) std::cout << all.substr(off[i], off[i + 1] - off[i]) << '\n';
    return 0;
}
