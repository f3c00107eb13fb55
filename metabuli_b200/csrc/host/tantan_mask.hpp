// Low-complexity masking of query reads (`metabuli classify --mask 1`): the reference runs tantan over every read before the
// k-mers are extracted (KmerExtractor.cpp:308-314 -> SeqIterator::maskLowComplexityRegions, SeqIterator.cpp:154-175 ->
// tantan::maskSequences, lib/mmseqs/lib/tantan) and replaces every letter whose repeat probability reaches --mask-prob by 'N'.
//
// tantan (M. Frith, NAR 2011) is a hidden Markov model with one background state and one "repeat with period k" state for
// every offset k = 1..50; a letter is emitted by repeat state k with the likelihood ratio of pairing it with the letter k
// positions back.  The posterior probability of being in any repeat state comes from one forward and one backward sweep in
// double precision, rescaled every 16 letters.  Parameters as the reference passes them: 50 offsets, repeat start 0.005,
// repeat end 0.05, offset decay 0.9, no gap states.
//
// This is host code (the HMM is a strictly sequential scan per read; reads are spread over threads) restated from the
// published model so that the masked letters agree with the reference build bit for bit: the sums over offsets are
// accumulated in four interleaved lanes and folded (0+2)+(1+3) like the reference's AVX2 code, then finished left to right;
// the multiply-adds the reference build fuses (GCC contracts them under -march=native; read off the disassembly of
// tantan::getProbabilities in oracle/_ref/metabuli) are explicit std::fma here and this file is compiled with
// -ffp-contract=off so that nothing else is; the float conversions before the threshold are the reference's.  Pinned by
// tests/golden/synth/mask_se.tsv.gz (written by the reference binary with --mask 1).
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "tantan_model.hpp"

namespace mblhost {

// seq[0..n) (ASCII) masked in place: letters with repeat probability >= mask_prob, and letters that are not nucleotides, become
// 'N'.  scratch vectors are the caller's (reused across reads of a thread).
inline void tantan_mask_read(const TantanModel& m, char* seq, size_t n, float mask_prob, std::vector<uint8_t>& num, std::vector<float>& prob,
                             std::vector<double>& scale) {
    constexpr int K = TantanModel::kOffsets;
    if (n == 0) return;
    num.resize(n); prob.resize(n); scale.assign(n / 16, 0.0);
    for (size_t t = 0; t < n; ++t) num[t] = m.code[(unsigned char)seq[t]];
    double fg[K];
    // ---- forward
    double bg = 1.0;
    for (int k = 0; k < K; ++k) fg[k] = 0.0;
    for (size_t t = 0; t < n; ++t) {
        const double* row = m.lr[num[t]];
        const int max_off = t < (size_t)K ? (int)t : K;
        double lane[4] = {0.0, 0.0, 0.0, 0.0};
        int k = 0;
        for (; k <= max_off - 4; k += 4)
            for (int u = 0; u < 4; ++u) {
                const double f = fg[k + u];
                lane[u] = lane[u] + f;
                fg[k + u] = std::fma(bg, m.b2f[k + u], f * m.f2f) * row[num[t - (size_t)(k + u) - 1]];
            }
        double from_fg = (lane[0] + lane[2]) + (lane[1] + lane[3]);
        for (; k < max_off; ++k) {
            const double f = fg[k];
            from_fg += f;
            fg[k] = std::fma(bg, m.b2f[k], f * m.f2f) * row[num[t - (size_t)k - 1]];
        }
        bg = std::fma(bg, m.b2b, from_fg * m.f2b);
        if (t % 16 == 15) {
            const double s = 1 / bg;
            scale[t / 16] = s;
            bg *= s;
            for (int j = 0; j < K; ++j) fg[j] *= s;
        }
        prob[t] = static_cast<float>(bg);
    }
    double total = 0.0;
    for (int k = 0; k < K; ++k) total += fg[k];
    const double z = std::fma(total, m.f2b, bg * m.b2b);
    // ---- backward
    bg = m.b2b;
    for (int k = 0; k < K; ++k) fg[k] = m.f2b;
    for (size_t t = n; t-- > 0;) {
        const double non_repeat = prob[t] * bg / z;
        prob[t] = 1 - static_cast<float>(non_repeat);
        if (t % 16 == 15) {
            const double s = scale[t / 16];
            bg *= s;
            for (int j = 0; j < K; ++j) fg[j] *= s;
        }
        const double to_bg = m.f2b * bg;
        const double* row = m.lr[num[t]];
        const int max_off = t < (size_t)K ? (int)t : K;
        double lane[4] = {0.0, 0.0, 0.0, 0.0};
        int k = 0;
        for (; k <= max_off - 4; k += 4)
            for (int u = 0; u < 4; ++u) {
                const double f = fg[k + u] * row[num[t - (size_t)(k + u) - 1]];
                lane[u] = std::fma(m.b2f[k + u], f, lane[u]);
                fg[k + u] = std::fma(f, m.f2f, to_bg);
            }
        double to_fg = (lane[0] + lane[2]) + (lane[1] + lane[3]);
        for (; k < max_off; ++k) {
            const double f = fg[k] * row[num[t - (size_t)k - 1]];
            to_fg = std::fma(m.b2f[k], f, to_fg);
            fg[k] = std::fma(f, m.f2f, to_bg);
        }
        bg = std::fma(bg, m.b2b, to_fg);
    }
    const double min_mask = mask_prob;               // the float --mask-prob promoted, as tantan::maskSequences receives it
    for (size_t t = 0; t < n; ++t)
        if (prob[t] >= min_mask || num[t] == 4) seq[t] = 'N';
}

// every read of an SoA batch, spread over threads
inline void tantan_mask_reads(char* bases, const uint64_t* offsets, size_t n_reads, float mask_prob, unsigned threads) {
    static const TantanModel model;
    unsigned T = threads ? threads : 1;
    if (n_reads < 256) T = 1;
    auto work = [&](size_t r0, size_t r1) {
        std::vector<uint8_t> num;
        std::vector<float> prob;
        std::vector<double> scale;
        for (size_t r = r0; r < r1; ++r) tantan_mask_read(model, bases + offsets[r], (size_t)(offsets[r + 1] - offsets[r]), mask_prob, num, prob, scale);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work, n_reads * t / T, n_reads * (t + 1) / T);
    work(0, n_reads / T);
    for (auto& x : th) x.join();
}

}  // namespace mblhost
