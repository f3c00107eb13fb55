// Parameters of the tantan repeat model as `metabuli classify --mask 1` sets them up (SeqIterator.cpp:154-175 passes 50 offsets,
// repeat start 0.005, repeat end 0.05, offset decay 0.9, no gaps; the likelihood ratios come from lib/mmseqs/data/nucleotide.out
// through SubstitutionMatrix::readProbMatrix and ProbabilityMatrix, BaseMatrix.h:83-97).  Shared by the host masker
// (tantan_mask.hpp) and the device masker (k0_mask.cu); only multiplications, divisions and library calls in here, so the values
// do not depend on the translation unit's floating-point contraction setting.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace mblhost {

struct TantanModel {
    static constexpr int kOffsets = 50;
    double lr[5][5];                 // likelihood ratio of pairing two letters (A C T G X, lib/mmseqs/data/nucleotide.out)
    double b2f[kOffsets];            // background -> repeat state k
    double b2b, f2b, f2f;
    uint8_t code[256];               // NucleotideMatrix::setupLetterMapping (NucleotideMatrix.cpp:18-57)
    TantanModel() {
        // SubstitutionMatrix::readProbMatrix (SubstitutionMatrix.cpp:344-422) on nucleotide.out, then ProbabilityMatrix
        // (BaseMatrix.h:83-97): P_ab = exp(lambda * S_ab) * p_a * p_b, ratio = P_ab / (p_a * p_b) — kept in that order
        const double lambda = strtod("0.6337314", nullptr);
        double pb[5];
        for (int i = 0; i < 4; ++i) pb[i] = strtod("0.2499975", nullptr);
        pb[4] = strtod("0.00001", nullptr);
        for (int i = 0; i < 4; ++i) pb[i] = pb[i] * (1.0 - pb[4]);
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                const double s = (i == j && i < 4) ? strtod("2.0000", nullptr) : strtod("-3.0000", nullptr);
                const double p = std::exp(lambda * s) * pb[i] * pb[j];
                lr[i][j] = p / (pb[i] * pb[j]);
            }
        const double repeat_prob = 0.005, repeat_end = 0.05, decay = 0.9;
        b2b = 1 - repeat_prob;
        f2b = repeat_end;
        f2f = 1 - repeat_end;
        double p = repeat_prob * ((1 - decay) / (1 - std::pow(decay, kOffsets)));
        for (int k = 0; k < kOffsets; ++k) { b2f[k] = p; p *= decay; }
        for (int c = 0; c < 256; ++c) {
            uint8_t v = 4;
            switch (std::toupper(c)) {
                case 'A': v = 0; break;
                case 'C': case 'M': case 'Y': case 'H': v = 1; break;
                case 'T': case 'U': case 'W': v = 2; break;
                case 'G': case 'K': case 'B': case 'D': case 'V': case 'R': case 'S': v = 3; break;
                default: v = 4;
            }
            code[c] = v;
        }
    }
};

}  // namespace mblhost
