// Host-side construction of the lookup tables the kernels stage into shared memory.
//   base code   : ASCII -> {A0 C1 T2 G3, 7 = not a base}      (reference common.cpp:13-17 + GeneticCode.h:6)
//   codon       : (c0,c1,c2) -> (aa << 3) | codon_id, 0xFF = no amino acid (GeneticCode.h:32-194)
//   hamming pair: two query codon ids x two target codon ids -> {sum, per-codon 2-bit fields}
//                 (KmerMatcher.h:66-158, 348-416)
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace mbl {

struct HostTables {
    uint8_t base_code[256];
    uint8_t codon[512];
    uint16_t ham_pair[4096];   // idx = (q6 << 6) | t6 ; bits 0-3 sum, 4-7 forward nibble, 8-11 reverse nibble
    uint8_t ham_sum[64];       // single codon: idx = q << 3 | t

    HostTables() {
        // --- base folding: IUPAC and lower case fold onto ACGT the way the reference's atcg[] does
        memset(base_code, 7, sizeof(base_code));
        struct { char from; char to; } fold[] = {{'A', 'A'}, {'B', 'G'}, {'C', 'C'}, {'D', 'G'}, {'G', 'G'},
                                                 {'H', 'T'}, {'K', 'G'}, {'M', 'C'}, {'R', 'A'}, {'S', 'C'},
                                                 {'T', 'T'}, {'U', 'G'}, {'W', 'A'}, {'Y', 'T'}};
        auto code_of = [](char c) -> uint8_t { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : 3; };
        for (auto& f : fold) {
            base_code[(uint8_t)f.from] = code_of(f.to);
            base_code[(uint8_t)(f.from + 32)] = code_of(f.to);
        }
        // --- standard genetic code in the reference's symbol order ARNDCQEGHILKMFPSTWYV + stop(20)
        memset(codon, 0xFF, sizeof(codon));
        const int A = 0, C = 1, T = 2, G = 3;
        int8_t aa[4][4][4];
        auto fam = [&](int a, int b, int v) { for (int c = 0; c < 4; ++c) aa[a][b][c] = (int8_t)v; };
        auto one = [&](int a, int b, int c, int v) { aa[a][b][c] = (int8_t)v; };
        fam(G, C, 0);
        fam(C, G, 1); one(A, G, A, 1); one(A, G, G, 1);
        one(A, A, T, 2); one(A, A, C, 2);
        one(G, A, T, 3); one(G, A, C, 3);
        one(T, G, T, 4); one(T, G, C, 4);
        one(C, A, A, 5); one(C, A, G, 5);
        one(G, A, A, 6); one(G, A, G, 6);
        fam(G, G, 7);
        one(C, A, T, 8); one(C, A, C, 8);
        one(A, T, T, 9); one(A, T, C, 9); one(A, T, A, 9);
        fam(C, T, 10); one(T, T, A, 10); one(T, T, G, 10);
        one(A, A, A, 11); one(A, A, G, 11);
        one(A, T, G, 12);
        one(T, T, T, 13); one(T, T, C, 13);
        fam(C, C, 14);
        fam(T, C, 15); one(A, G, T, 15); one(A, G, C, 15);
        fam(A, C, 16);
        one(T, G, G, 17);
        one(T, A, T, 18); one(T, A, C, 18);
        fam(G, T, 19);
        one(T, A, A, 20); one(T, A, G, 20); one(T, G, A, 20);
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b)
                for (int c = 0; c < 4; ++c) {
                    int id = c;                                   // third base ...
                    if (a == A && b == G) id = (c == G) ? 4 : (c == A) ? 5 : (c == T) ? 6 : 7;   // Arg AGG/AGA, Ser AGT/AGC
                    if (a == T && b == T && c == G) id = 4;       // Leu TTG
                    if (a == T && b == T && c == A) id = 5;       // Leu TTA
                    if (a == T && b == G && c == A) id = 5;       // stop TGA
                    codon[a * 64 + b * 8 + c] = (uint8_t)((aa[a][b][c] << 3) | id);
                }
        // --- synonymous-codon Hamming distances
        static const uint8_t H[8][8] = {{0, 1, 1, 1, 2, 1, 3, 3}, {1, 0, 1, 1, 2, 2, 3, 2}, {1, 1, 0, 1, 2, 2, 2, 3},
                                        {1, 1, 1, 0, 1, 2, 3, 3}, {2, 2, 2, 1, 0, 1, 4, 4}, {1, 2, 2, 2, 1, 0, 4, 4},
                                        {3, 3, 2, 3, 4, 4, 0, 1}, {3, 2, 3, 3, 4, 4, 1, 0}};
        for (int q = 0; q < 8; ++q)
            for (int t = 0; t < 8; ++t) ham_sum[q << 3 | t] = H[q][t];
        for (int q6 = 0; q6 < 64; ++q6)
            for (int t6 = 0; t6 < 64; ++t6) {
                int qlo = q6 & 7, qhi = q6 >> 3, tlo = t6 & 7, thi = t6 >> 3;
                int s = H[qlo][tlo] + H[qhi][thi];
                int flo = H[qlo][tlo] & 3, fhi = H[qhi][thi] & 3;   // the 2-bit tables store distance 4 as 0
                ham_pair[(q6 << 6) | t6] = (uint16_t)(s | ((flo | (fhi << 2)) << 4) | (((flo << 2) | fhi) << 8));
            }
    }
};

}  // namespace mbl
