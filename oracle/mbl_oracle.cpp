// ORACLE — TEST INFRASTRUCTURE ONLY (see mbl_oracle.hpp for the rules and the parity pin).
// CPU restatement of the `metabuli classify` hot path.  Paths cited are relative to /root/reference.
#include "mbl_oracle.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// =================================================================================================
// A1  GeneticCode (GeneticCode.h:6,23-29,32-194) and the atcg / iRCT folding tables (common.cpp:13-23)
// =================================================================================================
namespace {

struct CodonTables {
    int8_t aa[8][8][8];    // nuc2aa, -1 when any base is N
    int8_t num[8][8][8];   // nuc2num (3-bit synonymous-codon id)
    uint8_t fold[256];     // atcg[]: IUPAC / lower case -> one of A C G T N '.'
    uint8_t comp[256];     // iRCT[]: complement of a folded base
    CodonTables() {
        memset(aa, -1, sizeof(aa));
        memset(num, -1, sizeof(num));
        // nucleotide code: A0 C1 T2 G3 (nuc2int = (c & 14) >> 1)
        const int A = 0, C = 1, T = 2, G = 3;
        auto set = [&](int a, int b, int c, int v) { aa[a][b][c] = (int8_t)v; };
        auto set4 = [&](int a, int b, int v) { for (int c = 0; c < 4; ++c) aa[a][b][c] = (int8_t)v; };
        // amino-acid order "ARNDCQEGHILKMFPSTWYVX" (X = stop, a valid symbol: Q7)
        set4(G, C, 0);                                           // Ala
        set4(C, G, 1); set(A, G, A, 1); set(A, G, G, 1);         // Arg
        set(A, A, T, 2); set(A, A, C, 2);                        // Asn
        set(G, A, T, 3); set(G, A, C, 3);                        // Asp
        set(T, G, T, 4); set(T, G, C, 4);                        // Cys
        set(C, A, A, 5); set(C, A, G, 5);                        // Gln
        set(G, A, A, 6); set(G, A, G, 6);                        // Glu
        set4(G, G, 7);                                           // Gly
        set(C, A, T, 8); set(C, A, C, 8);                        // His
        set(A, T, T, 9); set(A, T, C, 9); set(A, T, A, 9);       // Ile
        set(T, T, A, 10); set(T, T, G, 10); set4(C, T, 10);      // Leu
        set(A, A, A, 11); set(A, A, G, 11);                      // Lys
        set(A, T, G, 12);                                        // Met
        set(T, T, T, 13); set(T, T, C, 13);                      // Phe
        set4(C, C, 14);                                          // Pro
        set4(T, C, 15); set(A, G, T, 15); set(A, G, C, 15);      // Ser
        set4(A, C, 16);                                          // Thr
        set(T, G, G, 17);                                        // Trp
        set(T, A, T, 18); set(T, A, C, 18);                      // Tyr
        set4(G, T, 19);                                          // Val
        set(T, A, A, 20); set(T, G, A, 20); set(T, A, G, 20);    // stop
        // codon id = third base, with the six-fold / split families remapped (GeneticCode.h:175-193)
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b)
                for (int c = 0; c < 4; ++c) num[a][b][c] = (int8_t)c;
        num[A][G][G] = 4; num[A][G][A] = 5;     // Arg AGG / AGA
        num[T][T][G] = 4; num[T][T][A] = 5;     // Leu TTG / TTA
        num[A][G][T] = 6; num[A][G][C] = 7;     // Ser AGT / AGC
        num[T][G][A] = 5;                       // stop TGA
        // base folding (common.cpp:13-17): everything unknown -> '.', which nuc2int maps to 7 like N
        for (int i = 0; i < 256; ++i) { fold[i] = '.'; comp[i] = '.'; }
        const char *from = "ABCDGHKMNRSTUWY";
        const char *to   = "AGCGGTGCNACTGAT";
        for (int i = 0; from[i]; ++i) {
            fold[(uint8_t)from[i]] = (uint8_t)to[i];
            fold[(uint8_t)(from[i] + 32)] = (uint8_t)(to[i] + 32);
        }
        // complement table (common.cpp:19-23); only folded symbols are ever looked up
        const char *cf = "ABCDGHKMNRSTUVWY";
        const char *ct = "TVGHCDMKNYSAABWR";
        for (int i = 0; cf[i]; ++i) {
            comp[(uint8_t)cf[i]] = (uint8_t)ct[i];
            comp[(uint8_t)(cf[i] + 32)] = (uint8_t)(ct[i] + 32);
        }
    }
};
const CodonTables &tables() { static CodonTables t; return t; }

inline int nuc2int(uint8_t c) { return (c & 14u) >> 1u; }

}  // namespace

// =================================================================================================
// A0  LocalUtil.h:46-60
// =================================================================================================
int max_covered_length(int len) {
    if (len % 3 == 2) return len - 2;
    if (len % 3 == 1) return len - 4;
    return len - 3;
}
int query_kmer_number(int len) { return (max_covered_length(len) / 3 - 8 + 1) * 6; }

// =================================================================================================
// A2 / A3  MetamerScanner / OldMetamerScanner (KmerScanner.h:74-117, 137-181)
// =================================================================================================
namespace {

struct Scanner {
    const CodonTables &t = tables();
    int format;
    const char *seq = nullptr;
    int seqStart = 0, seqEnd = 0, aaLen = 0, posStart = 0, loaded = 0;
    bool fwd = true;
    uint64_t dnaPart = 0, aaPart = 0;
    uint64_t aaQueue[8];  // format 1: contribution of each loaded codon to the base-21 number
    explicit Scanner(int fmt) : format(fmt) {}

    void init(const char *s, int start, int end, bool isForward) {
        seq = s; seqStart = start; seqEnd = end; fwd = isForward;
        aaLen = (end - start + 1) / 3; posStart = 0; loaded = 0; dnaPart = aaPart = 0; nq = 0;
    }
    // codon at reading offset k (0-based within the frame, in reading direction of this scanner)
    inline void codon(int k, int &aa, int &id) const {
        uint8_t b0, b1, b2;
        if (format == 2) {
            if (fwd) {                       // KmerScanner.h:91-93
                int ci = seqStart + k * 3;
                b0 = t.fold[(uint8_t)seq[ci]]; b1 = t.fold[(uint8_t)seq[ci + 1]]; b2 = t.fold[(uint8_t)seq[ci + 2]];
            } else {                         // KmerScanner.h:95-97
                int ci = seqEnd - k * 3;
                b0 = t.comp[t.fold[(uint8_t)seq[ci]]]; b1 = t.comp[t.fold[(uint8_t)seq[ci - 1]]];
                b2 = t.comp[t.fold[(uint8_t)seq[ci - 2]]];
            }
        } else {
            if (fwd) {                       // KmerScanner.h:145-148: forward frames read from the far end
                int ci = seqEnd - k * 3;
                b0 = t.fold[(uint8_t)seq[ci - 2]]; b1 = t.fold[(uint8_t)seq[ci - 1]]; b2 = t.fold[(uint8_t)seq[ci]];
            } else {                         // KmerScanner.h:149-152
                int ci = seqStart + k * 3;
                b0 = t.comp[t.fold[(uint8_t)seq[ci + 2]]]; b1 = t.comp[t.fold[(uint8_t)seq[ci + 1]]];
                b2 = t.comp[t.fold[(uint8_t)seq[ci]]];
            }
        }
        aa = t.aa[nuc2int(b0)][nuc2int(b1)][nuc2int(b2)];
        id = t.num[nuc2int(b0)][nuc2int(b1)][nuc2int(b2)];
    }
    // returns false when the frame is exhausted
    bool next(uint64_t &value, uint32_t &pos) {
        while (posStart <= aaLen - 8) {
            bool sawN = false;
            if (loaded == 8) loaded = 7;     // slide by one codon
            while (loaded < 8) {
                int aa, id;
                codon(posStart + loaded, aa, id);
                if (aa < 0) { sawN = true; break; }
                if (format == 2) {
                    aaPart = (aaPart << 5) | (uint64_t)aa;                 // KmerScanner.h:100-101
                } else {                                                  // KmerScanner.h:155-163
                    // base-21 number of the last 8 codons, newest codon least significant
                    if (nq == 8) { aaPart -= aaQueue[7]; nq = 7; }
                    for (int i = nq; i > 0; --i) aaQueue[i] = aaQueue[i - 1] * 21;
                    aaQueue[0] = (uint64_t)aa; ++nq;
                    aaPart = aaPart * 21 + (uint64_t)aa;
                }
                dnaPart = (dnaPart << 3) | (uint64_t)id;
                ++loaded;
            }
            if (sawN) {                      // KmerScanner.h:104-108
                posStart += loaded + 1; dnaPart = aaPart = 0; loaded = 0; nq = 0;
                continue;
            }
            value = (aaPart << 24) | (dnaPart & 0xFFFFFFull);
            bool leftAnchored = (format == 2) ? fwd : !fwd;
            if (leftAnchored) pos = (uint32_t)(seqStart + posStart * 3);
            else pos = (uint32_t)(seqEnd - (posStart + 8) * 3 + 1);
            ++posStart;
            return true;
        }
        return false;
    }
    int nq = 0;

    // ---- SyncmerScanner (SyncmerScanner.h:9-103): format-2 metamers whose smallest s-mer (s residues, leftmost on ties)
    // sits at the first or the last s-mer position of the 8-residue window ("closed syncmers")
    int smerLen = 0;                     // 0 = every window (MetamerScanner)
    struct SmerAt { uint64_t value; int pos; };
    std::vector<SmerAt> dq;              // monotone deque (front = minimum), SyncmerScanner.h:17
    size_t dqHead = 0;
    int smerCnt = 0, loadedChar = 0, prevPos = -8;
    uint64_t smer = 0;
    void init_syncmer() { dq.clear(); dqHead = 0; smerCnt = 0; loadedChar = 0; prevPos = -8; smer = 0; }
    bool next_syncmer(uint64_t &value, uint32_t &pos) {
        const uint64_t smerMask = (1ull << (5 * smerLen)) - 1;
        bool found = false;
        while (posStart <= aaLen - 8 && !found) {
            bool sawN = false;
            smerCnt -= (smerCnt > 0);
            while (smerCnt < 8 - smerLen + 1) {
                loadedChar -= (loadedChar == smerLen);
                while (loadedChar < smerLen) {
                    int aa, id;
                    codon(posStart + smerCnt + loadedChar, aa, id);
                    if (aa < 0) { sawN = true; break; }
                    smer = (smer << 5) | (uint64_t)aa;
                    ++loadedChar;
                }
                if (sawN) break;
                smer &= smerMask;
                while (dq.size() > dqHead && dq.back().value > smer) dq.pop_back();
                dq.push_back(SmerAt{smer, posStart + smerCnt});
                ++smerCnt;
            }
            if (sawN) {                                           // SyncmerScanner.h:62-69
                posStart += smerCnt + loadedChar + 1;
                prevPos = posStart - 8;
                dq.clear(); dqHead = 0;
                smerCnt = loadedChar = 0;
                smer = 0;
                continue;
            }
            if (dq.size() > dqHead && dq[dqHead].pos < posStart) ++dqHead;
            const int anchor1 = posStart, anchor2 = posStart + (8 - smerLen);
            if (dq.size() > dqHead && (dq[dqHead].pos == anchor1 || dq[dqHead].pos == anchor2)) {
                const int shifts = posStart - prevPos;
                for (int i = 0; i < shifts; ++i) {                // SyncmerScanner.h:75-88
                    int aa, id;
                    codon(prevPos + 8 + i, aa, id);
                    aaPart = (aaPart << 5) | (uint64_t)aa;
                    dnaPart = (dnaPart << 3) | (uint64_t)id;
                }
                prevPos = posStart;
                found = true;
            }
            ++posStart;
        }
        if (!found) return false;
        value = (aaPart << 24) | (dnaPart & 0xFFFFFFull);
        pos = fwd ? (uint32_t)(seqStart + prevPos * 3) : (uint32_t)(seqEnd - (prevPos + 8) * 3 + 1);
        return true;
    }
};

}  // namespace

// A3'  KmerExtractor.cpp:342-373 (fillQueryKmerBuffer), :292-340 (processSequence), :429-481 (loadChunkOfReads)
static void fill_query_kmers(Scanner &sc, const std::string &seq, Kmer *out, size_t &w, uint32_t seqId, uint32_t offset) {
    int L = (int)seq.size();
    int usedLen = max_covered_length(L);
    for (int frame = 0; frame < 6; ++frame) {
        bool fwd = frame < 3;
        int begin = fwd ? frame % 3 : (L % 3) - (frame % 3);
        if (begin < 0) begin += 3;
        sc.init(seq.c_str(), begin, begin + usedLen - 1, fwd);
        uint64_t v; uint32_t p;
        if (sc.smerLen > 0) {
            sc.init_syncmer();
            while (sc.next_syncmer(v, p)) out[w++] = Kmer{v, pack_qinfo(seqId, p + offset, (uint32_t)frame)};
            continue;
        }
        while (sc.next(v, p)) out[w++] = Kmer{v, pack_qinfo(seqId, p + offset, (uint32_t)frame)};
    }
}

void extract_kmers(const std::vector<Read> &m1, const std::vector<Read> *m2, int kmerFormat,
                   std::vector<QueryInfo> &queries, std::vector<Kmer> &kmers, int syncmer, int smerLen) {
    size_t n = m1.size();
    queries.assign(n, QueryInfo());
    std::vector<uint8_t> empty(n, 0);
    size_t total = 0;
    std::vector<size_t> off1(n, 0), off2(n, 0);
    for (size_t i = 0; i < n; ++i) {                      // loadChunkOfReads, forward branch
        int L = (int)m1[i].seq.size();
        queries[i].queryLength = max_covered_length(L);
        int kc = query_kmer_number(L);
        if (kc < 1) { empty[i] = 1; queries[i].kmerCnt = 0; } else queries[i].kmerCnt = kc;
    }
    if (m2) {
        for (size_t i = 0; i < n; ++i) {                  // loadChunkOfReads, isReverse branch
            int L = (int)(*m2)[i].seq.size();
            queries[i].queryLength2 = max_covered_length(L);
            if (empty[i]) continue;
            int kc = query_kmer_number(L);
            if (kc < 1) { empty[i] = 1; queries[i].kmerCnt2 = 0; } else queries[i].kmerCnt2 = kc;
        }
    }
    for (size_t i = 0; i < n; ++i) {
        if (empty[i]) continue;
        off1[i] = total; total += (size_t)queries[i].kmerCnt;
        if (m2) { off2[i] = total; total += (size_t)queries[i].kmerCnt2; }
    }
    kmers.assign(total, Kmer{0, 0});                      // Buffer::init memset (common.h:170-175)
#pragma omp parallel
    {
        Scanner sc(syncmer ? 2 : kmerFormat);            // KmerExtractor.cpp:18-20: SyncmerScanner is a MetamerScanner
        sc.smerLen = syncmer ? smerLen : 0;
#pragma omp for schedule(dynamic, 256)
        for (long long ii = 0; ii < (long long)n; ++ii) {
            size_t i = (size_t)ii;
            if (empty[i]) continue;
            size_t w = off1[i];
            fill_query_kmers(sc, m1[i].seq, kmers.data(), w, (uint32_t)i + 1, 0);
            if (m2) {
                w = off2[i];
                fill_query_kmers(sc, (*m2)[i].seq, kmers.data(), w, (uint32_t)i + 1,
                                 (uint32_t)(queries[i].queryLength + 3));      // KmerExtractor.cpp:322-329
            }
        }
    }
}

// =================================================================================================
// A4  Kmer::compareQueryKmer (Kmer.h:89-94)
// =================================================================================================
// Parallel sort used by the timed CPU baseline: bucket by a monotone key (counting partition, OpenMP), then
// std::sort the buckets in parallel — the role ips4o plays in the reference (FastSort.h:3-20).  The result is the
// same total order as a plain std::sort with `cmp` as long as bucket(a) < bucket(b) implies cmp(a, b).
template <class T, class Cmp, class BucketFn>
static void par_sort(std::vector<T> &v, Cmp cmp, BucketFn bucket, size_t nbuckets, int threads) {
    const size_t n = v.size();
    if (threads <= 1 || n < (1u << 16)) { std::sort(v.begin(), v.end(), cmp); return; }
    std::vector<size_t> cnt((size_t)threads * nbuckets, 0);
    std::vector<T> tmp(n);
#pragma omp parallel num_threads(threads)
    {
        const int t = omp_get_thread_num();
        const size_t b = n * (size_t)t / (size_t)threads, e = n * (size_t)(t + 1) / (size_t)threads;
        size_t *c = cnt.data() + (size_t)t * nbuckets;
        for (size_t i = b; i < e; ++i) ++c[bucket(v[i])];
    }
    std::vector<size_t> start(nbuckets + 1, 0);
    size_t run = 0;
    for (size_t k = 0; k < nbuckets; ++k) {
        start[k] = run;
        for (int t = 0; t < threads; ++t) { size_t c = cnt[(size_t)t * nbuckets + k]; cnt[(size_t)t * nbuckets + k] = run; run += c; }
    }
    start[nbuckets] = run;
#pragma omp parallel num_threads(threads)
    {
        const int t = omp_get_thread_num();
        const size_t b = n * (size_t)t / (size_t)threads, e = n * (size_t)(t + 1) / (size_t)threads;
        size_t *c = cnt.data() + (size_t)t * nbuckets;
        for (size_t i = b; i < e; ++i) tmp[c[bucket(v[i])]++] = v[i];
    }
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
    for (long long k = 0; k < (long long)nbuckets; ++k) std::sort(tmp.begin() + (long)start[k], tmp.begin() + (long)start[k + 1], cmp);
    v.swap(tmp);
}

void sort_kmers(std::vector<Kmer> &kmers, int threads) {
    par_sort(kmers, [](const Kmer &a, const Kmer &b) {
        if (a.value != b.value) return a.value < b.value;
        return qi_seq(a.qinfo) < qi_seq(b.qinfo);
    }, [](const Kmer &k) { return (size_t)(k.value >> 50); }, (size_t)1 << 14, threads);
}

// =================================================================================================
// A6  delta codec (KmerMatcher.h:282-297; IndexCreator.cpp:874-892)
// =================================================================================================
uint64_t next_target_kmer(uint64_t prev, const uint16_t *diff, size_t &idx) {
    uint64_t d = 0;
    uint16_t frag = diff[idx++];
    while (!(frag & 0x8000u)) {
        d |= frag;
        d <<= 15;
        frag = diff[idx++];
    }
    d |= (uint64_t)(frag & 0x7FFFu);
    return prev + d;
}

void encode_delta(uint64_t delta, std::vector<uint16_t> &out) {
    uint16_t buf[5];
    int idx = 3;
    buf[4] = (uint16_t)(0x8000u | (delta & 0x7FFFu));
    delta >>= 15;
    while (delta) { buf[idx--] = (uint16_t)(delta & 0x7FFFu); delta >>= 15; }
    for (int i = idx + 1; i <= 4; ++i) out.push_back(buf[i]);
}

// =================================================================================================
// A8  Hamming tables (KmerMatcher.h:66-158).  HAMMING_LUTk[q<<3|t] == field2[q][t] << 2k except for
// LUT7 rows 4-5 columns 6-7 (Q3), which hold 1<<14 instead of 0.
// =================================================================================================
namespace {
const uint8_t kHamSum[8][8] = {{0, 1, 1, 1, 2, 1, 3, 3}, {1, 0, 1, 1, 2, 2, 3, 2}, {1, 1, 0, 1, 2, 2, 2, 3},
                               {1, 1, 1, 0, 1, 2, 3, 3}, {2, 2, 2, 1, 0, 1, 4, 4}, {1, 2, 2, 2, 1, 0, 4, 4},
                               {3, 3, 2, 3, 4, 4, 0, 1}, {3, 2, 3, 3, 4, 4, 1, 0}};
inline uint16_t ham_field(int k, unsigned q, unsigned t) {
    unsigned v = kHamSum[q][t] & 3u;               // distance 4 (impossible pair) is stored as 0
    if (k == 7 && (q == 4 || q == 5) && (t == 6 || t == 7)) v = 1;   // HAMMING_LUT7 rows 4-5 (Q3)
    return (uint16_t)(v << (2 * k));
}
}  // namespace

uint8_t hamming_sum(uint64_t a, uint64_t b) {
    unsigned s = 0;
    for (int i = 0; i < 8; ++i) s += kHamSum[(a >> (3 * i)) & 7][(b >> (3 * i)) & 7];
    return (uint8_t)s;
}
uint16_t hammings_fwd(uint64_t a, uint64_t b) {
    uint16_t h = 0;
    for (int i = 0; i < 8; ++i) h |= ham_field(i, (a >> (3 * i)) & 7, (b >> (3 * i)) & 7);
    return h;
}
uint16_t hammings_rev(uint64_t a, uint64_t b) {
    uint16_t h = 0;
    for (int i = 0; i < 8; ++i) h |= ham_field(7 - i, (a >> (3 * i)) & 7, (b >> (3 * i)) & 7);
    return h;
}

// =================================================================================================
// A5 / A7 / A8  KmerMatcher::matchKmers (KmerMatcher.cpp:123-481) + compareDna (:1117-1146)
// =================================================================================================
namespace {
constexpr uint64_t kAaMask = ~0xFFFFFFull;
inline uint64_t AA(uint64_t v) { return v & kAaMask; }

struct SliceCtx {
    std::vector<uint64_t> cand;
    std::vector<int32_t> candInfo;
    std::vector<uint8_t> ham;
    std::vector<size_t> sel;
    std::vector<uint8_t> selSum;
    std::vector<uint16_t> selHam;
};

// compareDna (KmerMatcher.cpp:1117-1146)
void compare_dna(uint64_t query, SliceCtx &c, uint32_t frame, int kmerFormat) {
    size_t n = c.cand.size();
    c.ham.resize(n);
    uint8_t minH = 255;
    for (size_t i = 0; i < n; ++i) {
        c.ham[i] = hamming_sum(query, c.cand[i]);
        minH = std::min(minH, c.ham[i]);
    }
    c.sel.clear(); c.selSum.clear(); c.selHam.clear();
    uint8_t maxH = (uint8_t)std::min((int)minH * 2, 7);
    for (size_t h = 0; h < n; ++h) {
        if (c.ham[h] <= maxH) {
            bool plain = !((frame < 3) ^ (kmerFormat == 2));
            c.selSum.push_back(c.ham[h]);
            c.selHam.push_back(plain ? hammings_fwd(query, c.cand[h]) : hammings_rev(query, c.cand[h]));
            c.sel.push_back(h);
        }
    }
}
}  // namespace

bool match_kmers(const Database &db, const std::vector<Kmer> &q, std::vector<Match> &out, int threads, std::string *err) {
    out.clear();
    const size_t numDiff = db.diffIdx.size();
    const uint16_t *diff = db.diffIdx.data();
    // blanks sort first (KmerMatcher.cpp:143-151)
    size_t blank = 0;
    while (blank < q.size() && qi_seq(q[blank].qinfo) == 0) ++blank;
    size_t nq = q.size() - blank;
    if (nq == 0) return true;
    // usable checkpoints (KmerMatcher.cpp:156-164)
    std::vector<Split> sp(db.split);
    size_t use = sp.size();
    for (size_t i = 1; i < sp.size(); ++i)
        if (sp[i].adKmer == 0 || sp[i].adKmer == UINT64_MAX) { sp[i] = Split{UINT64_MAX, UINT64_MAX, UINT64_MAX}; --use; }
    if (threads < 1) threads = 1;
    if ((size_t)threads > nq) threads = (int)nq;
    // slices (KmerMatcher.cpp:166-194)
    struct Slice { size_t start, end; Split cp; };
    std::vector<Slice> slices;
    size_t quo = nq / (size_t)threads, rem = nq % (size_t)threads, s = blank;
    for (int i = 0; i < threads; ++i) {
        size_t e = s + quo - 1;
        if (rem > 0) { ++e; --rem; }
        uint64_t qa = AA(q[s].value);
        bool needLast = true;
        for (size_t j = 0; j < use; ++j) {
            if (qa <= AA(sp[j].adKmer)) {
                j = j - (j != 0);
                slices.push_back(Slice{s, e, sp[j]});
                needLast = false;
                break;
            }
        }
        if (needLast) slices.push_back(Slice{s, e, sp[use >= 2 ? use - 2 : 0]});
        s = e + 1;
    }
    const uint32_t mask = ~((uint32_t)(db.params.skipRedundancy == 0) << 31);   // :204-205
    std::vector<std::vector<Match>> parts(slices.size());
    bool bad = false;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (long long si = 0; si < (long long)slices.size(); ++si) {
        const Slice &sl = slices[(size_t)si];
        std::vector<Match> &dst = parts[(size_t)si];
        SliceCtx c;
        uint64_t curT = sl.cp.adKmer;
        size_t dpos = sl.cp.diffIdxOffset;                                  // == diffIdxPos
        size_t ipos = sl.cp.infoIdxOffset - (sl.cp.adKmer != 0);            // :259-260
        if (sl.cp.adKmer == 0 && sl.cp.diffIdxOffset == 0 && sl.cp.infoIdxOffset == 0)
            curT = next_target_kmer(curT, diff, dpos);                      // :267-271
        uint64_t curQ = UINT64_MAX, curQAA = UINT64_MAX;
        uint32_t curFrame = 0;
        auto emit = [&](const Kmer &k) {
            for (size_t x = 0; x < c.sel.size(); ++x) {
                size_t idx = c.sel[x];
                int32_t tid = c.candInfo[idx];
                int32_t spc = (tid >= 0 && (size_t)tid < db.taxid2species.size()) ? db.taxid2species[tid] : 0;
                if (tid == 0 || spc == 0) { bad = true; return; }           // Q2 (:292-300)
                dst.push_back(Match{k.qinfo, tid, spc, (uint32_t)(c.cand[idx] & 0xFFFFFFu), c.selHam[x], c.selSum[x], 0});
            }
        };
        for (size_t j = sl.start; j <= sl.end && !bad; ++j) {
            const Kmer &k = q[j];
            uint32_t frame = qi_frame(k.qinfo);
            if (curQ == k.value && curFrame / 3 == frame / 3) { emit(k); continue; }      // :277-311
            if (curQAA == AA(k.value)) {                                                   // :315-353
                compare_dna(k.value, c, frame, db.params.kmerFormat);
                emit(k);
                curQ = k.value; curQAA = AA(curQ); curFrame = frame;
                continue;
            }
            c.cand.clear(); c.candInfo.clear(); c.sel.clear(); c.selSum.clear(); c.selHam.clear();
            curQ = k.value; curQAA = AA(curQ); curFrame = frame;
            while (dpos != numDiff && curQAA > AA(curT)) {                                 // :363-371
                curT = next_target_kmer(curT, diff, dpos);
                ++ipos;
            }
            if (curQAA != AA(curT)) continue;                                              // :373-375
            while (dpos != numDiff && curQAA == AA(curT)) {                                // :378-406 (Q1)
                c.cand.push_back(curT);
                c.candInfo.push_back((int32_t)((uint32_t)db.info[ipos] & mask));
                curT = next_target_kmer(curT, diff, dpos);
                ++ipos;
            }
            compare_dna(curQ, c, frame, db.params.kmerFormat);
            emit(k);
        }
    }
    if (bad) { if (err) *err = "target k-mer with taxid 0 or unmapped species (reference exits)"; return false; }
    for (auto &p : parts) out.insert(out.end(), p.begin(), p.end());
    return true;
}

// =================================================================================================
// A9  KmerMatcher::compareMatches (KmerMatcher.cpp:1149-1166)
// =================================================================================================
static bool compare_matches(const Match &a, const Match &b) {
    if (qi_seq(a.qinfo) != qi_seq(b.qinfo)) return qi_seq(a.qinfo) < qi_seq(b.qinfo);
    if (a.speciesId != b.speciesId) return a.speciesId < b.speciesId;
    if (qi_frame(a.qinfo) != qi_frame(b.qinfo)) return qi_frame(a.qinfo) < qi_frame(b.qinfo);
    if (qi_pos(a.qinfo) != qi_pos(b.qinfo)) return qi_pos(a.qinfo) < qi_pos(b.qinfo);
    if (a.hamming != b.hamming) return a.hamming < b.hamming;
    return a.dnaEncoding < b.dnaEncoding;
}
void sort_matches(std::vector<Match> &m, int threads) {
    uint32_t maxSeq = 0;
    for (const Match &x : m) maxSeq = std::max(maxSeq, qi_seq(x.qinfo));
    const size_t nb = 1 << 14;
    const uint64_t div = (uint64_t)maxSeq / nb + 1;
    par_sort(m, compare_matches, [div](const Match &x) { return (size_t)(qi_seq(x.qinfo) / div); }, nb, threads);
}

// =================================================================================================
// A8''  taxonomy (TaxonomyWrapper.cpp:363-421, NcbiTaxonomy.cpp:250-330, 415-432)
// =================================================================================================
int rank_index(const char *rank) {
    static const std::map<std::string, int> ranks = {
        {"forma", 1}, {"varietas", 2}, {"subspecies", 3}, {"species", 4}, {"species subgroup", 5},
        {"species group", 6}, {"subgenus", 7}, {"genus", 8}, {"subtribe", 9}, {"tribe", 10},
        {"subfamily", 11}, {"family", 12}, {"superfamily", 13}, {"parvorder", 14}, {"infraorder", 15},
        {"suborder", 16}, {"order", 17}, {"superorder", 18}, {"infraclass", 19}, {"subclass", 20},
        {"class", 21}, {"superclass", 22}, {"subphylum", 23}, {"phylum", 24}, {"superphylum", 25},
        {"subkingdom", 26}, {"kingdom", 27}, {"superkingdom", 28}, {"domain", 28}};
    auto it = ranks.find(rank);
    return it == ranks.end() ? -1 : it->second;
}

static int flog2(size_t v) { int r = 0; while (v >>= 1) ++r; return r; }

bool Taxonomy::load(const std::string &path, std::string *err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::vector<char> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return load_blob(b.data(), b.size(), err);
}

bool Taxonomy::load_blob(const char *data, size_t size, std::string *err) {
    blob.assign(data, data + size);
    const char *p = blob.data();
    int32_t version; memcpy(&version, p, 4); p += 4;
    if (version != 2) { if (err) *err = "unsupported taxonomyDB version"; return false; }
    uint64_t flag; memcpy(&flag, p, 8);
    internalIds = (flag == 1);                       // TaxonomyWrapper.cpp:374-382 (detected by value)
    if (internalIds) p += 8;
    uint64_t mn; memcpy(&mn, p, 8); p += 8; maxNodes = (size_t)mn;
    memcpy(&maxTaxID, p, 4); p += 4;
    nodeTaxId.resize(maxNodes); nodeParent.resize(maxNodes); nodeRankIdx.resize(maxNodes); nodeNameIdx.resize(maxNodes);
    for (size_t i = 0; i < maxNodes; ++i) {          // TaxonNode: i32 id, i32 taxId, i32 parent, pad, u64 rankIdx, u64 nameIdx
        const char *n = p + 32 * i;
        memcpy(&nodeTaxId[i], n + 4, 4); memcpy(&nodeParent[i], n + 8, 4);
        memcpy(&nodeRankIdx[i], n + 16, 8); memcpy(&nodeNameIdx[i], n + 24, 8);
    }
    p += 32 * maxNodes;
    D = (const int32_t *)p; p += 4 * ((size_t)maxTaxID + 1);
    internal2org = (const int32_t *)p;
    if (internalIds) p += 4 * ((size_t)maxTaxID + 1);
    E = (const int32_t *)p; p += 4 * 2 * maxNodes;
    L = (const int32_t *)p; p += 4 * 2 * maxNodes;
    H = (const int32_t *)p; p += 4 * maxNodes;
    Mk = flog2(2 * maxNodes) + 1;
    M = (const int32_t *)p; p += 4 * 2 * maxNodes * (size_t)Mk;
    uint64_t byteCap; uint32_t entryCap, entryCnt;   // StringBlock.h:112-145
    memcpy(&byteCap, p, 8); p += 8; memcpy(&entryCap, p, 4); p += 4; memcpy(&entryCnt, p, 4); p += 4;
    strData = p; p += byteCap;
    strOffsets = (const uint32_t *)p; p += 4 * (size_t)entryCap;
    strCount = entryCnt;
    if ((size_t)(p - blob.data()) > blob.size()) { if (err) *err = "taxonomyDB truncated"; return false; }
    eukaryota = 0;                                   // TaxonomyWrapper.h setEukaryoteTaxID
    for (size_t i = 0; i < maxNodes; ++i) {
        if (nodeNameIdx[i] == 0) continue;
        if (strcmp(str(nodeNameIdx[i]), "Eukaryota") == 0) { eukaryota = nodeTaxId[i]; break; }
    }
    return true;
}

int Taxonomy::lcaHelper(int i, int j) const {       // NcbiTaxonomy.cpp:250-281 (Q5: node 0 short-circuit)
    if (i == 0 || j == 0) return 0;
    if (i == j) return i;
    int v1 = H[i], v2 = H[j];
    if (v1 > v2) std::swap(v1, v2);
    int k = flog2((size_t)(v2 - v1 + 1));
    int a = M[(size_t)v1 * Mk + k];
    int b = M[(size_t)(v2 - (1 << k) + 1) * Mk + k];
    int rmq = (L[a] <= L[b]) ? a : b;
    return E[rmq];
}
int32_t Taxonomy::lca(int32_t a, int32_t b) const {  // NcbiTaxonomy.cpp:300-307
    if (!nodeExists(a)) return b;
    if (!nodeExists(b)) return a;
    return nodeTaxId[lcaHelper(nodeId(a), nodeId(b))];
}
int32_t Taxonomy::lcaMany(const std::vector<int32_t> &taxa) const {   // NcbiTaxonomy.cpp:310-330
    size_t i = 0;
    while (i < taxa.size() && !nodeExists(taxa[i])) ++i;
    if (i == taxa.size()) return 0;
    int red = nodeId(taxa[i++]);
    for (; i < taxa.size(); ++i)
        if (nodeExists(taxa[i])) red = lcaHelper(red, nodeId(taxa[i]));
    return nodeTaxId[red];
}
bool Taxonomy::isAncestor(int32_t ancestor, int32_t child) const {     // NcbiTaxonomy.cpp:282-298
    if (ancestor == child) return true;
    if (ancestor == 0 || child == 0) return false;
    if (!nodeExists(child) || !nodeExists(ancestor)) return false;
    return lcaHelper(nodeId(child), nodeId(ancestor)) == nodeId(ancestor);
}
int32_t Taxonomy::taxIdAtRank(int32_t taxId, const char *rank) const { // TaxonomyWrapper.cpp:479-498
    if (taxId == 0 || !nodeExists(taxId) || taxId == 1) return 0;
    int want = rank_index(rank);
    int node = nodeId(taxId);
    int cnt = 0;
    while (cnt < 30 && rank_index(str(nodeRankIdx[node])) < want) {
        node = nodeId(nodeParent[node]);
        ++cnt;
    }
    if (cnt == 30) return taxId;
    return nodeTaxId[node];
}

// A8'  KmerMatcher::loadTaxIdList (KmerMatcher.cpp:96-119)
void build_taxid2species(const Taxonomy &tax, const std::vector<int32_t> &list, std::vector<int32_t> &out) {
    out.assign((size_t)tax.maxTaxID + 1, 0);
    for (int32_t taxId : list) {
        int32_t species = tax.taxIdAtRank(taxId, "species");
        int node = tax.nodeId(taxId);
        if (taxId != tax.nodeTaxId[node]) out[taxId] = species;
        while (tax.nodeTaxId[node] != species) {
            out[tax.nodeTaxId[node]] = species;
            node = tax.nodeId(tax.nodeParent[node]);
        }
        out[species] = species;
    }
}

// =================================================================================================
// DB files
// =================================================================================================
template <class T>
static bool slurp(const std::string &path, std::vector<T> &v, std::string *err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize((size_t)sz / sizeof(T));
    size_t got = v.empty() ? 0 : fread(v.data(), sizeof(T), v.size(), f);
    fclose(f);
    if (got != v.size()) { if (err) *err = "short read " + path; return false; }
    return true;
}

bool Database::load(const std::string &dir, std::string *err) {
    std::ifstream pf(dir + "/db.parameters");                 // common.cpp:88-133
    std::string line;
    while (std::getline(pf, line)) {
        size_t tab = line.find('\t');
        if (tab == std::string::npos) continue;
        std::string k = line.substr(0, tab), v = line.substr(tab + 1);
        if (k == "Reduced_alphabet") params.reducedAA = atoi(v.c_str());
        else if (k == "Skip_redundancy") { if (v == "1") params.skipRedundancy = 1; }
        else if (k == "Syncmer") { if (v == "1") params.syncmer = 1; }
        else if (k == "S-mer_len") params.smerLen = atoi(v.c_str());
        else if (k == "Kmer_format") params.kmerFormat = atoi(v.c_str());
        else if (k == "Accession_level") params.accessionLevelDb = atoi(v.c_str());
    }
    if (!slurp(dir + "/diffIdx", diffIdx, err)) return false;
    if (!slurp(dir + "/info", info, err)) return false;
    if (!slurp(dir + "/split", split, err)) return false;
    if (!tax.load(dir + "/taxonomyDB", err)) return false;
    std::ifstream tl(dir + "/taxID_list");
    if (!tl) { if (err) *err = "cannot open taxID_list"; return false; }
    std::vector<int32_t> list;
    while (std::getline(tl, line)) if (!line.empty()) list.push_back((int32_t)strtoul(line.c_str(), nullptr, 10));
    build_taxid2species(tax, list, taxid2species);
    return true;
}

// =================================================================================================
// FASTA / FASTQ reader with kseq semantics (name = first whitespace-delimited token; multi-line records)
// =================================================================================================
bool read_fastx(const std::string &path, std::vector<Read> &out, std::string *err) {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) { if (err) *err = "cannot open " + path; return false; }
    gzbuffer(g, 1 << 20);
    std::string data;
    std::vector<char> buf(1 << 20);
    int got;
    while ((got = gzread(g, buf.data(), (unsigned)buf.size())) > 0) data.append(buf.data(), (size_t)got);
    gzclose(g);
    size_t i = 0, n = data.size();
    auto getline_ = [&](size_t &b, size_t &e) {      // [b,e) without the newline / trailing CR
        b = i;
        while (i < n && data[i] != '\n') ++i;
        e = i;
        if (i < n) ++i;
        if (e > b && data[e - 1] == '\r') --e;
        return b < n;
    };
    size_t b, e;
    while (i < n) {
        while (i < n && data[i] != '>' && data[i] != '@') { getline_(b, e); }   // skip to a header
        if (i >= n) break;
        char tag = data[i];
        getline_(b, e);
        Read r;
        size_t p = b + 1;
        while (p < e && !isspace((unsigned char)data[p])) ++p;
        r.name.assign(data, b + 1, p - (b + 1));
        while (i < n && data[i] != '>' && data[i] != '@' && data[i] != '+') {
            getline_(b, e);
            for (size_t x = b; x < e; ++x) if (isgraph((unsigned char)data[x])) r.seq.push_back(data[x]);
        }
        if (tag == '@' && i < n && data[i] == '+') {
            getline_(b, e);
            size_t ql = 0;
            while (i < n && ql < r.seq.size()) { getline_(b, e); ql += e - b; }
        }
        out.push_back(std::move(r));
    }
    return true;
}

// =================================================================================================
// A10-A12  Taxonomer (Taxonomer.cpp)
// =================================================================================================
namespace {

struct Path {             // Taxonomer.h:33-52 MatchPath
    int start = 0, end = 0;
    float score = 0.f;
    int hammingDist = 0, depth = 0;
    const Match *startMatch = nullptr, *endMatch = nullptr;
};

inline float codon_score(int d) { return d == 0 ? 3.0f : 2.0f - 0.5f * (float)d; }
float match_score(const Match &m) {                                    // Match.h:32-44
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += codon_score((m.rightEndHamming >> (2 * i)) & 3);
    return s;
}
float right_part_score(const Match &m, int range) {                    // Match.h:46-57
    float s = 0.f;
    for (int i = 0; i < range; ++i) s += codon_score((m.rightEndHamming >> (2 * i)) & 3);
    return s;
}
float left_part_score(const Match &m, int range) {                     // Match.h:59-70
    float s = 0.f;
    for (int i = 0; i < range; ++i) s += codon_score((m.rightEndHamming >> (14 - 2 * i)) & 3);
    return s;
}
int right_part_ham(const Match &m, int range) { int s = 0; for (int i = 0; i < range; ++i) s += (m.rightEndHamming >> (2 * i)) & 3; return s; }
int left_part_ham(const Match &m, int range) { int s = 0; for (int i = 0; i < range; ++i) s += (m.rightEndHamming >> (14 - 2 * i)) & 3; return s; }

struct Scorer {
    const Database &db;
    const Options &opt;
    int kmerFormat, denominator;
    int maxCodonShift = 1, dnaShift = 3;                               // Taxonomer.cpp:34-42
    std::vector<Path> paths, combined, local;
    std::vector<uint8_t> connected;

    Scorer(const Database &d, const Options &o) : db(d), opt(o) {
        kmerFormat = d.params.kmerFormat;
        if (d.params.syncmer) { dnaShift = (8 - d.params.smerLen) * 3; maxCodonShift = 8 - d.params.smerLen; }
        denominator = (o.seqMode == 1 || o.seqMode == 2) ? 100 : 1000;  // Taxonomer.cpp:44-48
    }

    // isConsecutive / isConsecutive2 with shift (Taxonomer.cpp:677-699)
    bool consecutive(const Match *cur, const Match *next, bool forwardFrame, int shift) const {
        const Match *m1 = forwardFrame ? cur : next;
        const Match *m2 = forwardFrame ? next : cur;
        uint32_t lowMask = (1u << (24 - 3 * shift)) - 1;
        if (kmerFormat == 2) return (m1->dnaEncoding & lowMask) == (m2->dnaEncoding >> (3 * shift));
        return (m1->dnaEncoding >> (3 * shift)) == (m2->dnaEncoding & lowMask);
    }

    // getMatchPaths (Taxonomer.cpp:487-648); [start,end) is one (species, frame) group
    void match_paths(const Match *ml, size_t start, size_t end, int32_t species) {
        size_t i = start;
        size_t currPos = qi_pos(ml[start].qinfo);
        bool forwardFrame = qi_frame(ml[start].qinfo) < 3;
        int minDepth = opt.minConsCnt;
        if (db.tax.isAncestor(db.tax.eukaryota, species)) minDepth = opt.minConsCntEuk;
        connected.assign(end - start + 1, 0);
        local.assign(end - start + 1, Path());
        auto initPath = [&](size_t idx) {                              // MatchPath(const Match*) Taxonomer.h:39-46
            Path p;
            p.start = (int)qi_pos(ml[idx].qinfo); p.end = p.start + 23;
            p.score = match_score(ml[idx]); p.hammingDist = ml[idx].hamming; p.depth = 1;
            p.startMatch = p.endMatch = ml + idx;
            local[idx - start] = p;
        };
        size_t curS = i;
        while (i < end && qi_pos(ml[i].qinfo) == currPos) { initPath(i); ++i; }
        size_t curE = i;
        while (i < end) {
            uint32_t nextPos = qi_pos(ml[i].qinfo);
            size_t nxtS = i;
            while (i < end && nextPos == qi_pos(ml[i].qinfo)) { initPath(i); ++i; }
            size_t nxtE = i;
            int shift = (int)(((size_t)nextPos - currPos) / 3);
            if (shift > 0 && shift <= maxCodonShift) {
                for (size_t nx = nxtS; nx < nxtE; ++nx) {
                    float inc = 0.f; int hinc = 0;                     // calScoreIncrement / calHammingDistIncrement
                    for (int s = 0; s < shift; ++s) {
                        int h = (ml[nx].rightEndHamming >> (2 * s)) & 3;
                        inc += codon_score(h); hinc += h;
                    }
                    const Path *best = nullptr; float bestScore = 0.f;
                    for (size_t cu = curS; cu < curE; ++cu) {
                        if (consecutive(ml + cu, ml + nx, forwardFrame, shift)) {
                            connected[cu - start] = 1;
                            if (local[cu - start].score > bestScore) { best = &local[cu - start]; bestScore = best->score; }
                        }
                    }
                    if (best) {
                        Path &p = local[nx - start];
                        p.start = best->start; p.score = best->score + inc;
                        p.hammingDist = best->hammingDist + hinc; p.depth = best->depth + shift;
                        p.startMatch = best->startMatch;
                    }
                }
            }
            for (size_t cu = curS; cu < curE; ++cu)
                if (!connected[cu - start] && local[cu - start].depth >= minDepth) paths.push_back(local[cu - start]);
            if (i == end)
                for (size_t nx = nxtS; nx < nxtE; ++nx)
                    if (local[nx - start].depth >= minDepth) paths.push_back(local[nx - start]);
            curS = nxtS; curE = nxtE; currPos = nextPos;
        }
    }

    // trimMatchPath (Taxonomer.cpp:475-485)
    static void trim(Path &p1, const Path &p2, int overlap) {
        if (p1.start < p2.start) {
            p1.end = p2.start - 1;
            p1.hammingDist = std::max(0, p1.hammingDist - right_part_ham(*p1.endMatch, overlap / 3));
            p1.score = p1.score - right_part_score(*p1.endMatch, overlap / 3) - (overlap % 3);
        } else {
            p1.start = p2.end + 1;
            p1.hammingDist = std::max(0, p1.hammingDist - left_part_ham(*p1.startMatch, overlap / 3));
            p1.score = p1.score - left_part_score(*p1.startMatch, overlap / 3) - (overlap % 3);
        }
    }

    // combineMatchPaths (Taxonomer.cpp:410-468).  std::sort on purpose: Q4.
    float combine(size_t pathStart, size_t combStart, int readLength) {
        std::sort(paths.begin() + (long)pathStart, paths.end(), [](const Path &a, const Path &b) {
            if (a.score != b.score) return a.score > b.score;
            if (a.hammingDist != b.hammingDist) return a.hammingDist < b.hammingDist;
            return a.start > b.start;
        });
        float score = 0;
        for (size_t i = pathStart; i < paths.size(); ++i) {
            if (combStart == combined.size()) {
                combined.push_back(paths[i]);
                score += paths[i].score;
                continue;
            }
            bool overlapped = false;
            for (size_t j = combStart; j < combined.size(); ++j) {
                Path &p = paths[i];
                const Path &c = combined[j];
                if (!((p.end < c.start) || (c.end < p.start))) {
                    int ov = std::min(p.end, c.end) - std::max(p.start, c.start) + 1;
                    if (ov == p.end - p.start + 1) { overlapped = true; break; }
                    if (ov < 24) { trim(p, c, ov); continue; }
                    overlapped = true; break;
                }
            }
            if (!overlapped) { combined.push_back(paths[i]); score += paths[i].score; }
        }
        return score / readLength;
    }

    // chooseBestTaxon (Taxonomer.cpp:130-202) with getBestSpeciesMatches (:316-408) inlined
    void score_read(const Match *ml, size_t offset, size_t end /*inclusive*/, QueryInfo &q) {
        paths.clear(); combined.clear();
        std::vector<std::pair<int32_t, float>> sp2score;
        int queryLength = q.queryLength + q.queryLength2;
        float bestSpScore = 0;
        size_t meaningful = 0;
        std::pair<size_t, size_t> bestRange(0, 0);
        size_t i = offset;
        while (i < end + 1) {
            int32_t species = ml[i].speciesId;
            size_t spStart = i;
            size_t prev = paths.size();
            while (i < end + 1 && species == ml[i].speciesId) {
                uint32_t frame = qi_frame(ml[i].qinfo);
                size_t fs = i;
                while (i < end + 1 && species == ml[i].speciesId && frame == qi_frame(ml[i].qinfo)) ++i;
                if (i - fs > 1) match_paths(ml, fs, i, species);       // Q9
            }
            if (paths.size() > prev) {
                float score = combine(prev, combined.size(), queryLength);
                score = std::min(score, 1.0f);
                if (score < opt.minScore) continue;
                sp2score.emplace_back(species, score);
                if (score > 0.f) ++meaningful;
                if (score > bestSpScore) { bestSpScore = score; bestRange = {spStart, i}; }
            }
        }
        float finalScore = 0.f; int32_t taxId = 0; bool isLca = false;
        if (meaningful != 0) {
            std::vector<int32_t> maxSpecies;
            for (auto &s : sp2score)
                if (s.second >= bestSpScore * opt.tieRatio) { maxSpecies.push_back(s.first); finalScore += s.second; }
            if (maxSpecies.size() > 1) {
                isLca = true;
                taxId = db.tax.lcaMany(maxSpecies);
                finalScore /= maxSpecies.size();
            } else {
                taxId = maxSpecies[0];
            }
        }
        // --- chooseBestTaxon
        if (finalScore == 0 || finalScore < opt.minScore) {
            q.isClassified = false; q.classification = 0; q.score = finalScore; q.hammingDist = 0;   // Q6
            return;
        }
        if (isLca) { q.isClassified = true; q.classification = taxId; q.score = finalScore; q.hammingDist = 0; return; }
        // filterRedundantMatches (Taxonomer.cpp:205-241)
        size_t maxQuot = (size_t)(queryLength + 3) / (size_t)dnaShift;
        std::vector<const Match *> bestM(maxQuot + 1, nullptr);
        std::vector<int32_t> bestTax(maxQuot + 1, 0);
        std::vector<uint8_t> minHam(maxQuot + 1, 255);
        for (size_t k = bestRange.first; k < bestRange.second; ++k) {
            size_t quo = qi_pos(ml[k].qinfo) / (size_t)dnaShift;
            uint8_t h = ml[k].hamming;
            if (bestM[quo] == nullptr || h < minHam[quo]) { bestM[quo] = ml + k; bestTax[quo] = ml[k].targetId; minHam[quo] = h; }
            else if (h == minHam[quo]) bestTax[quo] = db.tax.lca(bestTax[quo], ml[k].targetId);
        }
        std::unordered_map<int32_t, unsigned> taxCnt;
        for (size_t k = 0; k <= maxQuot; ++k) if (bestM[k]) taxCnt[bestTax[k]]++;
        for (auto &t : taxCnt) q.taxCnt[t.first] = (int)t.second;
        if (finalScore < opt.minSpScore) {                                // Taxonomer.cpp:172-180
            q.isClassified = true;
            q.classification = db.tax.parentOf(db.tax.taxIdAtRank(taxId, "species"));
            q.score = finalScore; q.hammingDist = 0;
            return;
        }
        q.isClassified = true; q.score = finalScore; q.hammingDist = 0;
        q.classification = lower_rank(taxCnt, taxId, queryLength);
    }

    // lowerRankClassification / getSpeciesCladeCounts / BFS (Taxonomer.cpp:252-314)
    int32_t lower_rank(const std::unordered_map<int32_t, unsigned> &taxCnt, int32_t species, int queryLength) {
        unsigned minSub = (unsigned)((queryLength - 1) / denominator);
        struct Clade { unsigned taxCount = 0, cladeCount = 0; std::vector<int32_t> children; };
        std::unordered_map<int32_t, Clade> clade;
        for (auto &t : taxCnt) {
            int32_t cur = t.first;
            clade[cur].taxCount = t.second;
            clade[cur].cladeCount += t.second;
            while (cur != species) {
                int32_t par = db.tax.parentOf(cur);
                auto &ch = clade[par].children;
                if (std::find(ch.begin(), ch.end(), cur) == ch.end()) ch.push_back(cur);
                clade[par].cladeCount += t.second;
                cur = par;
            }
        }
        if (opt.accessionLevel == 2) {                                    // Taxonomer.cpp:256-267
            std::vector<int32_t> keys;
            for (auto &c : clade) keys.push_back(c.first);
            for (int32_t k : keys) {
                const char *rk = db.tax.rankOf(k);
                if (strcmp(rk, "") == 0 || strcmp(rk, "accession") == 0) {
                    auto &ch = clade[db.tax.parentOf(k)].children;
                    auto it = std::find(ch.begin(), ch.end(), k);
                    if (it != ch.end()) ch.erase(it);
                }
            }
        }
        int32_t root = species;
        while (true) {                                                    // BFS (tail recursion unrolled)
            const Clade &c = clade.at(root);
            if (c.children.empty()) return root;
            unsigned maxCnt = minSub;
            std::vector<int32_t> best;
            for (int32_t ch : c.children) {
                unsigned cnt = clade.at(ch).cladeCount;
                if (cnt > maxCnt) { best.clear(); best.push_back(ch); maxCnt = cnt; }
                else if (cnt == maxCnt) best.push_back(ch);
            }
            if (best.size() == 1) root = best[0]; else return root;
        }
    }
};

}  // namespace

// Classifier::assignTaxonomy (Classifier.cpp:166-208)
void score_reads(const Database &db, const Options &opt, const std::vector<Match> &m, std::vector<QueryInfo> &queries, int threads) {
    struct Block { size_t start, end; uint32_t id; };
    std::vector<Block> blocks;
    size_t i = 0;
    while (i < m.size()) {
        uint32_t id = qi_seq(m[i].qinfo);
        size_t s = i;
        while (i < m.size() && qi_seq(m[i].qinfo) == id) ++i;
        blocks.push_back(Block{s, i - 1, id});
    }
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        Scorer sc(db, opt);
#pragma omp for schedule(dynamic, 64)
        for (long long b = 0; b < (long long)blocks.size(); ++b)
            sc.score_read(m.data(), blocks[(size_t)b].start, blocks[(size_t)b].end, queries[blocks[(size_t)b].id - 1]);
    }
}

// =================================================================================================
// A13  Reporter::writeReadClassification (Reporter.cpp:35-80)
// =================================================================================================
void write_tsv_header(std::string &out, bool lineage) {
    out += "#is_classified\tname\ttaxID\tquery_length\tscore\trank";
    if (lineage) out += "\tlineage";
    out += "\ttaxID:match_count\n";
}

// TaxonomyWrapper::taxLineage2 (TaxonomyWrapper.cpp:431-454) with ExtendedShortRanks (TaxonomyWrapper.h:9-26)
std::string Taxonomy::lineage(int32_t taxId) const {
    static const std::map<std::string, std::string> shortRanks = {
        {"subspecies", "ss"}, {"species", "s"}, {"subgenus", "sg"}, {"genus", "g"}, {"subfamily", "sf"}, {"family", "f"},
        {"suborder", "so"}, {"order", "o"}, {"subclass", "sc"}, {"class", "c"}, {"subphylum", "sp"}, {"phylum", "p"},
        {"subkingdom", "sk"}, {"kingdom", "k"}, {"superkingdom", "d"}, {"domain", "d"}, {"realm", "r"}};
    std::vector<int> chain;
    int node = D[taxId];
    do {
        chain.push_back(node);
        node = D[nodeParent[node]];
    } while (nodeParent[node] != nodeTaxId[node]);
    std::string out;
    for (int i = (int)chain.size() - 1; i >= 0; --i) {
        auto it = shortRanks.find(str(nodeRankIdx[chain[(size_t)i]]));
        out += it == shortRanks.end() ? "-" : it->second;
        out += '_';
        out += str(nodeNameIdx[chain[(size_t)i]]);
        if (i > 0) out += ';';
    }
    return out;
}

static void append_float(std::string &out, float v) {   // ostream << float, default precision 6 (Q12)
    char buf[64];
    snprintf(buf, sizeof buf, "%g", (double)v);
    out += buf;
}

void write_tsv_rows(const Database &db, const std::vector<Read> &m1, const std::vector<QueryInfo> &qs, std::string &out, bool lineage) {
    for (size_t i = 0; i < qs.size(); ++i) {
        const QueryInfo &q = qs[i];
        out += q.isClassified ? "1\t" : "0\t";
        out += m1[i].name; out += '\t';
        out += std::to_string(db.tax.original(q.classification)); out += '\t';
        out += std::to_string(q.queryLength + q.queryLength2); out += '\t';
        append_float(out, q.score); out += '\t';
        if (q.isClassified) {
            out += db.tax.rankOf(q.classification); out += '\t';
            if (lineage) { out += db.tax.lineage(q.classification); out += '\t'; }
            for (auto &t : q.taxCnt) { out += std::to_string(db.tax.original(t.first)); out += ':'; out += std::to_string(t.second); out += ' '; }
            out += '\n';
        } else {
            out += lineage ? "-\t-\t-\t\n" : "-\t-\t\n";
        }
    }
}

// =================================================================================================
// Reporter::writeReportFile (Reporter.cpp:117-137), writeReport (:166-193), NcbiTaxonomy::getParentToChildren / getCladeCounts
// (NcbiTaxonomy.cpp:504-545)
// =================================================================================================
namespace {
struct CladeCnt { unsigned taxCount = 0, cladeCount = 0; std::vector<int32_t> children; };
void report_node(const Database &db, const std::unordered_map<int32_t, CladeCnt> &cc, unsigned long total, int32_t taxId, int depth,
                 std::string &out) {
    auto it = cc.find(taxId);
    const unsigned cladeCount = it == cc.end() ? 0 : it->second.cladeCount, taxCount = it == cc.end() ? 0 : it->second.taxCount;
    char line[512];
    if (taxId == 0) {
        if (cladeCount > 0) {
            snprintf(line, sizeof line, "%.4f\t%i\t%i\tno rank\t0\tunclassified\n", 100 * cladeCount / double(total), cladeCount, taxCount);
            out += line;
        }
        report_node(db, cc, total, 1, 0, out);
        return;
    }
    if (cladeCount == 0) return;
    const int node = db.tax.D[taxId];
    snprintf(line, sizeof line, "%.4f\t%i\t%i\t%s\t%i\t%s%s\n", 100 * cladeCount / double(total), cladeCount, taxCount,
             db.tax.str(db.tax.nodeRankIdx[(size_t)node]), db.tax.original(taxId), std::string((size_t)(2 * depth), ' ').c_str(),
             db.tax.str(db.tax.nodeNameIdx[(size_t)node]));
    out += line;
    std::vector<int32_t> children = it->second.children;
    auto val = [&](int32_t k) { auto f = cc.find(k); return f == cc.end() ? 0u : f->second.cladeCount; };
    std::sort(children.begin(), children.end(), [&](int a, int b) { return val(a) > val(b); });     // SORT_SERIAL = std::sort
    for (int32_t c : children) {
        if (cc.count(c)) report_node(db, cc, total, c, depth + 1, out); else break;
    }
}
}  // namespace

void write_report(const Database &db, const std::vector<QueryInfo> &qs, std::string &out) {
    std::unordered_map<int32_t, unsigned> taxCnt;                       // Classifier.cpp:196-203
    for (const QueryInfo &q : qs) ++taxCnt[q.classification];
    std::unordered_map<int32_t, std::vector<int32_t>> p2c;
    for (size_t i = 0; i < db.tax.maxNodes; ++i)
        if (db.tax.nodeParent[i] != db.tax.nodeTaxId[i]) p2c[db.tax.nodeParent[i]].push_back(db.tax.nodeTaxId[i]);
    std::unordered_map<int32_t, CladeCnt> cc;
    for (auto &kv : taxCnt) {
        cc[kv.first].taxCount = kv.second;
        cc[kv.first].cladeCount += kv.second;
        if (db.tax.nodeExists(kv.first)) {
            int node = db.tax.D[kv.first];
            while (db.tax.nodeParent[(size_t)node] != db.tax.nodeTaxId[(size_t)node] && db.tax.nodeExists(db.tax.nodeParent[(size_t)node])) {
                node = db.tax.D[db.tax.nodeParent[(size_t)node]];
                cc[db.tax.nodeTaxId[(size_t)node]].cladeCount += kv.second;
            }
        }
    }
    for (auto &kv : cc) { auto f = p2c.find(kv.first); if (f != p2c.end()) kv.second.children = f->second; }
    out += "#clade_proportion\tclade_count\ttaxon_count\trank\ttaxID\tname\n";
    report_node(db, cc, (unsigned long)qs.size(), 0, 0, out);
}

// =================================================================================================
// whole path
// =================================================================================================
bool classify_files(const std::string &q1, const std::string &q2, const std::string &dbDir, const Options &opt,
                    std::string &tsv, std::string *err, size_t *nKmers, size_t *nMatches, std::string *report) {
    Database db;
    if (!db.load(dbDir, err)) return false;
    std::vector<Read> m1, m2;
    if (!read_fastx(q1, m1, err)) return false;
    bool paired = opt.seqMode == 2;
    if (paired) {
        if (!read_fastx(q2, m2, err)) return false;
        if (m1.size() != m2.size()) { if (err) *err = "The number of reads in the two files are not equal."; return false; }
    }
    // loadDbParameters (common.cpp:101-108): the database's Accession_level adjusts --accession-level
    Options o = opt;
    if (db.params.accessionLevelDb == 0 && o.accessionLevel == 1) o.accessionLevel = 0;
    if (db.params.accessionLevelDb == 1 && o.accessionLevel == 0) o.accessionLevel = 2;
    std::vector<QueryInfo> queries;
    std::vector<Kmer> kmers;
    extract_kmers(m1, paired ? &m2 : nullptr, db.params.kmerFormat, queries, kmers, db.params.syncmer, db.params.smerLen);
    sort_kmers(kmers, opt.threads);
    if (nKmers) { size_t b = 0; while (b < kmers.size() && qi_seq(kmers[b].qinfo) == 0) ++b; *nKmers = kmers.size() - b; }
    std::vector<Match> matches;
    if (!match_kmers(db, kmers, matches, opt.threads, err)) return false;
    if (nMatches) *nMatches = matches.size();
    sort_matches(matches, opt.threads);
    score_reads(db, o, matches, queries, opt.threads);
    tsv.clear();
    write_tsv_header(tsv, opt.printLineage != 0);
    write_tsv_rows(db, m1, queries, tsv, opt.printLineage != 0);
    if (report) { report->clear(); write_report(db, queries, *report); }
    return true;
}

}  // namespace orc
