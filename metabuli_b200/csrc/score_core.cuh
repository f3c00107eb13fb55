// Per-read taxon scoring (reference rows A10-A12: Taxonomer.cpp:130-202 chooseBestTaxon, :205-241
// filterRedundantMatches, :252-314 lowerRankClassification/BFS, :316-408 getBestSpeciesMatches,
// :410-485 combineMatchPaths/trimMatchPath, :487-648 getMatchPaths, Match.h:32-88, NcbiTaxonomy.cpp:250-330).
//
// Written as __host__ __device__ code over flat arrays: the device kernel (k5_score.cu) runs one read per
// thread, and the CPU unit tests compile the very same functions to check the host logic without a GPU.
// All per-read state lives in caller-provided scratch indexed by match index (no allocation, no hash
// maps): the reference's unordered_map / vector state is replaced by small in-place lists whose results
// are provably order-independent — except the path sort, which replays libstdc++'s
// std::sort decision sequence exactly (quirk Q4).
#pragma once
#include "kernels.cuh"

namespace mbl {

// ---- libstdc++-compatible introsort on a permutation (std::sort, bits/stl_algo.h) --------------------
// Replays the comparison/move sequence of GCC's std::sort so that elements the comparator cannot
// separate end up in the same order as in the reference binary.
template <class Less>
MBL_HD void stl_unguarded_linear_insert(int32_t* a, int last, Less less) {
    int32_t val = a[last];
    int next = last - 1;
    while (less(val, a[next])) { a[last] = a[next]; last = next; --next; }
    a[last] = val;
}
template <class Less>
MBL_HD void stl_insertion_sort(int32_t* a, int first, int last, Less less) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(a[i], a[first])) {
            int32_t val = a[i];
            for (int j = i; j > first; --j) a[j] = a[j - 1];
            a[first] = val;
        } else {
            stl_unguarded_linear_insert(a, i, less);
        }
    }
}
template <class Less>
MBL_HD void stl_adjust_heap(int32_t* a, int first, int hole, int len, int32_t value, Less less) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(a[first + child], a[first + child - 1])) --child;
        a[first + hole] = a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[first + hole] = a[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && less(a[first + parent], value)) {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}
template <class Less>
MBL_HD void stl_heap_sort(int32_t* a, int first, int last, Less less) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            stl_adjust_heap(a, first, parent, len, a[first + parent], less);
            if (parent == 0) break;
            --parent;
        }
    }
    while (last - first > 1) {
        --last;
        int32_t value = a[last];
        a[last] = a[first];
        stl_adjust_heap(a, first, 0, last - first, value, less);
    }
}
template <class Less>
MBL_HD void stl_sort(int32_t* a, int n, Less less) {
    if (n <= 0) return;
    if (n > 16) {
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) ++lg;
        int st_first[64], st_last[64], st_depth[64];
        int sp = 0;
        st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; sp = 1;
        while (sp > 0) {
            --sp;
            int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
            while (last - first > 16) {
                if (depth == 0) { stl_heap_sort(a, first, last, less); break; }
                --depth;
                // median of (first+1, mid, last-1) to first
                const int mid = first + (last - first) / 2;
                const int x = first + 1, y = mid, z = last - 1;
                int m;
                if (less(a[x], a[y])) { m = less(a[y], a[z]) ? y : (less(a[x], a[z]) ? z : x); }
                else { m = less(a[x], a[z]) ? x : (less(a[y], a[z]) ? z : y); }
                { int32_t t = a[first]; a[first] = a[m]; a[m] = t; }
                // unguarded partition around a[first]
                int lo = first + 1, hi = last;
                while (true) {
                    while (less(a[lo], a[first])) ++lo;
                    --hi;
                    while (less(a[first], a[hi])) --hi;
                    if (!(lo < hi)) break;
                    int32_t t = a[lo]; a[lo] = a[hi]; a[hi] = t;
                    ++lo;
                }
                if (sp < 64) { st_first[sp] = lo; st_last[sp] = last; st_depth[sp] = depth; ++sp; }
                last = lo;
            }
        }
        stl_insertion_sort(a, 0, 16, less);
        for (int i = 16; i < n; ++i) stl_unguarded_linear_insert(a, i, less);
    } else {
        stl_insertion_sort(a, 0, n, less);
    }
}

// ---- taxonomy primitives -----------------------------------------------------------------------------
MBL_HD bool tax_exists(const DeviceTaxonomy& t, int32_t id) { return id >= 0 && id <= t.max_taxid && t.D[id] != -1; }
MBL_HD int tax_lca_nodes(const DeviceTaxonomy& t, int i, int j) {          // NcbiTaxonomy::lcaHelper (Q5)
    if (i == 0 || j == 0) return 0;
    if (i == j) return i;
    int v1 = t.H[i], v2 = t.H[j];
    if (v1 > v2) { int x = v1; v1 = v2; v2 = x; }
    int k = 0;
    for (unsigned span = (unsigned)(v2 - v1 + 1); span > 1; span >>= 1) ++k;
    const int a = t.M[(size_t)v1 * t.M_k + k];
    const int b = t.M[(size_t)(v2 - (1 << k) + 1) * t.M_k + k];
    return t.E[(t.L[a] <= t.L[b]) ? a : b];
}
MBL_HD int32_t tax_lca(const DeviceTaxonomy& t, int32_t a, int32_t b) {      // NcbiTaxonomy::LCA(TaxID,TaxID)
    if (!tax_exists(t, a)) return b;
    if (!tax_exists(t, b)) return a;
    return t.node_taxid[tax_lca_nodes(t, t.D[a], t.D[b])];
}
MBL_HD bool tax_is_ancestor(const DeviceTaxonomy& t, int32_t anc, int32_t child) {
    if (anc == child) return true;
    if (anc == 0 || child == 0) return false;
    if (!tax_exists(t, child) || !tax_exists(t, anc)) return false;
    return tax_lca_nodes(t, t.D[child], t.D[anc]) == t.D[anc];
}
MBL_HD int32_t tax_parent(const DeviceTaxonomy& t, int32_t id) { return tax_exists(t, id) ? t.node_parent[t.D[id]] : 0; }

// ---- Match score helpers (Match.h:32-88) ---------------------------------------------------------------
MBL_HD float codon_score(int d) { return d == 0 ? 3.0f : 2.0f - 0.5f * (float)d; }
MBL_HD int popc16(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// Match::getScore (Match.h:32-44): sum over the 8 codons of {0:3.0, 1:1.5, 2:1.0, 3:0.5}.  Every term is a multiple of
// 0.5, so the float sum is exact in any order; count the codons per distance instead of looping.
MBL_HD float match_score(uint32_t reh) {
    const uint32_t lo = reh & 0x5555u, hi = (reh >> 1) & 0x5555u;
    const int n3 = popc16(lo & hi), n2 = popc16(hi & ~lo), n1 = popc16(lo & ~hi);
    const int n0 = 8 - n1 - n2 - n3;
    return 3.0f * (float)n0 + 1.5f * (float)n1 + 1.0f * (float)n2 + 0.5f * (float)n3;
}
MBL_HD float right_part_score(uint32_t reh, int range) { float s = 0.f; for (int i = 0; i < range; ++i) s += codon_score((reh >> (2 * i)) & 3); return s; }
MBL_HD float left_part_score(uint32_t reh, int range) { float s = 0.f; for (int i = 0; i < range; ++i) s += codon_score((reh >> (14 - 2 * i)) & 3); return s; }
MBL_HD int right_part_ham(uint32_t reh, int range) { int s = 0; for (int i = 0; i < range; ++i) s += (reh >> (2 * i)) & 3; return s; }
MBL_HD int left_part_ham(uint32_t reh, int range) { int s = 0; for (int i = 0; i < range; ++i) s += (reh >> (14 - 2 * i)) & 3; return s; }

// ---- getMatchPaths for one (species, frame) group [gs, ge) (Taxonomer.cpp:487-648) ----------------------
// Emitted paths are appended at a.p_*[pbase + np].
//
// The DP only ever looks one position back, so when no position of the group holds more than kDpWidth
// matches (the common case: one or two target k-mers per query position and species) the state of the
// "current" and "next" position lives in registers and nothing but the emitted paths touches HBM.
// Groups with a wider position fall back to the scratch-array version below (same arithmetic, same order).
constexpr int kDpWidth = 2;

// Every array below is indexed by unrolled compile-time constants only, so the DP state really stays in registers (a version that
// appended with buf[n++] and swapped two buffers through pointers compiled to a 256-byte stack frame: every cell access was a
// local-memory load on the thread's dependent chain), and a row is fetched as three 8-byte words two rows ahead of its use.
struct DpCell { float score; int32_t start, ham, depth; uint32_t smatch, idx, dna, reh; bool conn; };   // (all cells of a buffer share the position)
struct Row3 { uint64_t w0, w1, w2; };                // qinfo | target, species | dna, right_end_hamming, hamming
MBL_HD Row3 load_row(const mbl_match_rec* ml, uint64_t i) {
    const uint64_t* s = reinterpret_cast<const uint64_t*>(ml + i);
    return Row3{s[0], s[1], s[2]};
}
MBL_HD void dp_put(DpCell (&buf)[kDpWidth], int at, const DpCell& x) {
#pragma unroll
    for (int k = 0; k < kDpWidth; ++k) if (k == at) buf[k] = x;
}

MBL_HD bool score_frame_group_fast(const ScoreArgs& a, uint64_t gs, uint64_t ge, int min_depth, uint64_t pbase, uint32_t& np) {
    const mbl_match_rec* ml = a.matches;
    const uint32_t np_in = np;                       // a position wider than kDpWidth: undo and let the caller fall back
    const bool fmt2 = a.par.kmer_format == 2;
    DpCell cur[kDpWidth], nxt[kDpWidth];
    int ncur = 0, nnxt = 0;
    uint64_t i = gs;
    Row3 r0 = load_row(ml, gs);
    Row3 r1 = gs + 1 < ge ? load_row(ml, gs + 1) : r0;
    const bool forward = (uint32_t)(r0.w0 >> 61) < 3;
    auto advance = [&]() {                           // row i consumed: r0 becomes row i + 1, row i + 2 is requested
        ++i;
        r0 = r1;
        if (i + 1 < ge) r1 = load_row(ml, i + 1);
    };
    auto cell = [&]() {                              // a fresh path that starts (and ends) at row i
        DpCell c;
        c.reh = (uint32_t)(r0.w2 >> 32) & 0xFFFFu;
        c.score = match_score(c.reh);
        c.start = (int32_t)(uint32_t)r0.w0;
        c.ham = (int32_t)((r0.w2 >> 48) & 0xFFu); c.depth = 1; c.smatch = (uint32_t)(i - gs); c.idx = (uint32_t)(i - gs);
        c.dna = (uint32_t)r0.w2; c.conn = false;
        return c;
    };
    auto push = [&](const DpCell& c, uint64_t pos) {
        const uint64_t o = pbase + np++;
        a.p_start[o] = c.start;
        a.p_end[o] = (int32_t)pos + 23;
        a.p_score[o] = c.score; a.p_ham[o] = c.ham; a.p_depth[o] = c.depth;
        a.p_smatch[o] = (uint32_t)(gs + c.smatch);                // absolute match indices
        a.p_ematch[o] = (uint32_t)(gs + c.idx);
    };
    uint64_t curPos = (uint32_t)r0.w0;
    bool wide = false;                               // keep going on overflow (results are discarded), no extra control flow
    while (i < ge && (uint32_t)r0.w0 == curPos) {
        if (ncur < kDpWidth) { dp_put(cur, ncur, cell()); ++ncur; } else wide = true;
        advance();
    }
    while (i < ge) {
        // one DP step: `cur` holds the paths ending at curPos, `nxt` receives those ending at the next position
        const uint32_t nextPos = (uint32_t)r0.w0;
        nnxt = 0;
        while (i < ge && (uint32_t)r0.w0 == nextPos) {
            if (nnxt < kDpWidth) { dp_put(nxt, nnxt, cell()); ++nnxt; } else wide = true;
            advance();
        }
        const int shift = (int)(((uint64_t)nextPos - curPos) / 3);
        if (shift > 0 && shift <= a.par.max_codon_shift) {              // maxCodonShift: 1, or 8 - s with syncmers (Taxonomer.cpp:34-42)
            const uint32_t lowMask = (1u << (24 - 3 * shift)) - 1;
#pragma unroll
            for (int nx = 0; nx < kDpWidth; ++nx) {
                if (nx < nnxt) {
                    const uint32_t reh = nxt[nx].reh;
                    int h = 0;                                           // calHammingDistIncrement / calScoreIncrement (:650-669)
                    float inc = 0.f;
                    for (int sft = 0; sft < shift; ++sft) { const int d = (reh >> (2 * sft)) & 3; h += d; inc += codon_score(d); }
                    int best = -1;
                    float bestScore = 0.f;
#pragma unroll
                    for (int cu = 0; cu < kDpWidth; ++cu) {
                        if (cu < ncur) {
                            const uint32_t m1 = forward ? cur[cu].dna : nxt[nx].dna, m2 = forward ? nxt[nx].dna : cur[cu].dna;
                            const bool cons = fmt2 ? ((m1 & lowMask) == (m2 >> (3 * shift))) : ((m1 >> (3 * shift)) == (m2 & lowMask));
                            if (cons) {
                                cur[cu].conn = true;
                                if (cur[cu].score > bestScore) { best = cu; bestScore = cur[cu].score; }
                            }
                        }
                    }
#pragma unroll
                    for (int cu = 0; cu < kDpWidth; ++cu)
                        if (cu == best) {
                            nxt[nx].start = cur[cu].start; nxt[nx].score = cur[cu].score + inc; nxt[nx].ham = cur[cu].ham + h;
                            nxt[nx].depth = cur[cu].depth + shift; nxt[nx].smatch = cur[cu].smatch;
                        }
                }
            }
        }
#pragma unroll
        for (int cu = 0; cu < kDpWidth; ++cu)
            if (cu < ncur && !cur[cu].conn && cur[cu].depth >= min_depth) push(cur[cu], curPos);
        if (i == ge) {
#pragma unroll
            for (int nx = 0; nx < kDpWidth; ++nx)
                if (nx < nnxt && nxt[nx].depth >= min_depth) push(nxt[nx], nextPos);
        }
#pragma unroll
        for (int k = 0; k < kDpWidth; ++k) cur[k] = nxt[k];
        ncur = nnxt;
        curPos = nextPos;
    }
    if (wide) { np = np_in; return false; }
    return true;
}

MBL_HD void score_frame_group(const ScoreArgs& a, uint64_t gs, uint64_t ge, int min_depth, uint64_t pbase, uint32_t& np) {
    if (!a.par.force_scratch_dp && score_frame_group_fast(a, gs, ge, min_depth, pbase, np)) return;
    const mbl_match_rec* ml = a.matches;
    const bool forward = qi_frame(ml[gs].qinfo) < 3;
    const bool fmt2 = a.par.kmer_format == 2;
    auto init = [&](uint64_t i) {
        a.l_start[i] = (int32_t)qi_pos(ml[i].qinfo);
        a.l_score[i] = match_score(ml[i].right_end_hamming);
        a.l_ham[i] = ml[i].hamming;
        a.l_depth[i] = 1;
        a.l_smatch[i] = (uint32_t)(i - gs);
        a.l_conn[i] = 0;
    };
    auto push = [&](uint64_t i) {
        const uint64_t o = pbase + np++;
        a.p_start[o] = a.l_start[i];
        a.p_end[o] = (int32_t)qi_pos(ml[i].qinfo) + 23;
        a.p_score[o] = a.l_score[i];
        a.p_ham[o] = a.l_ham[i];
        a.p_depth[o] = a.l_depth[i];
        a.p_smatch[o] = (uint32_t)(gs + a.l_smatch[i]);           // absolute match indices
        a.p_ematch[o] = (uint32_t)i;
    };
    uint64_t i = gs;
    uint64_t curPos = qi_pos(ml[gs].qinfo);
    uint64_t curS = i;
    while (i < ge && qi_pos(ml[i].qinfo) == curPos) { init(i); ++i; }
    uint64_t curE = i;
    while (i < ge) {
        const uint32_t nextPos = qi_pos(ml[i].qinfo);
        const uint64_t nxtS = i;
        while (i < ge && qi_pos(ml[i].qinfo) == nextPos) { init(i); ++i; }
        const uint64_t nxtE = i;
        const int shift = (int)(((uint64_t)nextPos - curPos) / 3);
        if (shift > 0 && shift <= a.par.max_codon_shift) {              // maxCodonShift = 1 without syncmers, 8 - s with
            const uint32_t lowMask = (1u << (24 - 3 * shift)) - 1;
            for (uint64_t nx = nxtS; nx < nxtE; ++nx) {
                const uint32_t reh = ml[nx].right_end_hamming;
                int h = 0;                                               // calScoreIncrement / calHammingDistIncrement
                float inc = 0.f;
                for (int sft = 0; sft < shift; ++sft) { const int d = (reh >> (2 * sft)) & 3; h += d; inc += codon_score(d); }
                const uint32_t ndna = ml[nx].dna_encoding;
                int64_t best = -1;
                float bestScore = 0.f;
                for (uint64_t cu = curS; cu < curE; ++cu) {
                    const uint32_t cdna = ml[cu].dna_encoding;
                    // isConsecutive2 / isConsecutive (Taxonomer.cpp:677-699); reverse frames swap the operands
                    const uint32_t m1 = forward ? cdna : ndna, m2 = forward ? ndna : cdna;
                    const bool cons = fmt2 ? ((m1 & lowMask) == (m2 >> (3 * shift))) : ((m1 >> (3 * shift)) == (m2 & lowMask));
                    if (cons) {
                        a.l_conn[cu] = 1;
                        if (a.l_score[cu] > bestScore) { best = (int64_t)cu; bestScore = a.l_score[cu]; }
                    }
                }
                if (best >= 0) {
                    a.l_start[nx] = a.l_start[best];
                    a.l_score[nx] = a.l_score[best] + inc;
                    a.l_ham[nx] = a.l_ham[best] + h;
                    a.l_depth[nx] = a.l_depth[best] + shift;
                    a.l_smatch[nx] = a.l_smatch[best];
                }
            }
        }
        for (uint64_t cu = curS; cu < curE; ++cu)
            if (!a.l_conn[cu] && a.l_depth[cu] >= min_depth) push(cu);
        if (i == ge)
            for (uint64_t nx = nxtS; nx < nxtE; ++nx)
                if (a.l_depth[nx] >= min_depth) push(nx);
        curS = nxtS; curE = nxtE; curPos = nextPos;
    }
}

// ---- combineMatchPaths for the np paths of one species at pbase (Taxonomer.cpp:410-468) ------------------
// `perm` (at a.l_start + pbase: the DP scratch is free by now, np <= matches of the species) lists the paths of the
// species as offsets from pbase in the order the reference generated them; perm_ready says it is already filled.
MBL_HD float score_combine(const ScoreArgs& a, uint64_t pbase, uint32_t np, int read_length, bool perm_ready = false) {
    int32_t* perm = a.l_start + pbase;
    if (!perm_ready) for (uint32_t i = 0; i < np; ++i) perm[i] = (int32_t)i;
    const float* ps = a.p_score + pbase;
    const int32_t* ph = a.p_ham + pbase;
    const int32_t* pst = a.p_start + pbase;
    stl_sort(perm, (int)np, [&](int32_t x, int32_t y) {
        if (ps[x] != ps[y]) return ps[x] > ps[y];
        if (ph[x] != ph[y]) return ph[x] < ph[y];
        return pst[x] > pst[y];
    });
    const mbl_match_rec* ml = a.matches;
    float score = 0.f;
    uint32_t nc = 0;
    for (uint32_t ii = 0; ii < np; ++ii) {
        const uint64_t p = pbase + (uint32_t)perm[ii];
        if (nc == 0) {
            a.c_start[pbase + nc] = a.p_start[p]; a.c_end[pbase + nc] = a.p_end[p]; ++nc;
            score += a.p_score[p];
            continue;
        }
        bool overlapped = false;
        for (uint32_t j = 0; j < nc; ++j) {
            const int cs = a.c_start[pbase + j], ce = a.c_end[pbase + j];
            if (!((a.p_end[p] < cs) || (ce < a.p_start[p]))) {
                const int ov = min(a.p_end[p], ce) - max(a.p_start[p], cs) + 1;
                if (ov == a.p_end[p] - a.p_start[p] + 1) { overlapped = true; break; }
                if (ov < 24) {                                           // trimMatchPath (Taxonomer.cpp:475-485)
                    if (a.p_start[p] < cs) {
                        const uint32_t reh = ml[a.p_ematch[p]].right_end_hamming;
                        a.p_end[p] = cs - 1;
                        a.p_ham[p] = max(0, a.p_ham[p] - right_part_ham(reh, ov / 3));
                        a.p_score[p] = a.p_score[p] - right_part_score(reh, ov / 3) - (float)(ov % 3);
                    } else {
                        const uint32_t reh = ml[a.p_smatch[p]].right_end_hamming;
                        a.p_start[p] = ce + 1;
                        a.p_ham[p] = max(0, a.p_ham[p] - left_part_ham(reh, ov / 3));
                        a.p_score[p] = a.p_score[p] - left_part_score(reh, ov / 3) - (float)(ov % 3);
                    }
                    continue;
                }
                overlapped = true;
                break;
            }
        }
        if (!overlapped) {
            a.c_start[pbase + nc] = a.p_start[p]; a.c_end[pbase + nc] = a.p_end[p]; ++nc;
            score += a.p_score[p];
        }
    }
    return score / (float)read_length;
}

// =====================================================================================================================
// Flat formulation used by the CUDA pipeline (k5_score.cu): the nested loops of chooseBestTaxon are cut into three
// passes over flat task lists so that the threads of a warp do the same kind of work —
//   A. one (species, frame) group per thread : getMatchPaths            -> paths at p_*[group start + k], g_np[group start]
//   B. one (read, species) group per thread   : combineMatchPaths        -> s_score[species group start]
//   C. one read per thread                    : tie set / LCA / votes / clade descent -> mbl_read_result
// fg_list / sp_list hold the first match index of every frame group / species group, ascending.
// =====================================================================================================================
MBL_HD uint32_t list_lower_bound(const uint32_t* list, uint32_t n, uint64_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((uint64_t)list[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// rows a (species, frame) group needs before getMatchPaths can emit anything: 2 (Q9); without syncmers depth == matches of the path
MBL_HD int min_group_rows(const ScoreParams& p) {
    if (p.max_codon_shift > 1) return 2;
    const int m = p.min_cons_cnt < p.min_cons_cnt_euk ? p.min_cons_cnt : p.min_cons_cnt_euk;
    return m > 2 ? m : 2;
}

MBL_HD void score_task_frame_group(const ScoreArgs& a, uint32_t t) {
    const uint32_t g = a.fg_order ? a.fg_order[t] : t;             // tasks ordered by group length (k5_score.cu)
    const uint64_t gs = a.fg_list[g];
    const uint64_t ge = g + 1 < a.n_fg ? a.fg_list[g + 1] : a.match_end;
    // Q9 (a frame group with a single match contributes nothing, Taxonomer.cpp:342) and its generalisation: without syncmers a
    // path's depth is its number of matches, so a group with fewer rows than the smaller minimum depth cannot emit a path — most
    // groups (chance hits of one to three rows) end here without touching their rows; score_task_species skips them the same way
    if (ge - gs < (uint64_t)min_group_rows(a.par)) return;
    uint32_t np = 0;
    const int32_t species = a.matches[gs].species_id;
    int min_depth = a.par.min_cons_cnt;
    if (tax_is_ancestor(a.tax, a.tax.eukaryota, species)) min_depth = a.par.min_cons_cnt_euk;
    score_frame_group(a, gs, ge, min_depth, gs, np);
    a.g_np[gs] = np;
}

MBL_HD void score_task_species(const ScoreArgs& a, uint32_t sidx) {
    const uint64_t spS = a.sp_list[sidx];
    const uint64_t spE = sidx + 1 < a.n_sp ? a.sp_list[sidx + 1] : a.match_end;
    int32_t* perm = a.l_start + spS;
    uint32_t np = 0;
    const uint64_t min_rows = (uint64_t)min_group_rows(a.par);
    float& result = a.sp_score ? a.sp_score[sidx] : a.s_score[spS];
    if (spE - spS < min_rows) { result = -3.0e38f; return; }                      // not even one frame group long enough for a path
    for (uint32_t g = a.sp_fg ? a.sp_fg[sidx] : list_lower_bound(a.fg_list, a.n_fg, spS); g < a.n_fg && a.fg_list[g] < spE; ++g) {
        const uint64_t gs = a.fg_list[g];
        const uint64_t ge = g + 1 < a.n_fg ? a.fg_list[g + 1] : a.match_end;
        if (ge - gs < min_rows) continue;                                         // no paths, g_np not written (score_task_frame_group)
        const uint32_t cnt = a.g_np[gs];
        for (uint32_t k = 0; k < cnt; ++k) perm[np++] = (int32_t)(gs - spS + k);
    }
    float out = -3.0e38f;                                                         // "no entry in sp2score"
    if (np > 0) {
        const uint32_t r = qi_seq(a.matches[spS].qinfo) - 1;
        float score = score_combine(a, spS, np, a.cov1[r] + a.cov2[r], true);
        score = fminf(score, 1.0f);
        if (!(score < a.par.min_score)) out = score;
    }
    result = out;
}

// ---- chooseBestTaxon for read r -----------------------------------------------------------------------
MBL_HD void score_read(const ScoreArgs& a, uint32_t r) {
    mbl_read_result res;
    res.classification = 0; res.score = 0.f; res.hamming = 0;
    res.query_length = a.cov1[r] + a.cov2[r];
    res.taxcnt_begin = a.quot_off[r]; res.taxcnt_len = 0; res.is_classified = 0;
    res.pad[0] = res.pad[1] = res.pad[2] = 0;
    const uint64_t ms = a.seg_begin[r], me = a.seg_end[r];
    if (me <= ms) { a.results[r] = res; return; }
    const mbl_match_rec* ml = a.matches;
    const DeviceTaxonomy& tx = a.tax;
    const int queryLength = res.query_length;

    // --- getBestSpeciesMatches, pass 1: per-species scores (kept at the species' first match index)
    float bestSpScore = 0.f;
    uint32_t meaningful = 0;
    uint64_t bestS = 0, bestE = 0;
    uint64_t i = ms;
    if (a.sp_list) {
        // flat pipeline: the species scores are already in s_score (tasks A and B)
        for (uint32_t sidx = a.read_sp ? a.read_sp[r] : list_lower_bound(a.sp_list, a.n_sp, ms); sidx < a.n_sp && a.sp_list[sidx] < me; ++sidx) {
            const uint64_t spS = a.sp_list[sidx];
            const uint64_t spE = sidx + 1 < a.n_sp && a.sp_list[sidx + 1] < me ? a.sp_list[sidx + 1] : me;
            const float score = a.sp_score ? a.sp_score[sidx] : a.s_score[spS];
            if (score > -1.0e38f) {
                if (score > 0.f) ++meaningful;
                if (score > bestSpScore) { bestSpScore = score; bestS = spS; bestE = spE; }
            }
        }
        i = me;
    }
    while (i < me) {
        const int32_t species = ml[i].species_id;
        const uint64_t spS = i;
        uint32_t np = 0;
        int min_depth = a.par.min_cons_cnt;
        if (tax_is_ancestor(tx, tx.eukaryota, species)) min_depth = a.par.min_cons_cnt_euk;
        while (i < me && ml[i].species_id == species) {
            const uint32_t frame = qi_frame(ml[i].qinfo);
            const uint64_t fs = i;
            while (i < me && ml[i].species_id == species && qi_frame(ml[i].qinfo) == frame) ++i;
            if (i - fs > 1) score_frame_group(a, fs, i, min_depth, spS, np);     // Q9
        }
        a.s_score[spS] = -3.0e38f;                                              // "no entry in sp2score"
        if (np > 0) {
            float score = score_combine(a, spS, np, queryLength);
            score = fminf(score, 1.0f);
            if (score < a.par.min_score) continue;
            a.s_score[spS] = score;
            if (score > 0.f) ++meaningful;
            if (score > bestSpScore) { bestSpScore = score; bestS = spS; bestE = i; }
        }
    }
    // --- pass 2: tie set, LCA (Taxonomer.cpp:372-407)
    float finalScore = 0.f;
    int32_t taxId = 0;
    uint32_t nMax = 0;
    if (meaningful != 0) {
        const float thr = bestSpScore * a.par.tie_ratio;
        int red = 0;
        bool haveRed = false;
        i = ms;
        uint32_t sidx = a.sp_list ? (a.read_sp ? a.read_sp[r] : list_lower_bound(a.sp_list, a.n_sp, ms)) : 0;
        while (i < me) {
            const int32_t species = ml[i].species_id;
            const uint64_t spS = i;
            const uint32_t this_sidx = sidx;
            if (a.sp_list) { ++sidx; i = sidx < a.n_sp && a.sp_list[sidx] < me ? a.sp_list[sidx] : me; }
            else while (i < me && ml[i].species_id == species) ++i;
            const float sc = (a.sp_list && a.sp_score) ? a.sp_score[this_sidx] : a.s_score[spS];
            if (sc > -1.0e38f && sc >= thr) {
                if (nMax == 0) taxId = species;
                ++nMax;
                finalScore += sc;
                if (tax_exists(tx, species)) {                                   // NcbiTaxonomy::LCA(vector) fold
                    red = haveRed ? tax_lca_nodes(tx, red, tx.D[species]) : tx.D[species];
                    haveRed = true;
                }
            }
        }
        if (nMax > 1) {
            taxId = haveRed ? tx.node_taxid[red] : 0;
            finalScore /= (float)nMax;
        }
    }
    // --- chooseBestTaxon
    if (finalScore == 0.f || finalScore < a.par.min_score) {
        res.score = finalScore;
        a.results[r] = res;
        return;
    }
    res.is_classified = 1;
    res.score = finalScore;
    if (nMax > 1) { res.classification = taxId; a.results[r] = res; return; }

    // --- filterRedundantMatches (Taxonomer.cpp:205-241): one vote per pos/3 quotient of the best species
    const uint32_t q0 = a.quot_off[r], nq = a.quot_off[r + 1] - q0;
    for (uint32_t k = 0; k < nq; ++k) a.q_has[q0 + k] = 0;
    for (uint64_t k = bestS; k < bestE; ++k) {
        const uint32_t quo = qi_pos(ml[k].qinfo) / (uint32_t)a.par.dna_shift;      // Taxonomer.cpp:217 (dnaShift: 3, or 3 (8 - s) with syncmers)
        if (quo >= nq) continue;
        const uint8_t h = ml[k].hamming;
        if (!a.q_has[q0 + quo] || h < a.q_ham[q0 + quo]) { a.q_has[q0 + quo] = 1; a.q_tax[q0 + quo] = ml[k].target_id; a.q_ham[q0 + quo] = h; }
        else if (h == a.q_ham[q0 + quo]) a.q_tax[q0 + quo] = tax_lca(tx, a.q_tax[q0 + quo], ml[k].target_id);
    }
    // distinct taxids with counts, ascending taxid (std::map order of Query::taxCnt)
    int32_t* pairs = a.taxcnt_pairs + 2ull * q0;
    uint32_t nt = 0;
    for (uint32_t k = 0; k < nq; ++k) {
        if (!a.q_has[q0 + k]) continue;
        const int32_t t = a.q_tax[q0 + k];
        uint32_t p = 0;
        while (p < nt && pairs[2 * p] < t) ++p;
        if (p < nt && pairs[2 * p] == t) { ++pairs[2 * p + 1]; continue; }
        for (uint32_t m = nt; m > p; --m) { pairs[2 * m] = pairs[2 * m - 2]; pairs[2 * m + 1] = pairs[2 * m - 1]; }
        pairs[2 * p] = t; pairs[2 * p + 1] = 1; ++nt;
    }
    res.taxcnt_len = nt;

    if (finalScore < a.par.min_sp_score) {                                      // Taxonomer.cpp:172-180
        int32_t sp = taxId;                                                     // getTaxIdAtRank(species, "species")
        if (taxId != 0 && taxId != 1 && tax_exists(tx, taxId)) {
            int node = tx.D[taxId], cnt = 0;
            while (cnt < 30 && tx.node_rank[node] < 4) { node = tx.D[tx.node_parent[node]]; ++cnt; }
            sp = cnt == 30 ? taxId : tx.node_taxid[node];
        } else sp = 0;
        res.classification = tax_parent(tx, sp);
        a.results[r] = res;
        return;
    }

    // --- lowerRankClassification + BFS (Taxonomer.cpp:252-314) without hash maps
    const unsigned minSub = (unsigned)((queryLength - 1) / a.par.denominator);
    int32_t* child = a.q_tax + q0;               // the quotient table is free now; nt <= nq
    int32_t root = taxId;
    for (int guard = 0; guard < 256; ++guard) {
        bool any = false;
        for (uint32_t k = 0; k < nt; ++k) {      // the child of `root` on the way to each counted taxon
            int32_t cur = pairs[2 * k], c = 0;
            for (int d = 0; d < 512 && cur != root && cur != 0; ++d) {
                const int32_t p = tax_parent(tx, cur);
                if (p == root) { c = cur; break; }
                if (p == cur) break;
                cur = p;
            }
            if (c != 0 && a.par.accession_level == 2 && tx.node_prune[tx.D[c]]) c = 0;
            child[k] = c;
            any |= c != 0;
        }
        if (!any) break;
        unsigned bestCnt = minSub, nBest = 0;
        int32_t bestChild = 0;
        for (uint32_t k = 0; k < nt; ++k) {
            const int32_t c = child[k];
            if (c == 0) continue;
            bool first = true;
            for (uint32_t m = 0; m < k; ++m) if (child[m] == c) { first = false; break; }
            if (!first) continue;
            unsigned cnt = 0;
            for (uint32_t m = k; m < nt; ++m) if (child[m] == c) cnt += (unsigned)pairs[2 * m + 1];
            if (cnt > bestCnt) { bestCnt = cnt; nBest = 1; bestChild = c; }
            else if (cnt == bestCnt) { if (nBest == 0) bestChild = c; ++nBest; }
        }
        if (nBest == 1) root = bestChild; else break;
    }
    res.classification = root;
    a.results[r] = res;
}

}  // namespace mbl
