// Host-side build of the product's scoring core (score_core.cuh is __host__ __device__) so that the
// "not gpu" tests can check the host logic: the libstdc++-order sort replay against std::sort, and the
// whole per-read scoring against the oracle, without a GPU.  Test helper only; not part of the product.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../metabuli_b200/csrc/score_core.cuh"

using namespace mbl;

static int g_smer_len = 0;     // syncmer database: s-mer length (0 = none); set with ht_set_syncmer
extern "C" void ht_set_syncmer(int smer_len) { g_smer_len = smer_len; }

extern "C" {

// comparator of combineMatchPaths: (score desc, hamming asc, start desc)
void ht_sort_paths(const float* score, const int* ham, const int* start, int n, int* perm_replay, int* perm_std) {
    for (int i = 0; i < n; ++i) perm_replay[i] = perm_std[i] = i;
    auto less = [&](int32_t x, int32_t y) {
        if (score[x] != score[y]) return score[x] > score[y];
        if (ham[x] != ham[y]) return ham[x] < ham[y];
        return start[x] > start[y];
    };
    stl_sort(perm_replay, n, less);
    std::sort(perm_std, perm_std + n, less);
}

int ht_score(const mbl_match_rec* matches, size_t n_match, uint32_t n_reads, const int32_t* cov1, const int32_t* cov2,
             const mbl_taxonomy* tx, int seq_mode, float min_score, float min_sp_score, float tie_ratio, int min_cons, int min_cons_euk,
             int accession_level, int kmer_format, int force_scratch_dp, int flat, mbl_read_result* results, int32_t* pairs_out, size_t cap_pairs, size_t* used_pairs) {
    std::vector<uint64_t> seg_b(n_reads, 0), seg_e(n_reads, 0);
    for (size_t i = 0; i < n_match; ++i) {
        uint32_t s = qi_seq(matches[i].qinfo);
        if (s == 0 || s > n_reads) continue;
        if (i == 0 || qi_seq(matches[i - 1].qinfo) != s) seg_b[s - 1] = i;
        if (i + 1 == n_match || qi_seq(matches[i + 1].qinfo) != s) seg_e[s - 1] = i + 1;
    }
    std::vector<uint32_t> quot_off(n_reads + 1, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        int ql = cov1[r] + (cov2 ? cov2[r] : 0);
        quot_off[r + 1] = quot_off[r] + (ql + 3 > 0 ? (uint32_t)((ql + 3) / 3 + 1) : 1u);
    }
    std::vector<int32_t> zero(n_reads, 0);
    const size_t M = n_match + 1, Q = quot_off[n_reads] + 1;
    std::vector<float> l_score(M), p_score(M), s_score(M);
    std::vector<int32_t> l_start(M), l_ham(M), l_depth(M), p_start(M), p_end(M), p_ham(M), p_depth(M), c_start(M), c_end(M), q_tax(Q), praw(2 * Q);
    std::vector<uint32_t> l_smatch(M), p_smatch(M), p_ematch(M);
    std::vector<uint8_t> l_conn(M), q_ham(Q), q_has(Q);
    ScoreArgs a{};
    a.matches = matches; a.n_match = n_match; a.n_reads = n_reads; a.seg_begin = seg_b.data(); a.seg_end = seg_e.data();
    a.cov1 = cov1; a.cov2 = cov2 ? cov2 : zero.data(); a.quot_off = quot_off.data();
    a.tax.D = tx->D; a.tax.E = tx->E; a.tax.L = tx->L; a.tax.H = tx->H; a.tax.M = tx->M; a.tax.node_taxid = tx->node_taxid;
    a.tax.node_parent = tx->node_parent; a.tax.node_prune = tx->node_prune; a.tax.node_rank = tx->node_rank;
    a.tax.taxid2species = tx->taxid2species; a.tax.max_taxid = tx->max_taxid; a.tax.M_k = tx->M_k; a.tax.eukaryota = tx->eukaryota;
    a.tax.max_nodes = (uint32_t)tx->max_nodes;
    a.par.min_score = min_score; a.par.min_sp_score = min_sp_score; a.par.tie_ratio = tie_ratio; a.par.min_cons_cnt = min_cons;
    a.par.min_cons_cnt_euk = min_cons_euk; a.par.accession_level = accession_level;
    a.par.denominator = (seq_mode == 1 || seq_mode == 2) ? 100 : 1000; a.par.kmer_format = kmer_format; a.par.force_scratch_dp = force_scratch_dp;
    a.par.max_codon_shift = g_smer_len ? 8 - g_smer_len : 1; a.par.dna_shift = 3 * a.par.max_codon_shift;   // Taxonomer.cpp:34-42
    a.l_score = l_score.data(); a.l_start = l_start.data(); a.l_ham = l_ham.data(); a.l_depth = l_depth.data();
    a.l_smatch = l_smatch.data(); a.l_conn = l_conn.data(); a.p_start = p_start.data(); a.p_end = p_end.data();
    a.p_score = p_score.data(); a.p_ham = p_ham.data(); a.p_depth = p_depth.data(); a.p_smatch = p_smatch.data();
    a.p_ematch = p_ematch.data(); a.c_start = c_start.data(); a.c_end = c_end.data(); a.s_score = s_score.data();
    a.q_tax = q_tax.data(); a.q_ham = q_ham.data(); a.q_has = q_has.data();
    a.results = results; a.taxcnt_pairs = praw.data();
    std::vector<uint32_t> fg, sp, gnp(M, 0);
    if (flat) {                      // the three-pass task formulation the CUDA pipeline uses
        for (size_t i = 0; i < n_match; ++i) {
            bool s = i == 0 || qi_seq(matches[i].qinfo) != qi_seq(matches[i - 1].qinfo) || matches[i].species_id != matches[i - 1].species_id;
            bool f = s || qi_frame(matches[i].qinfo) != qi_frame(matches[i - 1].qinfo);
            if (s) sp.push_back((uint32_t)i);
            if (f) fg.push_back((uint32_t)i);
        }
        a.fg_list = fg.data(); a.n_fg = (uint32_t)fg.size(); a.sp_list = sp.data(); a.n_sp = (uint32_t)sp.size();
        a.g_np = gnp.data(); a.match_end = n_match;
        for (uint32_t g = 0; g < a.n_fg; ++g) score_task_frame_group(a, g);
        for (uint32_t k = 0; k < a.n_sp; ++k) score_task_species(a, k);
    }
    for (uint32_t r = 0; r < n_reads; ++r) score_read(a, r);
    size_t used = 0;
    for (uint32_t r = 0; r < n_reads; ++r) used += results[r].taxcnt_len;
    *used_pairs = used;
    if (used > cap_pairs) return 2;
    used = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        memcpy(pairs_out + 2 * used, praw.data() + 2ull * quot_off[r], 8ull * results[r].taxcnt_len);
        results[r].taxcnt_begin = (uint32_t)used;
        used += results[r].taxcnt_len;
    }
    return 0;
}

}  // extern "C"
