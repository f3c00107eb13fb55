// TEST DOUBLE of libmetabuli_b200.so for the CPU suite: the C-ABI entry points the C++ host (`metabuli-b200 classify`) calls in
// its replica and index-sharded modes, answered by the oracle (oracle/mbl_oracle.cpp) instead of the GPU.  tests/test_cli_cpu.py puts this library
// next to a copy of the CLI binary (its rpath is $ORIGIN), so that the host's own code — streaming reader, batch pipeline, masking
// flags, row formatter, report — runs end to end without a device and is compared with the reference binary's files.  It is test
// infrastructure like the oracle: nothing under metabuli_b200/ links or loads it, and the real library never falls back to it.
// The database is read by the oracle from MBL_STUB_DB_DIR (the arrays the host passes to mbl_load_db are only size-checked).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/metabuli_b200.h"
#include "../../metabuli_b200/csrc/host/tantan_mask.hpp"
#include "../../oracle/mbl_oracle.hpp"

struct mbl_ctx {
    mbl_config cfg{};
    std::string err;
    orc::Database db;
    bool loaded = false;
    // staged batch (copied: the double keeps no pointers into the caller's buffers)
    std::vector<orc::Read> m1, m2;
    bool staged = false, paired = false;
    uint32_t staged_n = 0;
    // index-sharded phases: what this rank extracted / matched, its receive buffers (plain host memory here) and its peers'
    uint64_t seq_base = 0;
    std::vector<orc::QueryInfo> q;
    std::vector<std::vector<orc::Kmer>> kbucket;
    std::vector<std::vector<orc::Match>> mbucket;
    std::vector<uint64_t> recv_k;
    std::vector<orc::Match> recv_m;
    std::vector<uint64_t*> peer_k;
    std::vector<orc::Match*> peer_m;
    // results of the last classified batch
    std::vector<mbl_read_result> res;
    std::vector<int32_t> pairs;
    mbl_stats stats{};
};

namespace {
int fail(mbl_ctx* c, int rc, const std::string& m) { if (c) c->err = m; return rc; }

// MBL_STUB_NULL=1: no classification at all (every read comes back unclassified at once) — what is left is the host's own
// pipeline, so the CLI's wall time is its host-side ceiling (tools/host_ceiling.sh)
bool null_mode() { static const bool on = getenv("MBL_STUB_NULL") != nullptr; return on; }

void stage(mbl_ctx* c, const mbl_batch* b) {
    if (null_mode()) { c->staged_n = b->n_reads; c->staged = true; return; }
    auto fill = [&](const char* bases, const uint64_t* off, std::vector<orc::Read>& out) {
        out.assign(b->n_reads, orc::Read());
        std::string all(bases, bases + off[b->n_reads]);
        if (c->cfg.mask_mode) mblhost::tantan_mask_reads(&all[0], off, b->n_reads, c->cfg.mask_prob, 2);   // what K0 does on the device
        for (uint32_t i = 0; i < b->n_reads; ++i) out[i].seq.assign(all, off[i], off[i + 1] - off[i]);
    };
    fill(b->bases, b->offsets, c->m1);
    c->paired = b->bases2 && b->offsets2;
    if (c->paired) fill(b->bases2, b->offsets2, c->m2); else c->m2.clear();
    c->staged = true;
}

orc::Options options_of(const mbl_ctx* c) {
    orc::Options opt;
    opt.seqMode = c->cfg.seq_mode; opt.minScore = c->cfg.min_score; opt.minSpScore = c->cfg.min_sp_score; opt.tieRatio = c->cfg.tie_ratio;
    opt.minConsCnt = c->cfg.min_cons_cnt; opt.minConsCntEuk = c->cfg.min_cons_cnt_euk; opt.accessionLevel = c->cfg.accession_level;
    return opt;
}

void fill_results(mbl_ctx* c, const std::vector<orc::QueryInfo>& q) {
    c->res.assign(q.size(), mbl_read_result{});
    c->pairs.clear();
    for (size_t i = 0; i < q.size(); ++i) {
        mbl_read_result& r = c->res[i];
        r.classification = q[i].classification; r.score = q[i].score; r.hamming = q[i].hammingDist;
        r.query_length = q[i].queryLength + q[i].queryLength2; r.is_classified = q[i].isClassified ? 1 : 0;
        r.taxcnt_begin = (uint32_t)(c->pairs.size() / 2); r.taxcnt_len = (uint32_t)q[i].taxCnt.size();
        for (auto& t : q[i].taxCnt) { c->pairs.push_back(t.first); c->pairs.push_back(t.second); }
    }
}

int classify_staged(mbl_ctx* c) {
    if (null_mode()) { c->res.assign(c->staged_n, mbl_read_result{}); c->pairs.clear(); c->stats = mbl_stats{}; c->staged = false; return MBL_OK; }
    const orc::Options opt = options_of(c);
    std::vector<orc::QueryInfo> q;
    std::vector<orc::Kmer> k;
    orc::extract_kmers(c->m1, c->paired ? &c->m2 : nullptr, c->db.params.kmerFormat, q, k, c->db.params.syncmer, c->db.params.smerLen);
    orc::sort_kmers(k, 2);
    std::vector<orc::Match> m;
    std::string err;
    if (!orc::match_kmers(c->db, k, m, 2, &err)) return fail(c, MBL_E_BAD_DB, err);
    orc::sort_matches(m, 2);
    orc::score_reads(c->db, opt, m, q, 2);
    fill_results(c, q);
    c->stats = mbl_stats{};
    for (auto& x : k) if ((x.qinfo >> 32) & 0x1FFFFFFFu) ++c->stats.n_query_kmers;
    c->stats.n_matches = m.size();
    c->staged = false;
    return MBL_OK;
}
}  // namespace

extern "C" {

int mbl_create(const mbl_config* cfg, mbl_ctx** out) {
    if (!cfg || !out) return MBL_E_BAD_ARG;
    *out = new mbl_ctx();
    (*out)->cfg = *cfg;
    return MBL_OK;
}
void mbl_destroy(mbl_ctx* c) { delete c; }
const char* mbl_last_error(const mbl_ctx* c) { return c ? c->err.c_str() : "null context"; }

int mbl_load_db(mbl_ctx* c, const mbl_db* db, const mbl_taxonomy* tx) {
    if (!c || !db || !tx) return MBL_E_BAD_ARG;
    const char* dir = getenv("MBL_STUB_DB_DIR");
    std::string err;
    if (!dir || !c->db.load(dir, &err)) return fail(c, MBL_E_BAD_DB, "stub backend: MBL_STUB_DB_DIR: " + err);
    if (c->db.diffIdx.size() != db->n_u16 || c->db.info.size() != db->n_kmers || c->db.tax.maxNodes != tx->max_nodes)
        return fail(c, MBL_E_BAD_DB, "stub backend: the host passed other arrays than MBL_STUB_DB_DIR holds");
    c->loaded = true;
    return MBL_OK;
}

int mbl_download_results(mbl_ctx* c, mbl_read_result* out, int32_t* taxcnt_pairs, size_t cap_pairs, size_t* used_pairs) {
    if (!c || !used_pairs) return MBL_E_BAD_ARG;
    *used_pairs = c->pairs.size() / 2;
    if (*used_pairs > cap_pairs) return fail(c, MBL_E_CAPACITY, "pair buffer too small");
    if (!c->res.empty()) memcpy(out, c->res.data(), c->res.size() * sizeof(mbl_read_result));
    if (!c->pairs.empty()) memcpy(taxcnt_pairs, c->pairs.data(), c->pairs.size() * 4);
    return MBL_OK;
}

int mbl_classify_batch(mbl_ctx* c, const mbl_batch* b, mbl_read_result* out, int32_t* taxcnt_pairs, size_t cap_pairs, size_t* used_pairs) {
    if (!c || !b) return MBL_E_BAD_ARG;
    stage(c, b);
    const int rc = classify_staged(c);
    return rc != MBL_OK ? rc : mbl_download_results(c, out, taxcnt_pairs, cap_pairs, used_pairs);
}

int mbl_prefetch_batch(mbl_ctx* c, const mbl_batch* b) {
    if (!c || !b || !b->bases || !b->offsets) return fail(c, MBL_E_BAD_ARG, "null argument");
    stage(c, b);
    return MBL_OK;
}

int mbl_classify_prefetched(mbl_ctx* c, const mbl_batch* next, mbl_read_result* out, int32_t* taxcnt_pairs, size_t cap_pairs, size_t* used_pairs) {
    if (!c) return MBL_E_BAD_ARG;
    if (!c->staged) return fail(c, MBL_E_BAD_ARG, "mbl_prefetch_batch has not been called");
    int rc = classify_staged(c);
    if (rc != MBL_OK) return rc;
    if (next) stage(c, next);                       // results of the current batch are already in c->res / c->pairs
    return mbl_download_results(c, out, taxcnt_pairs, cap_pairs, used_pairs);
}

int mbl_get_stats(const mbl_ctx* c, mbl_stats* out) { if (!c || !out) return MBL_E_BAD_ARG; *out = c->stats; return MBL_OK; }
int mbl_host_register(void*, size_t) { return MBL_OK; }
int mbl_host_unregister(void*) { return MBL_OK; }

int mbl_mask_reads(char* bases, const uint64_t* offsets, uint32_t n_reads, float mask_prob, int threads) {
    if (!offsets) return MBL_E_BAD_ARG;
    mblhost::tantan_mask_reads(bases, offsets, n_reads, mask_prob, threads > 0 ? (unsigned)threads : 2u);
    return MBL_OK;
}

// ---- index-sharded phases on host memory: "device pointers" are plain pointers, every rank keeps the whole database (a metamer
// still goes to exactly one rank, by value range, so the matches are the same), the exchanges are the pushes below -----------------
int mbl_plan_shards(const mbl_db* db, uint32_t n_shards, mbl_shard* out) {
    if (!db || !out || n_shards < 1 || n_shards > MBL_MAX_SHARDS) return MBL_E_BAD_ARG;
    for (uint32_t s = 0; s < n_shards; ++s) {
        out[s] = mbl_shard{};
        out[s].first_value = s == 0 ? 0 : ((~0ull / n_shards) * s) & ~0xFFFFFFull;       // amino-acid-group aligned
    }
    return MBL_OK;
}
int mbl_load_db_shard(mbl_ctx* c, const mbl_db* db, const mbl_taxonomy* tx, const mbl_shard*) { return mbl_load_db(c, db, tx); }
int mbl_shard_filter(mbl_ctx* c, void** d_words, uint64_t* n_bytes) { if (!c || !d_words || !n_bytes) return MBL_E_BAD_ARG; *d_words = nullptr; *n_bytes = 0; return MBL_OK; }
int mbl_shard_filter_or(mbl_ctx*, const void*, uint64_t, int) { return MBL_OK; }

int mbl_shard_extract(mbl_ctx* c, const mbl_batch* b, uint64_t seq_base, uint32_t n_shards, const uint64_t* first, uint64_t* send_counts) {
    if (!c || !b || !first || !send_counts) return MBL_E_BAD_ARG;
    stage(c, b);
    c->staged = false;
    c->seq_base = seq_base;
    std::vector<orc::Kmer> k;
    orc::extract_kmers(c->m1, c->paired ? &c->m2 : nullptr, c->db.params.kmerFormat, c->q, k, c->db.params.syncmer, c->db.params.smerLen);
    c->kbucket.assign(n_shards, {});
    c->stats = mbl_stats{};
    for (const orc::Kmer& x : k) {
        if (!((x.qinfo >> 32) & 0x1FFFFFFFu)) continue;                      // blank slot
        ++c->stats.n_query_kmers;
        uint32_t s = 0;
        while (s + 1 < n_shards && first[s + 1] <= x.value) ++s;
        c->kbucket[s].push_back(orc::Kmer{x.value, x.qinfo + (seq_base << 32)});   // seqIDs are global on the wire
    }
    for (uint32_t s = 0; s < n_shards; ++s) send_counts[s] = c->kbucket[s].size();
    return MBL_OK;
}

int mbl_shard_recv_buffers(mbl_ctx* c, uint64_t kmer_rows, uint64_t match_rows, void** d_kmers, void** d_matches, uint8_t*, uint8_t*) {
    if (!c || !d_kmers || !d_matches) return MBL_E_BAD_ARG;
    c->recv_k.assign(2 * kmer_rows + 2, 0);
    c->recv_m.assign(match_rows + 1, orc::Match{});
    *d_kmers = c->recv_k.data(); *d_matches = c->recv_m.data();
    return MBL_OK;
}
int mbl_shard_attach_peer(mbl_ctx* c, uint32_t peer, const uint8_t*, const uint8_t*, void* raw_kmers, void* raw_matches) {
    if (!c || peer >= MBL_MAX_SHARDS) return MBL_E_BAD_ARG;
    if (c->peer_k.size() <= peer) { c->peer_k.resize(peer + 1, nullptr); c->peer_m.resize(peer + 1, nullptr); }
    c->peer_k[peer] = static_cast<uint64_t*>(raw_kmers); c->peer_m[peer] = static_cast<orc::Match*>(raw_matches);
    return MBL_OK;
}
int mbl_shard_detach_peers(mbl_ctx* c) { if (!c) return MBL_E_BAD_ARG; c->peer_k.clear(); c->peer_m.clear(); return MBL_OK; }

int mbl_shard_push_kmers(mbl_ctx* c, const uint64_t* off, const uint64_t* tot) {
    if (!c || !off || !tot) return MBL_E_BAD_ARG;
    for (size_t h = 0; h < c->kbucket.size(); ++h) {
        if (c->kbucket[h].empty()) continue;
        if (h >= c->peer_k.size() || !c->peer_k[h]) return fail(c, MBL_E_BAD_ARG, "peer not attached");
        uint64_t* buf = c->peer_k[h];
        for (size_t i = 0; i < c->kbucket[h].size(); ++i) { buf[off[h] + i] = c->kbucket[h][i].value; buf[tot[h] + off[h] + i] = c->kbucket[h][i].qinfo; }
    }
    return MBL_OK;
}

int mbl_shard_match(mbl_ctx* c, const uint64_t* value, const uint64_t* qinfo, uint64_t n, uint32_t n_owners, const uint64_t* owner_first_read,
                    uint64_t* send_counts) {
    if (!c || !owner_first_read || !send_counts || (n && (!value || !qinfo))) return MBL_E_BAD_ARG;
    std::vector<orc::Kmer> k(n);
    for (uint64_t i = 0; i < n; ++i) k[i] = orc::Kmer{value[i], qinfo[i]};
    orc::sort_kmers(k, 2);
    std::vector<orc::Match> m;
    std::string err;
    if (!orc::match_kmers(c->db, k, m, 2, &err)) return fail(c, MBL_E_BAD_DB, err);
    c->mbucket.assign(n_owners, {});
    for (const orc::Match& x : m) {
        const uint64_t read = ((x.qinfo >> 32) & 0x1FFFFFFFu) - 1;            // global read index
        uint32_t o = 0;
        while (o + 1 < n_owners && owner_first_read[o + 1] <= read) ++o;
        c->mbucket[o].push_back(x);
    }
    for (uint32_t o = 0; o < n_owners; ++o) send_counts[o] = c->mbucket[o].size();
    c->stats.n_matches = m.size();
    return MBL_OK;
}

int mbl_shard_push_matches(mbl_ctx* c, const uint64_t* off) {
    if (!c || !off) return MBL_E_BAD_ARG;
    for (size_t o = 0; o < c->mbucket.size(); ++o) {
        if (c->mbucket[o].empty()) continue;
        if (o >= c->peer_m.size() || !c->peer_m[o]) return fail(c, MBL_E_BAD_ARG, "peer not attached");
        memcpy(c->peer_m[o] + off[o], c->mbucket[o].data(), c->mbucket[o].size() * sizeof(orc::Match));
    }
    return MBL_OK;
}

int mbl_shard_score(mbl_ctx* c, const mbl_match_rec* d_match, uint64_t n_match) {
    if (!c || (n_match && !d_match)) return MBL_E_BAD_ARG;
    static_assert(sizeof(orc::Match) == sizeof(mbl_match_rec), "match record layout");
    std::vector<orc::Match> m(n_match);
    if (n_match) memcpy(m.data(), d_match, n_match * sizeof(orc::Match));
    for (orc::Match& x : m) x.qinfo -= c->seq_base << 32;                    // back to the batch's own read numbers
    orc::sort_matches(m, 2);
    orc::score_reads(c->db, options_of(c), m, c->q, 2);
    fill_results(c, c->q);
    return MBL_OK;
}

}  // extern "C"
