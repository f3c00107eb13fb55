// TEST DOUBLE of libmetabuli_b200.so for the CPU suite: the C-ABI entry points the C++ host (`metabuli-b200 classify`) calls in
// its replica mode, answered by the oracle (oracle/mbl_oracle.cpp) instead of the GPU.  tests/test_cli_cpu.py puts this library
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

int classify_staged(mbl_ctx* c) {
    if (null_mode()) { c->res.assign(c->staged_n, mbl_read_result{}); c->pairs.clear(); c->stats = mbl_stats{}; c->staged = false; return MBL_OK; }
    orc::Options opt;
    opt.seqMode = c->cfg.seq_mode; opt.minScore = c->cfg.min_score; opt.minSpScore = c->cfg.min_sp_score; opt.tieRatio = c->cfg.tie_ratio;
    opt.minConsCnt = c->cfg.min_cons_cnt; opt.minConsCntEuk = c->cfg.min_cons_cnt_euk; opt.accessionLevel = c->cfg.accession_level;
    std::vector<orc::QueryInfo> q;
    std::vector<orc::Kmer> k;
    orc::extract_kmers(c->m1, c->paired ? &c->m2 : nullptr, c->db.params.kmerFormat, q, k, c->db.params.syncmer, c->db.params.smerLen);
    orc::sort_kmers(k, 2);
    std::vector<orc::Match> m;
    std::string err;
    if (!orc::match_kmers(c->db, k, m, 2, &err)) return fail(c, MBL_E_BAD_DB, err);
    orc::sort_matches(m, 2);
    orc::score_reads(c->db, opt, m, q, 2);
    c->res.assign(q.size(), mbl_read_result{});
    c->pairs.clear();
    for (size_t i = 0; i < q.size(); ++i) {
        mbl_read_result& r = c->res[i];
        r.classification = q[i].classification; r.score = q[i].score; r.hamming = q[i].hammingDist;
        r.query_length = q[i].queryLength + q[i].queryLength2; r.is_classified = q[i].isClassified ? 1 : 0;
        r.taxcnt_begin = (uint32_t)(c->pairs.size() / 2); r.taxcnt_len = (uint32_t)q[i].taxCnt.size();
        for (auto& t : q[i].taxCnt) { c->pairs.push_back(t.first); c->pairs.push_back(t.second); }
    }
    c->stats = mbl_stats{};
    for (auto& x : k) if ((x.qinfo >> 32) & 0x1FFFFFFFu) ++c->stats.n_query_kmers;
    c->stats.n_matches = m.size();
    c->stats.kernel_launches = 0;
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

// the index-sharded phases need devices: the double refuses them
#define MBL_STUB_UNSUPPORTED(name, ...) int name(__VA_ARGS__) { return MBL_E_UNSUPPORTED; }
MBL_STUB_UNSUPPORTED(mbl_load_db_shard, mbl_ctx*, const mbl_db*, const mbl_taxonomy*, const mbl_shard*)
MBL_STUB_UNSUPPORTED(mbl_plan_shards, const mbl_db*, uint32_t, mbl_shard*)
MBL_STUB_UNSUPPORTED(mbl_shard_filter, mbl_ctx*, void**, uint64_t*)
MBL_STUB_UNSUPPORTED(mbl_shard_filter_or, mbl_ctx*, const void*, uint64_t, int)
MBL_STUB_UNSUPPORTED(mbl_shard_extract, mbl_ctx*, const mbl_batch*, uint64_t, uint32_t, const uint64_t*, uint64_t*)
MBL_STUB_UNSUPPORTED(mbl_shard_match, mbl_ctx*, const uint64_t*, const uint64_t*, uint64_t, uint32_t, const uint64_t*, uint64_t*)
MBL_STUB_UNSUPPORTED(mbl_shard_recv_buffers, mbl_ctx*, uint64_t, uint64_t, void**, void**, uint8_t*, uint8_t*)
MBL_STUB_UNSUPPORTED(mbl_shard_attach_peer, mbl_ctx*, uint32_t, const uint8_t*, const uint8_t*, void*, void*)
MBL_STUB_UNSUPPORTED(mbl_shard_detach_peers, mbl_ctx*)
MBL_STUB_UNSUPPORTED(mbl_shard_push_kmers, mbl_ctx*, const uint64_t*, const uint64_t*)
MBL_STUB_UNSUPPORTED(mbl_shard_push_matches, mbl_ctx*, const uint64_t*)
MBL_STUB_UNSUPPORTED(mbl_shard_score, mbl_ctx*, const mbl_match_rec*, uint64_t)

}  // extern "C"
