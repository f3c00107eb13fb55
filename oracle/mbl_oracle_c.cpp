// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the CPU restatement (mbl_oracle.cpp) so
// that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg can drive it
// through ctypes with the same array layouts as the product's C-ABI.  Never linked into the product.
#include <chrono>
#include <cstring>
#include <fstream>

#include "mbl_oracle.hpp"

using namespace orc;

namespace {
struct ResultRec {      // same layout as mbl_read_result (include/metabuli_b200.h)
    int32_t classification; float score; int32_t hamming; int32_t query_length;
    uint32_t taxcnt_begin, taxcnt_len; uint8_t is_classified; uint8_t pad[3];
};
static_assert(sizeof(ResultRec) == 28, "layout");

void make_reads(const uint8_t* bases, const uint64_t* off, uint32_t n, std::vector<Read>& out) {
    out.resize(n);
    for (uint32_t i = 0; i < n; ++i) out[i].seq.assign((const char*)bases + off[i], (size_t)(off[i + 1] - off[i]));
}
void set_err(char* err, size_t len, const std::string& s) { if (err && len) { strncpy(err, s.c_str(), len - 1); err[len - 1] = 0; } }
}  // namespace

extern "C" {

void* orc_db_open(const char* dir, char* err, size_t errlen) {
    Database* db = new Database();
    std::string e;
    if (!db->load(dir, &e)) { set_err(err, errlen, e); delete db; return nullptr; }
    return db;
}
// in-memory variant (bench.py builds multi-GiB synthetic indexes without touching the disk)
void* orc_db_from_arrays(const uint16_t* diff, size_t n_u16, const int32_t* info, size_t n_kmers, const uint64_t* split, size_t n_split,
                         const char* taxonomy_blob, size_t blob_size, const int32_t* taxid_list, size_t n_list, int kmer_format,
                         int skip_redundancy, char* err, size_t errlen) {
    Database* db = new Database();
    std::string e;
    db->params.kmerFormat = kmer_format; db->params.skipRedundancy = skip_redundancy;
    db->diffIdx.assign(diff, diff + n_u16);
    db->info.assign(info, info + n_kmers);
    db->split.resize(n_split);
    memcpy(db->split.data(), split, n_split * 24);
    if (!db->tax.load_blob(taxonomy_blob, blob_size, &e)) { set_err(err, errlen, e); delete db; return nullptr; }
    build_taxid2species(db->tax, std::vector<int32_t>(taxid_list, taxid_list + n_list), db->taxid2species);
    return db;
}
void orc_db_close(void* db) { delete (Database*)db; }
// Syncmer / S-mer_len of db.parameters for databases built from arrays (common.cpp:117-125)
void orc_db_set_syncmer(void* db, int syncmer, int smer_len) { ((Database*)db)->params.syncmer = syncmer; ((Database*)db)->params.smerLen = smer_len; }
int orc_db_syncmer(void* db) { return ((Database*)db)->params.syncmer ? ((Database*)db)->params.smerLen : 0; }
int orc_db_kmer_format(void* db) { return ((Database*)db)->params.kmerFormat; }

// A0-A3': returns the number of slots (reference reservation order, blanks all-zero)
int orc_extract2(const uint8_t* bases1, const uint64_t* off1, const uint8_t* bases2, const uint64_t* off2, uint32_t n,
                 int kmer_format, int syncmer, int smer_len, uint64_t* value, uint64_t* qinfo, size_t cap, size_t* n_out, int32_t* cov1,
                 int32_t* cov2);
int orc_extract(const uint8_t* bases1, const uint64_t* off1, const uint8_t* bases2, const uint64_t* off2, uint32_t n,
                int kmer_format, uint64_t* value, uint64_t* qinfo, size_t cap, size_t* n_out, int32_t* cov1, int32_t* cov2) {
    return orc_extract2(bases1, off1, bases2, off2, n, kmer_format, 0, 5, value, qinfo, cap, n_out, cov1, cov2);
}
int orc_extract2(const uint8_t* bases1, const uint64_t* off1, const uint8_t* bases2, const uint64_t* off2, uint32_t n,
                 int kmer_format, int syncmer, int smer_len, uint64_t* value, uint64_t* qinfo, size_t cap, size_t* n_out, int32_t* cov1,
                 int32_t* cov2) {
    std::vector<Read> m1, m2;
    make_reads(bases1, off1, n, m1);
    if (bases2) make_reads(bases2, off2, n, m2);
    std::vector<QueryInfo> q;
    std::vector<Kmer> k;
    extract_kmers(m1, bases2 ? &m2 : nullptr, kmer_format, q, k, syncmer, smer_len);
    *n_out = k.size();
    for (uint32_t i = 0; i < n; ++i) { if (cov1) cov1[i] = q[i].queryLength; if (cov2) cov2[i] = q[i].queryLength2; }
    if (k.size() > cap) return 2;
    for (size_t i = 0; i < k.size(); ++i) { value[i] = k[i].value; qinfo[i] = k[i].qinfo; }
    return 0;
}

void orc_sort_kmers(uint64_t* value, uint64_t* qinfo, size_t n, int threads) {
    std::vector<Kmer> k(n);
    for (size_t i = 0; i < n; ++i) k[i] = Kmer{value[i], qinfo[i]};
    sort_kmers(k, threads);
    for (size_t i = 0; i < n; ++i) { value[i] = k[i].value; qinfo[i] = k[i].qinfo; }
}

// A5-A8: kmers sorted by (value, seqID), blanks (seqID 0) first
int orc_match(void* dbp, const uint64_t* value, const uint64_t* qinfo, size_t n, void* out24, size_t cap, size_t* n_match, int threads) {
    std::vector<Kmer> k(n);
    for (size_t i = 0; i < n; ++i) k[i] = Kmer{value[i], qinfo[i]};
    std::vector<Match> m;
    std::string err;
    if (!match_kmers(*(Database*)dbp, k, m, threads, &err)) return -4;
    *n_match = m.size();
    if (m.size() > cap) return 1;
    if (!m.empty()) memcpy(out24, m.data(), sizeof(Match) * m.size());
    return 0;
}

void orc_sort_matches(void* m24, size_t n, int threads) {
    std::vector<Match> m((Match*)m24, (Match*)m24 + n);
    sort_matches(m, threads);
    if (n) memcpy(m24, m.data(), sizeof(Match) * n);
}

int orc_score(void* dbp, int seq_mode, float min_score, float min_sp_score, float tie_ratio, int min_cons, int min_cons_euk,
              int accession_level, const void* m24, size_t n_match, uint32_t n_reads, const int32_t* cov1, const int32_t* cov2,
              void* results, int32_t* pairs, size_t cap_pairs, size_t* used_pairs, int threads) {
    Options opt;
    opt.seqMode = seq_mode; opt.minScore = min_score; opt.minSpScore = min_sp_score; opt.tieRatio = tie_ratio;
    opt.minConsCnt = min_cons; opt.minConsCntEuk = min_cons_euk; opt.accessionLevel = accession_level;
    std::vector<Match> m((const Match*)m24, (const Match*)m24 + n_match);
    std::vector<QueryInfo> q(n_reads);
    for (uint32_t i = 0; i < n_reads; ++i) { q[i].queryLength = cov1[i]; q[i].queryLength2 = cov2 ? cov2[i] : 0; }
    score_reads(*(Database*)dbp, opt, m, q, threads);
    ResultRec* out = (ResultRec*)results;
    size_t used = 0;
    for (uint32_t i = 0; i < n_reads; ++i) used += q[i].taxCnt.size();
    *used_pairs = used;
    if (used > cap_pairs) return 2;
    used = 0;
    for (uint32_t i = 0; i < n_reads; ++i) {
        ResultRec r{};
        r.classification = q[i].classification; r.score = q[i].score; r.hamming = q[i].hammingDist;
        r.query_length = q[i].queryLength + q[i].queryLength2; r.is_classified = q[i].isClassified ? 1 : 0;
        r.taxcnt_begin = (uint32_t)used; r.taxcnt_len = (uint32_t)q[i].taxCnt.size();
        for (auto& t : q[i].taxCnt) { pairs[2 * used] = t.first; pairs[2 * used + 1] = t.second; ++used; }
        out[i] = r;
    }
    return 0;
}

// the whole path over SoA reads, timed (bench.py cpu_baseline / --impl reference).  Returns seconds, <0 on error.
double orc_classify_arrays(void* dbp, int seq_mode, int threads, const uint8_t* bases1, const uint64_t* off1,
                           const uint8_t* bases2, const uint64_t* off2, uint32_t n, void* results, size_t* n_kmers, size_t* n_matches) {
    Database& db = *(Database*)dbp;
    Options opt;
    opt.seqMode = seq_mode; opt.threads = threads;
    std::vector<Read> m1, m2;
    make_reads(bases1, off1, n, m1);
    if (bases2) make_reads(bases2, off2, n, m2);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<QueryInfo> q;
    std::vector<Kmer> k;
    extract_kmers(m1, bases2 ? &m2 : nullptr, db.params.kmerFormat, q, k, db.params.syncmer, db.params.smerLen);
    sort_kmers(k, threads);
    std::vector<Match> m;
    std::string err;
    if (!match_kmers(db, k, m, threads, &err)) return -1.0;
    sort_matches(m, threads);
    score_reads(db, opt, m, q, threads);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (n_kmers) *n_kmers = k.size();
    if (n_matches) *n_matches = m.size();
    if (results) {
        ResultRec* out = (ResultRec*)results;
        for (uint32_t i = 0; i < n; ++i) {
            ResultRec r{};
            r.classification = q[i].classification; r.score = q[i].score; r.query_length = q[i].queryLength + q[i].queryLength2;
            r.is_classified = q[i].isClassified ? 1 : 0; r.taxcnt_len = (uint32_t)q[i].taxCnt.size();
            out[i] = r;
        }
    }
    return sec;
}

// classify flags for orc_classify_files (classify.cpp:10-37 defaults until set): --min-score, --min-sp-score, --tie-ratio,
// --min-cons-cnt, --min-cons-cnt-euk, --accession-level
static Options g_flags;
void orc_set_flags(float min_score, float min_sp_score, float tie_ratio, int min_cons, int min_cons_euk, int accession_level) {
    g_flags.minScore = min_score; g_flags.minSpScore = min_sp_score; g_flags.tieRatio = tie_ratio;
    g_flags.minConsCnt = min_cons; g_flags.minConsCntEuk = min_cons_euk; g_flags.accessionLevel = accession_level;
}
void orc_set_lineage(int print_lineage) { g_flags.printLineage = print_lineage; }

int orc_classify_files(const char* q1, const char* q2, const char* db_dir, int seq_mode, int threads, const char* out_path,
                       size_t* n_kmers, size_t* n_matches, char* err, size_t errlen) {
    Options opt = g_flags;
    opt.seqMode = seq_mode; opt.threads = threads;
    std::string tsv, e, report;
    if (!classify_files(q1, q2 ? q2 : "", db_dir, opt, tsv, &e, n_kmers, n_matches, &report)) { set_err(err, errlen, e); return -1; }
    std::ofstream f(out_path, std::ios::binary);
    f << tsv;
    std::ofstream r(std::string(out_path) + ".report", std::ios::binary);      // <out_path>.report = text of <jobid>_report.tsv
    r << report;
    return f.good() && r.good() ? 0 : -1;
}

// primitives for unit tests
uint64_t orc_next_target_kmer(uint64_t prev, const uint16_t* diff, size_t* idx) { return next_target_kmer(prev, diff, *idx); }
uint8_t orc_hamming_sum(uint64_t a, uint64_t b) { return hamming_sum(a, b); }
uint16_t orc_hammings_fwd(uint64_t a, uint64_t b) { return hammings_fwd(a, b); }
uint16_t orc_hammings_rev(uint64_t a, uint64_t b) { return hammings_rev(a, b); }
int orc_max_covered_length(int len) { return max_covered_length(len); }
int orc_query_kmer_number(int len) { return query_kmer_number(len); }

}  // extern "C"
