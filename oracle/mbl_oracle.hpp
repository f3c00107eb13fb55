// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the `metabuli classify` hot path (reference @ 22e7026). It exists so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg can check and
// time the CUDA path against the reference's algorithm.  Nothing under metabuli_b200/ may include,
// link or call it.
//
// Parity pin: this restatement is pinned to the reference by the end-to-end known answers of
// SURVEY.md §8(c) / BASELINE.md §5 (md5 of <job>_classifications.tsv on the four regression-fixture
// configurations, query-k-mer and match counts) and by the golden TSVs in tests/golden/ which were
// produced by the reference binary itself in the build container (tests/golden/gen_golden.sh).
//
// Every function cites the reference file:line it restates (paths relative to /root/reference).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace orc {

// ---- PODs (same layouts as the reference's) -------------------------------------------------
// Kmer.h:11-16  QueryKmerInfo bit-field: pos[31:0] | sequenceID[60:32] | frame[63:61]
inline uint64_t pack_qinfo(uint32_t seqId, uint32_t pos, uint32_t frame) {
    return (uint64_t)pos | ((uint64_t)(seqId & 0x1FFFFFFFu) << 32) | ((uint64_t)(frame & 7u) << 61);
}
inline uint32_t qi_pos(uint64_t q) { return (uint32_t)q; }
inline uint32_t qi_seq(uint64_t q) { return (uint32_t)((q >> 32) & 0x1FFFFFFFu); }
inline uint32_t qi_frame(uint64_t q) { return (uint32_t)(q >> 61); }

struct Kmer {            // Kmer.h:24-46 (16 B)
    uint64_t value;
    uint64_t qinfo;
};

struct Match {           // Match.h:9-26 without the vptr (24 B)
    uint64_t qinfo;
    int32_t targetId;
    int32_t speciesId;
    uint32_t dnaEncoding;
    uint16_t rightEndHamming;
    uint8_t hamming;
    uint8_t pad;
};
static_assert(sizeof(Match) == 24, "Match must be 24 bytes");

struct Split {           // Kmer.h:111-119 DiffIdxSplit (24 B)
    uint64_t adKmer;
    uint64_t diffIdxOffset;
    uint64_t infoIdxOffset;
};

// ---- Taxonomy (TaxonomyWrapper.cpp:363-421 unserialize; NcbiTaxonomy.cpp:250-330) --------------
struct Taxonomy {
    std::vector<char> blob;            // the raw taxonomyDB file
    bool internalIds = false;
    size_t maxNodes = 0;
    int32_t maxTaxID = 0;
    // per node (from TaxonNode[maxNodes], 32 B each)
    std::vector<int32_t> nodeTaxId, nodeParent;
    std::vector<uint64_t> nodeRankIdx, nodeNameIdx;
    const int32_t *D = nullptr, *internal2org = nullptr, *E = nullptr, *L = nullptr, *H = nullptr, *M = nullptr;
    int32_t Mk = 0;
    // StringBlock
    const char *strData = nullptr;
    const uint32_t *strOffsets = nullptr;
    uint32_t strCount = 0;
    int32_t eukaryota = 0;

    bool load(const std::string &path, std::string *err);
    bool load_blob(const char *data, size_t size, std::string *err);
    bool nodeExists(int32_t t) const { return t <= maxTaxID && D[t] != -1; }
    int nodeId(int32_t t) const { return D[t]; }
    const char *str(uint64_t idx) const { return strData + strOffsets[idx]; }
    int lcaHelper(int i, int j) const;
    int32_t lca(int32_t a, int32_t b) const;
    int32_t lcaMany(const std::vector<int32_t> &taxa) const;
    bool isAncestor(int32_t ancestor, int32_t child) const;
    int32_t taxIdAtRank(int32_t taxId, const char *rank) const;
    int32_t original(int32_t internal) const { return internalIds ? internal2org[internal] : internal; }
    int32_t parentOf(int32_t t) const { return nodeParent[D[t]]; }
    const char *rankOf(int32_t t) const { return str(nodeRankIdx[D[t]]); }
    std::string lineage(int32_t taxId) const;        // TaxonomyWrapper::taxLineage2
};
int rank_index(const char *rank);   // NcbiTaxonomy.h:52-80 / NcbiTaxonomy.cpp:374-380

// ---- Database --------------------------------------------------------------------------------
struct DbParams {                    // common.cpp:88-133
    int kmerFormat = 1;              // classify.cpp:13 default when db.parameters lacks Kmer_format
    int reducedAA = 0, skipRedundancy = 0, syncmer = 0, smerLen = 5, accessionLevelDb = -1;
};
struct Database {
    DbParams params;
    std::vector<uint16_t> diffIdx;
    std::vector<int32_t> info;
    std::vector<Split> split;
    Taxonomy tax;
    std::vector<int32_t> taxid2species;   // dense, KmerMatcher.cpp:56-120
    bool load(const std::string &dir, std::string *err);
};
void build_taxid2species(const Taxonomy &tax, const std::vector<int32_t> &taxIdList, std::vector<int32_t> &out);

// ---- Options (classify.cpp:10-37 defaults) ----------------------------------------------------
struct Options {
    int seqMode = 2;                 // 1 SE, 2 PE, 3 long
    float minScore = 0.f, minSpScore = 0.f, tieRatio = 0.95f;
    int minConsCnt = 4, minConsCntEuk = 9;
    int accessionLevel = 0;
    int threads = 1;
    int printLineage = 0;            // --lineage
};

// ---- Reads ----------------------------------------------------------------------------------
struct Read { std::string name, seq; };
bool read_fastx(const std::string &path, std::vector<Read> &out, std::string *err);  // kseq semantics

struct QueryInfo {                   // common.h:94-122 Query (the fields the path touches)
    int classification = 0;
    float score = 0.f;
    int hammingDist = 0;
    int queryLength = 0, queryLength2 = 0;
    int kmerCnt = 0, kmerCnt2 = 0;
    bool isClassified = false;
    std::map<int32_t, int> taxCnt;
};

// ---- Stages ----------------------------------------------------------------------------------
int max_covered_length(int len);                               // LocalUtil.h:51-60
int query_kmer_number(int len);                                // LocalUtil.h:46-49 (spaceNum 0, k 8)
// A2/A3/A3': six-frame metamer extraction of one batch.  kmers gets exactly sum(kmerCnt[+kmerCnt2])
// slots in the reference's reservation order; unused slots are all-zero (seqID 0 == blank).
// syncmer != 0: only closed syncmers with s = smerLen are emitted (SyncmerScanner.h:9-103; KmerExtractor.cpp:18-20)
void extract_kmers(const std::vector<Read> &m1, const std::vector<Read> *m2, int kmerFormat,
                   std::vector<QueryInfo> &queries, std::vector<Kmer> &kmers, int syncmer = 0, int smerLen = 5);
void sort_kmers(std::vector<Kmer> &kmers, int threads);        // A4 (Kmer.h:89-94)
// A5-A8: linear merge.  Returns false on Q2 (taxid 0 / unmapped species).
bool match_kmers(const Database &db, const std::vector<Kmer> &sortedKmers, std::vector<Match> &out,
                 int threads, std::string *err);
void sort_matches(std::vector<Match> &m, int threads);         // A9 (KmerMatcher.cpp:1149-1166)
void score_reads(const Database &db, const Options &opt, const std::vector<Match> &sortedMatches,
                 std::vector<QueryInfo> &queries, int threads);  // A10-A12
// A13 Reporter.cpp:35-80; lineage = --lineage 1 (TaxonomyWrapper::taxLineage2, TaxonomyWrapper.cpp:431-454)
void write_tsv_header(std::string &out, bool lineage = false);
void write_tsv_rows(const Database &db, const std::vector<Read> &m1, const std::vector<QueryInfo> &q, std::string &out, bool lineage = false);

// Reporter::writeReportFile / writeReport (Reporter.cpp:117-193): text of <jobid>_report.tsv from the per-read classifications
void write_report(const Database &db, const std::vector<QueryInfo> &q, std::string &out);

// delta codec (A6) — exposed for tests
uint64_t next_target_kmer(uint64_t prev, const uint16_t *diff, size_t &idx);   // KmerMatcher.h:282-297
void encode_delta(uint64_t delta, std::vector<uint16_t> &out);                // IndexCreator.cpp:874-892
uint8_t hamming_sum(uint64_t a, uint64_t b);                                   // KmerMatcher.h:348-360
uint16_t hammings_fwd(uint64_t a, uint64_t b);                                 // KmerMatcher.h:386-400
uint16_t hammings_rev(uint64_t a, uint64_t b);                                 // KmerMatcher.h:402-416

// whole path, file to TSV text (Classifier.cpp:44-164 minus batching, which does not change results)
bool classify_files(const std::string &q1, const std::string &q2, const std::string &dbDir,
                    const Options &opt, std::string &tsv, std::string *err,
                    size_t *nKmers = nullptr, size_t *nMatches = nullptr, std::string *report = nullptr);

}  // namespace orc
