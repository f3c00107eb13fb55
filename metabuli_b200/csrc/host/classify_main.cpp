// metabuli-b200 — C++ host for the classify hot path: keeps the `metabuli classify` command line and the
// on-disk DB / output formats of the reference (src/workflow/classify.cpp:39-200, Classifier.cpp:44-164,
// Reporter.cpp:35-80) and drives the CUDA path through the C-ABI (include/metabuli_b200.h).
//
//   metabuli-b200 classify [--seq-mode 1|2|3] [flags] <fastx> [<fastx2>] <dbdir> <outdir> <jobid>
//
// Host work kept here: argument checks, db.parameters, taxonomyDB / taxID_list parsing, FASTA/FASTQ
// reading (kseq semantics), batching, and writing <jobid>_classifications.tsv.  No CPU fallback: if the
// library cannot reach a CUDA device the run stops with an error.
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/metabuli_b200.h"
#include "fastx_tsv.hpp"

namespace {

[[noreturn]] void die(const std::string& msg) {
    fprintf(stderr, "Error: %s\n", msg.c_str());
    exit(1);
}
bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
bool ends_with(const std::string& s, const char* suf) { size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }
bool valid_query_file(const std::string& p) {          // LocalUtil.h:19-34
    for (const char* e : {".fna", ".fasta", ".fa", ".fq", ".fastq", ".fna.gz", ".fasta.gz", ".fa.gz", ".fq.gz", ".fastq.gz"})
        if (ends_with(p, e)) return true;
    return false;
}

template <class T>
std::vector<T> slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v((size_t)sz / sizeof(T));
    if (!v.empty() && fread(v.data(), sizeof(T), v.size(), f) != v.size()) die("short read on " + path);
    fclose(f);
    return v;
}

// The index files are mapped, not copied (the reference does the same: mmapData, KmerMatcher.cpp:137-139): a 40 GiB diffIdx is
// neither read into a second host copy nor zero-filled first; the library's upload reads the pages straight from the page cache.
template <class T>
class MappedFile {
public:
    explicit MappedFile(const std::string& path) {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) die("cannot open " + path);
        struct stat st;
        if (fstat(fd, &st) != 0) die("cannot stat " + path);
        bytes_ = (size_t)st.st_size;
        if (bytes_) {
            p_ = mmap(nullptr, bytes_, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p_ == MAP_FAILED) die("cannot map " + path);
            madvise(p_, bytes_, MADV_SEQUENTIAL);
            madvise(p_, bytes_, MADV_WILLNEED);
        }
        close(fd);
    }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    ~MappedFile() { if (p_ && p_ != MAP_FAILED) munmap(p_, bytes_); }
    const T* data() const { return static_cast<const T*>(p_); }
    size_t size() const { return bytes_ / sizeof(T); }
private:
    void* p_ = nullptr;
    size_t bytes_ = 0;
};

// ---- taxonomyDB (TaxonomyWrapper.cpp:363-421) -------------------------------------------------------------
struct TaxonomyHost {
    std::vector<char> blob;
    bool internalIds = false;
    size_t maxNodes = 0;
    int32_t maxTaxID = 0, Mk = 0, eukaryota = 0;
    std::vector<int32_t> nodeTaxId, nodeParent;
    std::vector<uint64_t> rankIdx, nameIdx;
    std::vector<uint8_t> prune;
    std::vector<int8_t> rank;
    const int32_t *D = nullptr, *i2o = nullptr, *E = nullptr, *L = nullptr, *H = nullptr, *M = nullptr;
    const char* strData = nullptr;
    const uint32_t* strOff = nullptr;

    const char* str(uint64_t i) const { return strData + strOff[i]; }
    static int rankIndex(const std::string& r) {     // NcbiTaxonomy.h:52-80
        static const std::map<std::string, int> m = {
            {"forma", 1}, {"varietas", 2}, {"subspecies", 3}, {"species", 4}, {"species subgroup", 5}, {"species group", 6},
            {"subgenus", 7}, {"genus", 8}, {"subtribe", 9}, {"tribe", 10}, {"subfamily", 11}, {"family", 12},
            {"superfamily", 13}, {"parvorder", 14}, {"infraorder", 15}, {"suborder", 16}, {"order", 17}, {"superorder", 18},
            {"infraclass", 19}, {"subclass", 20}, {"class", 21}, {"superclass", 22}, {"subphylum", 23}, {"phylum", 24},
            {"superphylum", 25}, {"subkingdom", 26}, {"kingdom", 27}, {"superkingdom", 28}, {"domain", 28}};
        auto it = m.find(r);
        return it == m.end() ? -1 : it->second;
    }
    void load(const std::string& path) {
        blob = slurp<char>(path);
        const char* p = blob.data();
        int32_t version; memcpy(&version, p, 4); p += 4;
        if (version != 2) die("Outdated taxonomy information, please recreate with createtaxdb.");
        uint64_t flag; memcpy(&flag, p, 8);
        internalIds = flag == 1;
        if (internalIds) p += 8;
        uint64_t mn; memcpy(&mn, p, 8); p += 8; maxNodes = (size_t)mn;
        memcpy(&maxTaxID, p, 4); p += 4;
        nodeTaxId.resize(maxNodes); nodeParent.resize(maxNodes); rankIdx.resize(maxNodes); nameIdx.resize(maxNodes);
        for (size_t i = 0; i < maxNodes; ++i) {
            const char* n = p + 32 * i;
            memcpy(&nodeTaxId[i], n + 4, 4); memcpy(&nodeParent[i], n + 8, 4); memcpy(&rankIdx[i], n + 16, 8); memcpy(&nameIdx[i], n + 24, 8);
        }
        p += 32 * maxNodes;
        D = (const int32_t*)p; p += 4 * ((size_t)maxTaxID + 1);
        i2o = (const int32_t*)p; if (internalIds) p += 4 * ((size_t)maxTaxID + 1);
        E = (const int32_t*)p; p += 8 * maxNodes;
        L = (const int32_t*)p; p += 8 * maxNodes;
        H = (const int32_t*)p; p += 4 * maxNodes;
        Mk = 1; for (size_t v = 2 * maxNodes; v > 1; v >>= 1) ++Mk;
        M = (const int32_t*)p; p += 4 * 2 * maxNodes * (size_t)Mk;
        uint64_t byteCap; uint32_t entryCap, entryCnt;
        memcpy(&byteCap, p, 8); p += 8; memcpy(&entryCap, p, 4); p += 4; memcpy(&entryCnt, p, 4); p += 4;
        strData = p; p += byteCap;
        strOff = (const uint32_t*)p; p += 4 * (size_t)entryCap;
        if ((size_t)(p - blob.data()) > blob.size()) die("taxonomyDB truncated");
        prune.resize(maxNodes); rank.resize(maxNodes);
        for (size_t i = 0; i < maxNodes; ++i) {
            std::string r = str(rankIdx[i]);
            rank[i] = (int8_t)rankIndex(r);
            prune[i] = (r.empty() || r == "accession") ? 1 : 0;
        }
        for (size_t i = 0; i < maxNodes; ++i)
            if (nameIdx[i] != 0 && strcmp(str(nameIdx[i]), "Eukaryota") == 0) { eukaryota = nodeTaxId[i]; break; }
    }
    std::string lineage(int32_t taxId) const {        // TaxonomyWrapper::taxLineage2 (TaxonomyWrapper.cpp:431-454)
        static const std::map<std::string, std::string> shortRanks = {     // ExtendedShortRanks (TaxonomyWrapper.h:9-26)
            {"subspecies", "ss"}, {"species", "s"}, {"subgenus", "sg"}, {"genus", "g"}, {"subfamily", "sf"}, {"family", "f"},
            {"suborder", "so"}, {"order", "o"}, {"subclass", "sc"}, {"class", "c"}, {"subphylum", "sp"}, {"phylum", "p"},
            {"subkingdom", "sk"}, {"kingdom", "k"}, {"superkingdom", "d"}, {"domain", "d"}, {"realm", "r"}};
        std::vector<int> chain;
        int node = D[taxId];
        do { chain.push_back(node); node = D[nodeParent[node]]; } while (nodeParent[node] != nodeTaxId[node]);
        std::string out;
        for (int i = (int)chain.size() - 1; i >= 0; --i) {
            auto it = shortRanks.find(str(rankIdx[chain[(size_t)i]]));
            out += it == shortRanks.end() ? "-" : it->second;
            out += '_';
            out += str(nameIdx[chain[(size_t)i]]);
            if (i > 0) out += ';';
        }
        return out;
    }
    bool exists(int32_t t) const { return t <= maxTaxID && D[t] != -1; }
    int32_t original(int32_t t) const { return internalIds ? i2o[t] : t; }
    int32_t atSpecies(int32_t taxId) const {          // TaxonomyWrapper.cpp:479-498 with rank "species"
        if (taxId == 0 || !exists(taxId) || taxId == 1) return 0;
        int node = D[taxId], cnt = 0;
        while (cnt < 30 && rank[node] < 4) { node = D[nodeParent[node]]; ++cnt; }
        return cnt == 30 ? taxId : nodeTaxId[node];
    }
    std::vector<int32_t> taxid2species(const std::string& listPath) const {   // KmerMatcher.cpp:96-119
        std::vector<int32_t> out((size_t)maxTaxID + 1, 0);
        std::ifstream in(listPath);
        if (!in) die("Cannot open the taxID list file.");
        std::string line;
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            int32_t taxId = (int32_t)strtoul(line.c_str(), nullptr, 10);
            int32_t sp = atSpecies(taxId);
            int node = D[taxId];
            if (taxId != nodeTaxId[node]) out[taxId] = sp;
            while (nodeTaxId[node] != sp) { out[nodeTaxId[node]] = sp; node = D[nodeParent[node]]; }
            out[sp] = sp;
        }
        return out;
    }
};

struct Params {
    int lineage = 0;
    int seqMode = 2, threads = 0, accessionLevel = 0, minConsCnt = 4, minConsCntEuk = 9, matchPerKmer = 4, device = 0;
    float minScore = 0.f, minSpScore = 0.f, tieRatio = 0.95f;
    int syncmer = 0, smerLen = 5, kmerFormat = 1;   // classify.cpp:11-13 defaults; db.parameters overrides them (loadDbParameters)
    int maskMode = 0;                    // --mask 1: tantan masking of the queries before extraction (classify.cpp:31-32 defaults)
    float maskProb = 0.9f;
    int maskHost = 0;                    // --mask-host 1: mask in the reader thread (mbl_mask_reads) instead of on the device (K0)
    size_t batchReads = 0;
    std::vector<int> devices;            // --gpus N (devices 0..N-1) or --devices a,b,c: one replica of the index per device
    int indexSharded = 0;                // --index-sharded 1: the index is range-partitioned over the devices instead of replicated
    std::vector<std::string> files;
};

int classify(int argc, char** argv) {
    Params par;
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) die("missing value for " + a); return argv[++i]; };
        if (a == "--seq-mode") par.seqMode = atoi(val());
        else if (a == "--threads") par.threads = atoi(val());
        else if (a == "--min-score") par.minScore = (float)atof(val());
        else if (a == "--min-sp-score") par.minSpScore = (float)atof(val());
        else if (a == "--tie-ratio") par.tieRatio = (float)atof(val());
        else if (a == "--min-cons-cnt") par.minConsCnt = atoi(val());
        else if (a == "--min-cons-cnt-euk") par.minConsCntEuk = atoi(val());
        else if (a == "--accession-level") par.accessionLevel = atoi(val());
        else if (a == "--match-per-kmer") par.matchPerKmer = atoi(val());
        else if (a == "--device") par.device = atoi(val());
        else if (a == "--index-sharded") par.indexSharded = atoi(val());
        else if (a == "--gpus") { const int n = atoi(val()); if (n < 1 || n > 64) die("--gpus takes 1..64"); par.devices.clear(); for (int d = 0; d < n; ++d) par.devices.push_back(d); }
        else if (a == "--devices") { par.devices.clear(); std::string v = val(); size_t p0 = 0; while (p0 <= v.size()) { size_t c = v.find(',', p0); if (c == std::string::npos) c = v.size(); if (c > p0) par.devices.push_back(atoi(v.substr(p0, c - p0).c_str())); p0 = c + 1; } if (par.devices.empty()) die("--devices takes a comma-separated list"); }
        else if (a == "--batch-reads") par.batchReads = (size_t)atoll(val());
        else if (a == "--lineage") par.lineage = atoi(val());
        // --mask 1 masks low-complexity regions before extraction (KmerExtractor.cpp:308-314, SeqIterator.cpp:154-175): on the
        // device after every upload (mbl_config.mask_mode), or with --mask-host 1 by the reader thread through mbl_mask_reads
        else if (a == "--mask") par.maskMode = atoi(val());
        else if (a == "--mask-prob") par.maskProb = (float)atof(val());
        else if (a == "--mask-host") par.maskHost = atoi(val());
        // --taxonomy-path: the reference reads it only when <dbdir>/taxonomyDB is missing (loadTaxonomy, common.cpp:50-86; checked
        // against the reference binary: with taxonomyDB present even a non-existent path changes nothing).  This host needs
        // taxonomyDB (checked below), so the flag is accepted and has no effect, as in the reference.
        else if (a == "--taxonomy-path") val();
        // defaults for databases whose db.parameters lacks the entry (older builds); the database's own values win
        else if (a == "--syncmer") par.syncmer = atoi(val());
        else if (a == "--smer-len") par.smerLen = atoi(val());
        else if (a == "--kmer-format") par.kmerFormat = atoi(val());
        else if (a == "--print-log") val();
        // --em (a switch, optionally followed by a boolean): EM re-assignment changes what classify writes and is not implemented
        else if (a == "--em") {
            bool on = true;
            if (i + 1 < argc) {
                std::string v = argv[i + 1];
                for (auto& ch : v) ch = (char)tolower((unsigned char)ch);
                if (v == "0" || v == "false") { on = false; ++i; } else if (v == "1" || v == "true") ++i;
            }
            if (on) die("--em (expectation-maximisation re-assignment) is not supported by the B200 path");
        }
        // flags that change the reference's output and are not implemented here must fail, never be dropped silently:
        else if (a == "--reduced-aa") { if (atoi(val()) != 0) die("--reduced-aa 1 is not supported by the B200 path"); }
        // flags without influence on the classifications: --max-ram only sizes the reference's query splits (here: HBM budget),
        // --hamming-margin is parsed but unused by the reference (KmerMatcher.cpp:1117-1146), -v is verbosity
        else if (a == "--max-ram" || a == "-v" || a == "--hamming-margin") val();
        // the validators only check the inputs and exit on malformed files; on valid inputs the output is unchanged
        else if (a == "--validate-input" || a == "--validate-db") { if (atoi(val()) != 0) fprintf(stderr, "warning: %s is ignored by the B200 path (inputs are not validated)\n", a.c_str()); }
        else if (a.rfind("--", 0) == 0) die("unknown flag " + a);
        else par.files.push_back(a);
    }
    size_t want = par.seqMode == 2 ? 5 : 4;
    if (par.files.size() != want) die("expected <fastx>" + std::string(par.seqMode == 2 ? " <fastx2>" : "") + " <dbdir> <outdir> <jobid>");
    const std::string q1 = par.files[0], q2 = par.seqMode == 2 ? par.files[1] : "";
    const std::string dbDir = par.files[want - 3], outDir = par.files[want - 2], jobId = par.files[want - 1];
    // classify.cpp:44-190 checks
    if (!valid_query_file(q1)) die(q1 + " is not a valid query file.");
    if (par.seqMode == 2 && !valid_query_file(q2)) die(q2 + " is not a valid query file.");
    if (!file_exists(q1)) die("Query file " + q1 + " is NOT found.");
    if (par.seqMode == 2 && !file_exists(q2)) die("Query file " + q2 + " is NOT found.");
    for (const char* f : {"/diffIdx", "/info", "/split", "/taxID_list"})
        if (!file_exists(dbDir + f)) die(dbDir + f + " is NOT found.");
    if (!file_exists(dbDir + "/taxonomyDB"))
        die(dbDir + "/taxonomyDB is NOT found (building the taxonomy from names.dmp / nodes.dmp — <dbdir>/taxonomy or --taxonomy-path — is not implemented on the B200 path).");
    if (!file_exists(outDir)) mkdir(outDir.c_str(), 0755);

    // loadDbParameters (common.cpp:88-133)
    mbl_config cfg{};
    cfg.kmer_format = par.kmerFormat; cfg.smer_len = par.smerLen; cfg.syncmer = par.syncmer ? 1 : 0; cfg.seq_mode = par.seqMode;
    cfg.min_score = par.minScore; cfg.min_sp_score = par.minSpScore; cfg.tie_ratio = par.tieRatio;
    cfg.min_cons_cnt = par.minConsCnt; cfg.min_cons_cnt_euk = par.minConsCntEuk;
    cfg.accession_level = par.accessionLevel; cfg.device = par.device; cfg.match_per_kmer = par.matchPerKmer;
    cfg.mask_mode = (par.maskMode && !par.maskHost) ? 1 : 0; cfg.mask_prob = par.maskProb;
    {
        std::ifstream pf(dbDir + "/db.parameters");
        std::string line;
        while (std::getline(pf, line)) {
            size_t tab = line.find('\t');
            if (tab == std::string::npos) continue;
            std::string k = line.substr(0, tab), v = line.substr(tab + 1);
            if (k == "Reduced_alphabet") cfg.reduced_aa = atoi(v.c_str());
            else if (k == "Skip_redundancy") { if (v == "1") cfg.skip_redundancy = 1; }
            else if (k == "Syncmer") { if (v == "1") cfg.syncmer = 1; }
            else if (k == "S-mer_len") cfg.smer_len = atoi(v.c_str());
            else if (k == "Kmer_format") cfg.kmer_format = atoi(v.c_str());
            else if (k == "DB_name") printf("Database name : %s\n", v.c_str());
            else if (k == "Creation_date") printf("Creation date : %s\n", v.c_str());
            else if (k == "Accession_level") {
                if (v == "0" && cfg.accession_level == 1) { cfg.accession_level = 0; printf("Warning: Current DB doesn't support accession-level classification.\n"); }
                if (v == "1" && cfg.accession_level == 0) cfg.accession_level = 2;
            }
        }
    }
    if (cfg.reduced_aa) die("reduced-alphabet databases are not supported by the B200 path yet");

    TaxonomyHost tax;
    tax.load(dbDir + "/taxonomyDB");
    std::vector<int32_t> t2s = tax.taxid2species(dbDir + "/taxID_list");
    const MappedFile<uint16_t> diff(dbDir + "/diffIdx");
    const MappedFile<int32_t> info(dbDir + "/info");
    std::vector<uint64_t> split = slurp<uint64_t>(dbDir + "/split");

    // one context per device, each holding a replica of the index (the host arrays are shared)
    if (par.devices.empty()) par.devices.push_back(par.device);
    const size_t G = par.devices.size();
    std::vector<mbl_ctx*> ctxs(G, nullptr);
    mbl_db db{diff.data(), diff.size(), info.data(), info.size(), split.data(), split.size() / 3};
    mbl_taxonomy tx{tax.maxNodes, tax.maxTaxID, tax.eukaryota, tax.D, tax.E, tax.L, tax.H, tax.M, tax.Mk,
                    tax.nodeTaxId.data(), tax.nodeParent.data(), tax.prune.data(), tax.rank.data(), t2s.data()};
    const bool sharded = par.indexSharded != 0 && G > 1;
    std::vector<mbl_shard> shards(G);
    if (sharded && mbl_plan_shards(&db, (uint32_t)G, shards.data()) != MBL_OK) die("mbl_plan_shards failed");
    {
        std::vector<std::thread> loaders;
        std::vector<std::string> lerr(G);
        for (size_t g = 0; g < G; ++g) loaders.emplace_back([&, g] {
            mbl_config cg = cfg;
            cg.device = par.devices[g];
            int rc = mbl_create(&cg, &ctxs[g]);
            if (rc != MBL_OK) { lerr[g] = rc == MBL_E_NO_DEVICE ? "no usable CUDA device " + std::to_string(cg.device) + " (this build has no CPU fallback)" : "mbl_create failed"; return; }
            rc = sharded ? mbl_load_db_shard(ctxs[g], &db, &tx, &shards[g]) : mbl_load_db(ctxs[g], &db, &tx);
            if (rc != MBL_OK) lerr[g] = std::string("mbl_load_db: ") + mbl_last_error(ctxs[g]);
        });
        for (auto& t : loaders) t.join();
        for (size_t g = 0; g < G; ++g) if (!lerr[g].empty()) die(lerr[g]);
    }
    if (sharded) {
        // every shard's presence filter covers its own value range only: OR the parts together (each context reads its peers'
        // filters through peer access; OR-ing a part that already holds other parts changes nothing)
        std::vector<void*> fptr(G, nullptr);
        uint64_t fbytes = 0;
        for (size_t g = 0; g < G; ++g) if (mbl_shard_filter(ctxs[g], &fptr[g], &fbytes) != MBL_OK) die(mbl_last_error(ctxs[g]));
        for (size_t g = 0; g < G; ++g)
            for (size_t h = 0; h < G; ++h)
                if (h != g && fptr[h] && mbl_shard_filter_or(ctxs[g], fptr[h], fbytes, 0) != MBL_OK) die(std::string("mbl_shard_filter_or: ") + mbl_last_error(ctxs[g]));
        for (size_t g = 0; g < G; ++g) if (mbl_shard_filter_or(ctxs[g], nullptr, fbytes, 1) != MBL_OK) die(mbl_last_error(ctxs[g]));
    }

    const unsigned T = par.threads > 0 ? (unsigned)par.threads : std::max(1u, std::thread::hardware_concurrency());
    const std::string tsvPath = outDir + "/" + jobId + "_classifications.tsv";
    FILE* out = fopen(tsvPath.c_str(), "wb");
    if (!out) die("cannot write " + tsvPath);
    fputs(par.lineage ? "#is_classified\tname\ttaxID\tquery_length\tscore\trank\tlineage\ttaxID:match_count\n"
                      : "#is_classified\tname\ttaxID\tquery_length\tscore\trank\ttaxID:match_count\n", out);      // Reporter.cpp:37-41

    struct TaxView {
        const TaxonomyHost& t;
        int32_t original(int32_t x) const { return t.original(x); }
        const char* rank_name(int32_t x) const { return t.str(t.rankIdx[t.D[x]]); }
        std::string lineage(int32_t x) const { return t.lineage(x); }
    } tv{tax};
    // taxonomy in the shape mblhost::write_report takes (Reporter::writeReportFile, Reporter.cpp:117-193)
    std::vector<const char*> nodeRankStr(tax.maxNodes), nodeNameStr(tax.maxNodes);
    for (size_t i = 0; i < tax.maxNodes; ++i) { nodeRankStr[i] = tax.str(tax.rankIdx[i]); nodeNameStr[i] = tax.str(tax.nameIdx[i]); }
    mblhost::ArrayTax rt{tax.maxNodes, tax.maxTaxID, tax.nodeTaxId.data(), tax.nodeParent.data(), tax.D,
                         tax.internalIds ? tax.i2o : nullptr, nodeRankStr.data(), nodeNameStr.data()};
    std::vector<uint64_t> taxCounts((size_t)tax.maxTaxID + 1, 0);           // Classifier.cpp:196-203: ++taxCounts[classification]

    // ---- the QuerySplit loop (Classifier.cpp:81-140) as a pipeline of batches --------------------------------------------
    //   reader thread : FASTA/FASTQ(.gz) in bounded chunks -> batches of --batch-reads reads (mblhost::FastxStream)
    //   device threads: one per context; batch i+1 of a device uploads while its batch i is classified
    //                   (mbl_prefetch_batch / mbl_classify_prefetched)
    //   writer thread : rows formatted by the host threads and written in batch order, taxon counts for the report
    // Batch objects rotate through a free list, so their buffers (pinned once per growth) are reused and what is held in memory
    // is bounded by the batches in flight, not by the size of the input (the reference bounds its splits by --max-ram).
    struct Batch {
        size_t index = 0;
        mblhost::ReadSet r1, r2;
        mbl_batch b{};
        std::vector<mbl_read_result> res;
        std::vector<int32_t> pairs;
        size_t used = 0;
        const void* pinned[2] = {nullptr, nullptr};
        size_t pinned_cap[2] = {0, 0};
    };
    // index-sharded: a batch is one exchange round over all devices (receive buffers and match buffers are sized by what one
    // round moves: 1.25 M reads or pairs per device by default)
    // Default batch (replicas): every batch streams the whole index through the merge kernel once (~4.6 ms per GiB on a B200,
    // DESIGN §7) while the host needs ~0.1 us per read to parse and format it (§6), so a batch of ~110 k reads per GiB of index
    // keeps the device's fixed cost under half of the host's time for the same batch; smaller batches than that starve the GPU,
    // larger ones only delay the first output and stop the reader and the writer from overlapping.  Clamped to [1 M, 8 M].
    const double index_gib = (2.0 * (double)diff.size() + 4.0 * (double)info.size()) / (double)(1ull << 30);
    const size_t auto_step = (size_t)std::min(8.0e6, std::max(1.0e6, 110.0e3 * index_gib));
    const size_t step = par.batchReads ? par.batchReads : (sharded ? (size_t)1250000 * G : auto_step);
    const size_t n_batches_in_flight = sharded ? 4 : 2 * G + 2;
    std::vector<std::unique_ptr<Batch>> pool;
    for (size_t i = 0; i < n_batches_in_flight; ++i) pool.emplace_back(new Batch());
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Batch*> free_list, ready;               // reader -> devices
    std::map<size_t, Batch*> finished;                 // devices -> writer, by batch index
    for (auto& p : pool) free_list.push_back(p.get());
    bool reader_done = false, failed = false;
    std::string fail_msg;
    size_t n_batches = 0, total_reads = 0, total_bases = 0;
    auto fail_with = [&](const std::string& m) { std::lock_guard<std::mutex> lk(mu); if (!failed) { failed = true; fail_msg = m; } cv.notify_all(); };
    auto pin = [&](Batch* bt, int k, const mblhost::ByteBuf& v) {          // (re)register a batch buffer when it moved or grew
        if (v.data() == bt->pinned[k] && v.capacity() == bt->pinned_cap[k]) return;
        if (bt->pinned[k]) mbl_host_unregister(const_cast<void*>(bt->pinned[k]));
        bt->pinned[k] = nullptr; bt->pinned_cap[k] = 0;
        if (v.capacity() && mbl_host_register(const_cast<char*>(v.data()), v.capacity()) == MBL_OK) { bt->pinned[k] = v.data(); bt->pinned_cap[k] = v.capacity(); }
    };
    for (auto& p : pool) {
        Batch* bt = p.get();
        auto unpin = [bt](int k) { if (bt->pinned[k]) mbl_host_unregister(const_cast<void*>(bt->pinned[k])); bt->pinned[k] = nullptr; bt->pinned_cap[k] = 0; };
        bt->r1.before_realloc = [unpin] { unpin(0); };
        bt->r2.before_realloc = [unpin] { unpin(1); };
    }
    auto tl0 = std::chrono::steady_clock::now();

    std::thread reader([&] {
        try {
        mblhost::FastxStream s1, s2;
        std::string err;
        if (!s1.open(q1, &err) || (par.seqMode == 2 && !s2.open(q2, &err))) { fail_with(err); return; }
        for (size_t index = 0;; ++index) {
            Batch* bt = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || !free_list.empty(); });
                if (failed) return;
                bt = free_list.front(); free_list.pop_front();
            }
            if (par.seqMode == 2) {
                // the two mate files are read, inflated and parsed side by side
                std::string err2;
                bool ok2 = true;
                std::thread mate2([&] { ok2 = s2.next(bt->r2, step, std::max(1u, T / 2), &err2); });
                const bool ok1 = s1.next(bt->r1, step, std::max(1u, T - T / 2), &err);
                mate2.join();
                if (!ok1) { fail_with(err); return; }
                if (!ok2) { fail_with(err2); return; }
                if (bt->r1.size() != bt->r2.size()) { fail_with("The number of reads in the two files are not equal."); return; }
            } else if (!s1.next(bt->r1, step, T, &err)) { fail_with(err); return; }
            if (par.maskMode && par.maskHost) {
                // the names and lengths the Reporter prints stay the file's; only the letters the extractor sees change
                if (mbl_mask_reads(bt->r1.bases.data(), bt->r1.offsets.data(), (uint32_t)bt->r1.size(), par.maskProb, (int)T) != MBL_OK ||
                    (par.seqMode == 2 && mbl_mask_reads(bt->r2.bases.data(), bt->r2.offsets.data(), (uint32_t)bt->r2.size(), par.maskProb, (int)T) != MBL_OK)) {
                    fail_with("masking the queries failed"); return;
                }
            }
            std::unique_lock<std::mutex> lk(mu);
            if (bt->r1.size() == 0) { free_list.push_back(bt); n_batches = index; reader_done = true; cv.notify_all(); return; }
            bt->index = index;
            bt->b = mbl_batch{};
            bt->b.bases = bt->r1.bases.data(); bt->b.offsets = bt->r1.offsets.data(); bt->b.n_reads = (uint32_t)bt->r1.size();
            if (par.seqMode == 2) { bt->b.bases2 = bt->r2.bases.data(); bt->b.offsets2 = bt->r2.offsets.data(); }
            total_reads += bt->r1.size(); total_bases += bt->r1.bases.size() + bt->r2.bases.size();
            lk.unlock();
            pin(bt, 0, bt->r1.bases);
            if (par.seqMode == 2) pin(bt, 1, bt->r2.bases);
            lk.lock();
            ready.push_back(bt);
            cv.notify_all();
        }
        } catch (const std::exception& e) { fail_with(std::string("host thread failed: ") + e.what()); }
          catch (...) { fail_with("host thread failed"); }
    });

    uint64_t kmers = 0, matches = 0;
    auto t0 = std::chrono::steady_clock::now();
    auto next_ready = [&]() -> Batch* {                                     // blocks; nullptr at the end of the input
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed || !ready.empty() || reader_done; });
        if (failed || ready.empty()) return nullptr;
        Batch* bt = ready.front(); ready.pop_front();
        return bt;
    };
    std::vector<std::thread> workers;
    // ---- index-sharded rounds: the batch's reads are cut evenly over the devices; every device thread runs the three
    // phases of its rank around two exchanges, in lock step through a barrier; the exchanges are the library's push kernels
    // storing straight into the owners' receive buffers (peer access inside one process), the counts travel through host memory
    struct Barrier {
        std::mutex m; std::condition_variable c; size_t n, waiting = 0, gen = 0; bool aborted = false;
        explicit Barrier(size_t k) : n(k) {}
        void wait() { std::unique_lock<std::mutex> lk(m); if (aborted) return; const size_t g = gen; if (++waiting == n) { waiting = 0; ++gen; c.notify_all(); } else c.wait(lk, [&] { return gen != g || aborted; }); }
        void abort() { std::lock_guard<std::mutex> lk(m); aborted = true; c.notify_all(); }      // a rank left the loop (exception): nobody may wait for it
    } bar(G);
    struct Round {
        Batch* bt = nullptr;
        std::vector<uint64_t> first_read;                 // [G + 1] read ranges of the ranks
        std::vector<std::vector<uint64_t>> K, M;          // counts [src][dst]: metamers to shards, matches to read owners
        std::vector<std::vector<int32_t>> pairs;          // per rank
        std::vector<size_t> used;
        std::vector<void*> rk, rm;                        // receive buffers of the ranks
        uint64_t cap_k = 0, cap_m = 0;
        bool stop = false;
    } rd;
    rd.K.assign(G, std::vector<uint64_t>(G, 0)); rd.M = rd.K; rd.pairs.resize(G); rd.used.assign(G, 0); rd.rk.assign(G, nullptr); rd.rm.assign(G, nullptr);
    std::vector<uint64_t> first_values(G);
    for (size_t g = 0; g < G; ++g) first_values[g] = shards[g].first_value;
    auto ensure_recv = [&](uint64_t need_k, uint64_t need_m) {   // thread 0, between barriers: grow every rank's receive buffers together
        if (need_k <= rd.cap_k && need_m <= rd.cap_m) return;
        rd.cap_k = std::max(rd.cap_k, need_k + need_k / 4 + 4096); rd.cap_m = std::max(rd.cap_m, need_m + need_m / 4 + 4096);
        for (size_t g = 0; g < G; ++g) mbl_shard_detach_peers(ctxs[g]);
        for (size_t g = 0; g < G; ++g)
            if (mbl_shard_recv_buffers(ctxs[g], rd.cap_k, rd.cap_m, &rd.rk[g], &rd.rm[g], nullptr, nullptr) != MBL_OK) { fail_with(std::string("mbl_shard_recv_buffers: ") + mbl_last_error(ctxs[g])); return; }
        for (size_t g = 0; g < G; ++g)
            for (size_t h = 0; h < G; ++h)
                if (mbl_shard_attach_peer(ctxs[g], (uint32_t)h, nullptr, nullptr, rd.rk[h], rd.rm[h]) != MBL_OK) { fail_with(std::string("mbl_shard_attach_peer: ") + mbl_last_error(ctxs[g])); return; }
    };
    if (sharded) for (size_t g = 0; g < G; ++g) workers.emplace_back([&, g] {
        try {
        mbl_ctx* ctx = ctxs[g];
        auto check = [&](int rc, const char* what) { if (rc != MBL_OK) fail_with(std::string(what) + ": " + mbl_last_error(ctx)); };
        while (true) {
            if (g == 0) {
                rd.bt = next_ready();
                rd.stop = rd.bt == nullptr;
                if (!rd.stop) {
                    const size_t n = rd.bt->r1.size();
                    rd.first_read.assign(G + 1, 0);
                    for (size_t k = 0; k <= G; ++k) rd.first_read[k] = n * k / G;
                    rd.bt->res.assign(n, mbl_read_result{});
                }
            }
            bar.wait();
            if (rd.stop || failed) return;
            Batch* bt = rd.bt;
            const uint64_t r0 = rd.first_read[g], r1 = rd.first_read[g + 1];
            // this rank's reads as a batch of their own (offsets rebased)
            std::vector<uint64_t> o1(r1 - r0 + 1), o2;
            for (uint64_t k = 0; k <= r1 - r0; ++k) o1[k] = bt->r1.offsets[r0 + k] - bt->r1.offsets[r0];
            mbl_batch b{};
            b.bases = bt->r1.bases.data() + bt->r1.offsets[r0]; b.offsets = o1.data(); b.n_reads = (uint32_t)(r1 - r0);
            if (par.seqMode == 2) {
                o2.resize(r1 - r0 + 1);
                for (uint64_t k = 0; k <= r1 - r0; ++k) o2[k] = bt->r2.offsets[r0 + k] - bt->r2.offsets[r0];
                b.bases2 = bt->r2.bases.data() + bt->r2.offsets[r0]; b.offsets2 = o2.data();
            }
            check(mbl_shard_extract(ctx, &b, r0, (uint32_t)G, first_values.data(), rd.K[g].data()), "mbl_shard_extract");
            bar.wait();
            if (failed) return;
            std::vector<uint64_t> tot(G, 0), off(G, 0);
            for (size_t h = 0; h < G; ++h) for (size_t src = 0; src < G; ++src) { tot[h] += rd.K[src][h]; if (src < g) off[h] += rd.K[src][h]; }
            if (g == 0) ensure_recv(*std::max_element(tot.begin(), tot.end()), 0);
            bar.wait();
            if (failed) return;
            check(mbl_shard_push_kmers(ctx, off.data(), tot.data()), "mbl_shard_push_kmers");
            bar.wait();
            if (failed) return;
            const uint64_t nk = tot[g];
            check(mbl_shard_match(ctx, (const uint64_t*)rd.rk[g], (const uint64_t*)rd.rk[g] + nk, nk, (uint32_t)G, rd.first_read.data(), rd.M[g].data()), "mbl_shard_match");
            bar.wait();
            if (failed) return;
            std::fill(tot.begin(), tot.end(), 0); std::fill(off.begin(), off.end(), 0);
            for (size_t h = 0; h < G; ++h) for (size_t src = 0; src < G; ++src) { tot[h] += rd.M[src][h]; if (src < g) off[h] += rd.M[src][h]; }
            if (g == 0) ensure_recv(0, *std::max_element(tot.begin(), tot.end()));
            bar.wait();
            if (failed) return;
            check(mbl_shard_push_matches(ctx, off.data()), "mbl_shard_push_matches");
            bar.wait();
            if (failed) return;
            check(mbl_shard_score(ctx, (const mbl_match_rec*)rd.rm[g], tot[g]), "mbl_shard_score");
            std::vector<int32_t>& pr = rd.pairs[g];
            if (pr.size() < 10 * (r1 - r0) + 32) pr.assign(10 * (r1 - r0) + 32, 0);
            int rc = mbl_download_results(ctx, bt->res.data() + r0, pr.data(), pr.size() / 2, &rd.used[g]);
            if (rc == MBL_E_CAPACITY) { pr.assign(2 * (rd.used[g] + 16), 0); rc = mbl_download_results(ctx, bt->res.data() + r0, pr.data(), pr.size() / 2, &rd.used[g]); }
            check(rc, "mbl_download_results");
            mbl_stats st;
            mbl_get_stats(ctx, &st);
            { std::lock_guard<std::mutex> lk(mu); matches += st.n_matches; kmers += st.n_query_kmers; }
            bar.wait();
            if (failed) return;
            if (g == 0) {                                   // the ranks' pair lists behind one another, offsets re-based
                size_t total = 0;
                for (size_t h = 0; h < G; ++h) total += rd.used[h];
                bt->pairs.assign(2 * total + 2, 0);
                size_t base = 0;
                for (size_t h = 0; h < G; ++h) {
                    if (rd.used[h]) memcpy(bt->pairs.data() + 2 * base, rd.pairs[h].data(), 8 * rd.used[h]);
                    for (uint64_t r = rd.first_read[h]; r < rd.first_read[h + 1]; ++r) bt->res[r].taxcnt_begin += (uint32_t)base;
                    base += rd.used[h];
                }
                bt->used = total;
                { std::lock_guard<std::mutex> lk(mu); finished[bt->index] = bt; }
                cv.notify_all();
            }
        }
        } catch (const std::exception& e) { fail_with(std::string("host thread failed: ") + e.what()); bar.abort(); }
          catch (...) { fail_with("host thread failed"); bar.abort(); }
    });
    else for (size_t g = 0; g < G; ++g) workers.emplace_back([&, g] {
        try {
        mbl_ctx* ctx = ctxs[g];
        Batch* cur = next_ready();
        if (cur && mbl_prefetch_batch(ctx, &cur->b) != MBL_OK) { fail_with(std::string("mbl_prefetch_batch: ") + mbl_last_error(ctx)); return; }
        while (cur) {
            Batch* nxt = next_ready();
            const size_t n = cur->r1.size();
            cur->res.assign(n, mbl_read_result{});
            if (cur->pairs.size() < 10 * n + 32) cur->pairs.assign(10 * n + 32, 0);
            int rc = mbl_classify_prefetched(ctx, nxt ? &nxt->b : nullptr, cur->res.data(), cur->pairs.data(), cur->pairs.size() / 2, &cur->used);
            if (rc == MBL_E_CAPACITY) {                       // the batch is classified and resident: only the download is repeated
                cur->pairs.assign(2 * (cur->used + 16), 0);
                rc = mbl_download_results(ctx, cur->res.data(), cur->pairs.data(), cur->pairs.size() / 2, &cur->used);
            }
            if (rc != MBL_OK) { fail_with(std::string("mbl_classify_prefetched: ") + mbl_last_error(ctx)); return; }
            mbl_stats st;
            mbl_get_stats(ctx, &st);
            {
                std::lock_guard<std::mutex> lk(mu);
                kmers += st.n_query_kmers; matches += st.n_matches;
                finished[cur->index] = cur;
            }
            cv.notify_all();
            cur = nxt;
        }
        } catch (const std::exception& e) { fail_with(std::string("host thread failed: ") + e.what()); }
          catch (...) { fail_with("host thread failed"); }
    });

    std::thread writer([&] {
        try {
        size_t processed = 0;
        for (size_t want = 0;; ++want) {
            Batch* bt = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || finished.count(want) || (reader_done && want >= n_batches); });
                if (failed || !finished.count(want)) return;
                bt = finished[want]; finished.erase(want);
            }
            std::vector<std::string> rows;
            mblhost::format_rows(tv, bt->r1.names, 0, bt->r1.size(), bt->res.data(), bt->pairs.data(), T, rows, par.lineage != 0);
            for (const std::string& x : rows) fwrite(x.data(), 1, x.size(), out);
            for (size_t i = 0; i < bt->r1.size(); ++i) ++taxCounts[(size_t)bt->res[i].classification];
            processed += bt->r1.size();
            printf("Processed read count   : %zu\n", processed);
            std::lock_guard<std::mutex> lk(mu);
            free_list.push_back(bt);
            cv.notify_all();
        }
        } catch (const std::exception& e) { fail_with(std::string("host thread failed: ") + e.what()); }
          catch (...) { fail_with("host thread failed"); }
    });

    reader.join();
    for (auto& w : workers) w.join();
    { std::lock_guard<std::mutex> lk(mu); reader_done = true; }
    cv.notify_all();
    writer.join();
    if (failed) die(fail_msg);
    fclose(out);
    printf("--------------------\nTotal read count : %zu\nTotal read length: %zunt\n--------------------\n", total_reads, total_bases);
    {                                                           // <jobid>_report.tsv (Reporter.cpp:19, :117-137)
        std::string report;
        mblhost::write_report(rt, taxCounts, total_reads, report);
        FILE* rf = fopen((outDir + "/" + jobId + "_report.tsv").c_str(), "wb");
        if (!rf) die("cannot write " + outDir + "/" + jobId + "_report.tsv");
        fwrite(report.data(), 1, report.size(), rf);
        fclose(rf);
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    (void)tl0;
    printf("Query k-mer number     : %llu\nTotal k-mer match count: %llu\nTaxonomic classification completed on %zu GPU(s)%s. (%.3f s)\n",
           (unsigned long long)kmers, (unsigned long long)matches, G, sharded ? ", index sharded" : "", sec);
    for (auto& p : pool) for (int k = 0; k < 2; ++k) if (p->pinned[k]) mbl_host_unregister(const_cast<void*>(p->pinned[k]));
    for (mbl_ctx* c : ctxs) mbl_destroy(c);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2 || strcmp(argv[1], "classify") != 0) {
        fprintf(stderr, "usage: %s classify [--seq-mode 1|2|3] [--min-score F] [--min-sp-score F] [--tie-ratio F] [--min-cons-cnt N]\n"
                        "          [--min-cons-cnt-euk N] [--accession-level N] [--lineage 0|1] [--match-per-kmer N] [--device N]\n"
                        "          [--gpus N | --devices a,b,...] [--index-sharded 0|1] [--batch-reads N] [--threads N]\n"
                        "          <fastx> [<fastx2>] <dbdir> <outdir> <jobid>\n", argv[0]);
        return 2;
    }
    return classify(argc - 2, argv + 2);
}
