// metabuli-b200 — C++ host for the classify hot path: keeps the `metabuli classify` command line and the
// on-disk DB / output formats of the reference (src/workflow/classify.cpp:39-200, Classifier.cpp:44-164,
// Reporter.cpp:35-80) and drives the CUDA path through the C-ABI (include/metabuli_b200.h).
//
//   metabuli-b200 classify [--seq-mode 1|2|3] [flags] <fastx> [<fastx2>] <dbdir> <outdir> <jobid>
//
// Host work kept here: argument checks, db.parameters, taxonomyDB / taxID_list parsing, FASTA/FASTQ
// reading (kseq semantics), batching, and writing <jobid>_classifications.tsv.  No CPU fallback: if the
// library cannot reach a CUDA device the run stops with an error.
#include <sys/stat.h>
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../../../include/metabuli_b200.h"

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

// ---- FASTA/FASTQ with kseq semantics -------------------------------------------------------------------------
struct ReadFile {
    std::vector<std::string> names;
    std::vector<char> bases;
    std::vector<uint64_t> offsets{0};
    void load(const std::string& path) {
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) die("cannot open " + path);
        gzbuffer(g, 1 << 20);
        std::string data;
        std::vector<char> buf(1 << 22);
        int got;
        while ((got = gzread(g, buf.data(), (unsigned)buf.size())) > 0) data.append(buf.data(), (size_t)got);
        gzclose(g);
        size_t i = 0, n = data.size();
        auto line = [&](size_t& b, size_t& e) { b = i; while (i < n && data[i] != '\n') ++i; e = i; if (i < n) ++i; if (e > b && data[e - 1] == '\r') --e; };
        size_t b, e, entry = 0;
        while (i < n) {
            while (i < n && data[i] != '>' && data[i] != '@') line(b, e);
            if (i >= n) break;
            char tag = data[i];
            line(b, e);
            size_t p = b + 1;
            while (p < e && !isspace((unsigned char)data[p])) ++p;
            names.emplace_back(data, b + 1, p - (b + 1));
            size_t start = bases.size();
            while (i < n && data[i] != '>' && data[i] != '@' && data[i] != '+') {
                line(b, e);
                for (size_t x = b; x < e; ++x) if (isgraph((unsigned char)data[x])) bases.push_back(data[x]);
            }
            if (tag == '@' && i < n && data[i] == '+') {
                line(b, e);
                size_t ql = 0, sl = bases.size() - start;
                while (i < n && ql < sl) { line(b, e); ql += e - b; }
            }
            ++entry;
            if (bases.size() == start || names.back().empty()) {   // QueryIndexer.cpp:50-53
                printf("%zuth entry has no sequence or name.\n", entry);
                exit(1);
            }
            offsets.push_back(bases.size());
        }
    }
};

struct Params {
    int seqMode = 2, threads = 1, accessionLevel = 0, minConsCnt = 4, minConsCntEuk = 9, matchPerKmer = 4, device = 0;
    float minScore = 0.f, minSpScore = 0.f, tieRatio = 0.95f;
    size_t batchReads = 0;
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
        else if (a == "--batch-reads") par.batchReads = (size_t)atoll(val());
        else if (a == "--max-ram" || a == "--mask" || a == "--mask-prob" || a == "-v" || a == "--hamming-margin" ||
                 a == "--validate-input" || a == "--validate-db" || a == "--lineage" || a == "--taxonomy-path") val();
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
    for (const char* f : {"/diffIdx", "/info", "/split", "/taxonomyDB", "/taxID_list"})
        if (!file_exists(dbDir + f)) die(dbDir + f + " is NOT found.");
    if (!file_exists(outDir)) mkdir(outDir.c_str(), 0755);

    // loadDbParameters (common.cpp:88-133)
    mbl_config cfg{};
    cfg.kmer_format = 1; cfg.smer_len = 5; cfg.seq_mode = par.seqMode;
    cfg.min_score = par.minScore; cfg.min_sp_score = par.minSpScore; cfg.tie_ratio = par.tieRatio;
    cfg.min_cons_cnt = par.minConsCnt; cfg.min_cons_cnt_euk = par.minConsCntEuk;
    cfg.accession_level = par.accessionLevel; cfg.device = par.device; cfg.match_per_kmer = par.matchPerKmer;
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
    if (cfg.reduced_aa || cfg.syncmer) die("reduced-alphabet and syncmer databases are not supported by the B200 path yet");

    TaxonomyHost tax;
    tax.load(dbDir + "/taxonomyDB");
    std::vector<int32_t> t2s = tax.taxid2species(dbDir + "/taxID_list");
    std::vector<uint16_t> diff = slurp<uint16_t>(dbDir + "/diffIdx");
    std::vector<int32_t> info = slurp<int32_t>(dbDir + "/info");
    std::vector<uint64_t> split = slurp<uint64_t>(dbDir + "/split");

    mbl_ctx* ctx = nullptr;
    int rc = mbl_create(&cfg, &ctx);
    if (rc != MBL_OK) die(rc == MBL_E_NO_DEVICE ? "no usable CUDA device (this build has no CPU fallback)" : "mbl_create failed");
    mbl_db db{diff.data(), diff.size(), info.data(), info.size(), split.data(), split.size() / 3};
    mbl_taxonomy tx{tax.maxNodes, tax.maxTaxID, tax.eukaryota, tax.D, tax.E, tax.L, tax.H, tax.M, tax.Mk,
                    tax.nodeTaxId.data(), tax.nodeParent.data(), tax.prune.data(), tax.rank.data(), t2s.data()};
    rc = mbl_load_db(ctx, &db, &tx);
    if (rc != MBL_OK) die(std::string("mbl_load_db: ") + mbl_last_error(ctx));

    ReadFile r1, r2;
    r1.load(q1);
    if (par.seqMode == 2) {
        r2.load(q2);
        if (r1.names.size() != r2.names.size()) die("The number of reads in the two files are not equal.");
    }
    const size_t total = r1.names.size();
    printf("--------------------\nTotal read count : %zu\nTotal read length: %zunt\n--------------------\n", total,
           r1.bases.size() + r2.bases.size());

    const std::string tsvPath = outDir + "/" + jobId + "_classifications.tsv";
    FILE* out = fopen(tsvPath.c_str(), "wb");
    if (!out) die("cannot write " + tsvPath);
    fputs("#is_classified\tname\ttaxID\tquery_length\tscore\trank\ttaxID:match_count\n", out);      // Reporter.cpp:37-41
    const size_t step = par.batchReads ? par.batchReads : (total ? total : 1);
    std::vector<mbl_read_result> res;
    std::vector<int32_t> pairs;
    uint64_t kmers = 0, matches = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (size_t r0 = 0; r0 < total; r0 += step) {
        const size_t n = std::min(step, total - r0);
        mbl_batch b{};
        b.bases = r1.bases.data(); b.offsets = r1.offsets.data() + r0; b.n_reads = (uint32_t)n;
        if (par.seqMode == 2) { b.bases2 = r2.bases.data(); b.offsets2 = r2.offsets.data() + r0; }
        res.assign(n, mbl_read_result{});
        size_t cap = pairs.size() / 2, used = 0;
        rc = mbl_classify_batch(ctx, &b, res.data(), pairs.data(), cap, &used);
        if (rc == MBL_E_CAPACITY) {
            pairs.assign(2 * (used + 16), 0);
            rc = mbl_classify_batch(ctx, &b, res.data(), pairs.data(), used + 16, &used);
        }
        if (rc != MBL_OK) die(std::string("mbl_classify_batch: ") + mbl_last_error(ctx));
        mbl_stats st;
        mbl_get_stats(ctx, &st);
        kmers += st.n_query_kmers; matches += st.n_matches;
        // Reporter::writeReadClassification (Reporter.cpp:43-79)
        for (size_t i = 0; i < n; ++i) {
            const mbl_read_result& q = res[i];
            if (q.is_classified) {
                fprintf(out, "1\t%s\t%d\t%d\t%g\t%s\t", r1.names[r0 + i].c_str(), tax.original(q.classification), q.query_length,
                        (double)q.score, tax.str(tax.rankIdx[tax.D[q.classification]]));
                for (uint32_t k = q.taxcnt_begin; k < q.taxcnt_begin + q.taxcnt_len; ++k)
                    fprintf(out, "%d:%d ", tax.original(pairs[2 * k]), pairs[2 * k + 1]);
                fputc('\n', out);
            } else {
                fprintf(out, "0\t%s\t%d\t%d\t%g\t-\t-\t\n", r1.names[r0 + i].c_str(), tax.original(q.classification), q.query_length, (double)q.score);
            }
        }
        printf("Processed read count   : %zu (%g)\n", r0 + n, (double)(r0 + n) / (double)total);
    }
    fclose(out);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Query k-mer number     : %llu\nTotal k-mer match count: %llu\nTaxonomic classification completed. (%.3f s)\n",
           (unsigned long long)kmers, (unsigned long long)matches, sec);
    mbl_destroy(ctx);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2 || strcmp(argv[1], "classify") != 0) {
        fprintf(stderr, "usage: %s classify [--seq-mode 1|2|3] [--min-score F] [--min-sp-score F] [--tie-ratio F] [--min-cons-cnt N]\n"
                        "          [--min-cons-cnt-euk N] [--accession-level N] [--match-per-kmer N] [--device N] [--batch-reads N]\n"
                        "          <fastx> [<fastx2>] <dbdir> <outdir> <jobid>\n", argv[0]);
        return 2;
    }
    return classify(argc - 2, argv + 2);
}
