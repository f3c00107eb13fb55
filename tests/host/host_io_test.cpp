// TEST HARNESS for metabuli_b200/csrc/host/fastx_tsv.hpp (the C++ host's parallel FASTA/FASTQ reader and TSV row formatter):
// C entry points so the CPU tests can compare them with the Python mirror (metabuli_b200/fastx.py, Classifier.format_tsv).
#include "../../metabuli_b200/csrc/host/fastx_tsv.hpp"
#include "../../metabuli_b200/csrc/host/fast_inflate.hpp"

namespace {
mblhost::ReadSet g_reads;
std::string g_text;
struct Tax {
    const int32_t* orig; const char* const* ranks; const char* const* lineages;
    int32_t original(int32_t x) const { return orig[x]; }
    const char* rank_name(int32_t x) const { return ranks[x]; }
    std::string lineage(int32_t x) const { return lineages ? lineages[x] : ""; }
};
}  // namespace

extern "C" {
// -> number of reads, or -1 (error text via hio_text)
long long hio_load(const char* path, unsigned threads) {
    std::string err;
    if (!mblhost::load_fastx(path, g_reads, threads, &err)) { g_text = err; return -1; }
    return (long long)g_reads.size();
}
// the same file through the streaming reader in batches of batch_reads records and raw chunks of chunk_bytes, concatenated
long long hio_load_stream(const char* path, unsigned threads, unsigned long long batch_reads, unsigned long long chunk_bytes) {
    std::string err;
    mblhost::FastxStream st;
    if (!st.open(path, &err)) { g_text = err; return -1; }
    g_reads = mblhost::ReadSet();
    mblhost::ReadSet b;
    while (true) {
        if (!st.next(b, (size_t)batch_reads, threads, &err, (size_t)chunk_bytes)) { g_text = err; return -1; }
        if (b.size() == 0) break;
        if (b.size() > batch_reads) { g_text = "batch larger than requested"; return -1; }
        const uint64_t base = g_reads.bases.size();
        for (auto& n : b.names) g_reads.names.push_back(n);
        g_reads.bases.insert(g_reads.bases.end(), b.bases.begin(), b.bases.end());
        for (size_t k = 1; k < b.offsets.size(); ++k) g_reads.offsets.push_back(base + b.offsets[k]);
    }
    return (long long)g_reads.size();
}
unsigned long long hio_total_bases() { return g_reads.bases.size(); }
void hio_copy(char* bases, unsigned long long* offsets) {
    if (!g_reads.bases.empty()) memcpy(bases, g_reads.bases.data(), g_reads.bases.size());
    memcpy(offsets, g_reads.offsets.data(), 8 * g_reads.offsets.size());
}
const char* hio_name(unsigned long long i) { return g_reads.names[i].c_str(); }
// formats rows [0, n) of the loaded reads' names; -> length of the text (hio_text)
unsigned long long hio_format(unsigned long long n, const mbl_read_result* res, const int32_t* pairs, const int32_t* orig,
                              const char* const* ranks, unsigned threads, const char* const* lineages) {
    Tax t{orig, ranks, lineages};
    std::vector<std::string> rows;
    mblhost::format_rows(t, g_reads.names, 0, (size_t)n, res, pairs, threads, rows, lineages != nullptr);
    g_text.clear();
    for (auto& r : rows) g_text += r;
    return g_text.size();
}
// GzInflater on a gzip stream in memory, read in pieces of `piece` bytes (0 = one call for everything): 0 when the output equals
// want[0, want_n), 1 on a decoder error (text via hio_text), 2 on different output, 3 when more bytes came than expected
int hio_inflate_check(const unsigned char* gz, unsigned long long n, const unsigned char* want, unsigned long long want_n, unsigned long long piece) {
    mblhost::GzInflater inf;
    inf.attach(gz, (size_t)n);
    std::vector<char> out((size_t)want_n + 1024);
    size_t got = 0;
    while (true) {
        const size_t ask = piece ? std::min<size_t>((size_t)piece, out.size() - got) : out.size() - got;
        if (!ask) return 3;
        const long long r = inf.read(out.data() + got, ask);
        if (r < 0) { g_text = inf.error(); return 1; }
        got += (size_t)r;
        if ((size_t)r < ask) break;
    }
    if (got != want_n || (want_n && memcmp(out.data(), want, (size_t)want_n) != 0)) return 2;
    return 0;
}
// GzParallel on a gzip stream in memory with `threads` decoder threads, taken in calls of about `want` bytes: same return codes
int hio_parallel_inflate_check(const unsigned char* gz, unsigned long long n, const unsigned char* want, unsigned long long want_n,
                               unsigned threads, unsigned long long piece) {
    mblhost::GzParallel inf;
    inf.attach(gz, (size_t)n);
    mblhost::ByteBuf out;
    size_t got = 0;
    while (true) {
        const long long r = inf.read_some(out, got, (size_t)piece, threads);
        if (r < 0) { g_text = inf.error(); return 1; }
        if (r == 0) break;
        got += (size_t)r;
        if (got > want_n + 1024) return 3;
    }
    g_text = "groups " + std::to_string(inf.groups()) + " segments " + std::to_string(inf.segments()) + " false " + std::to_string(inf.false_starts()) + " bgzf " + std::to_string(inf.bgzf_groups());
    if (got != want_n || (want_n && memcmp(out.data(), want, (size_t)want_n) != 0)) return 2;
    return 0;
}
// put_float_g against printf("%g") on every `stride`-th float of [1e-4, 10) (plus a margin into the snprintf fallback on both
// sides) and on a list of values outside it; -> number of mismatches
unsigned long long hio_check_float_g(unsigned stride, unsigned threads) {
    float lo = 1e-4f, hi = 10.0f;
    uint32_t a, b;
    memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
    a -= 1000; b += 1000;
    std::vector<unsigned long long> bad(threads ? threads : 1, 0);
    const unsigned T = (unsigned)bad.size();
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] {
        char x[64], y[64];
        for (uint64_t u = (uint64_t)a + (uint64_t)t * stride; u <= b; u += (uint64_t)T * stride) {
            const uint32_t v = (uint32_t)u;
            float f;
            memcpy(&f, &v, 4);
            *mblhost::put_float_g(x, f) = 0;
            snprintf(y, sizeof y, "%g", (double)f);
            if (strcmp(x, y)) ++bad[t];
        }
    });
    for (auto& x : th) x.join();
    unsigned long long n = 0;
    for (auto v : bad) n += v;
    const float odd[] = {0.0f, -0.0f, 1e-5f, 9.99999e-5f, 10.0f, 123456.0f, 1e6f, 1e7f, -0.5f, 1e-30f, 3.4e38f, NAN, INFINITY, 0.99999964f, 0.9999995f, 9.999996f};
    for (float f : odd) {
        char x[64], y[64];
        *mblhost::put_float_g(x, f) = 0;
        snprintf(y, sizeof y, "%g", (double)f);
        if (strcmp(x, y)) ++n;
    }
    return n;
}
// <jobid>_report.tsv from per-read classifications (internal taxids) and the taxonomy arrays
unsigned long long hio_report(unsigned long long n_reads, const int32_t* classification, unsigned long long n_nodes, int32_t max_taxid,
                              const int32_t* node_taxid, const int32_t* node_parent, const int32_t* D, const int32_t* orig,
                              const char* const* node_rank, const char* const* node_name) {
    mblhost::ArrayTax t{(size_t)n_nodes, max_taxid, node_taxid, node_parent, D, orig, node_rank, node_name};
    std::vector<uint64_t> counts((size_t)max_taxid + 1, 0);
    for (unsigned long long i = 0; i < n_reads; ++i) ++counts[(size_t)classification[i]];
    g_text.clear();
    mblhost::write_report(t, counts, n_reads, g_text);
    return g_text.size();
}
const char* hio_text() { return g_text.c_str(); }
}
