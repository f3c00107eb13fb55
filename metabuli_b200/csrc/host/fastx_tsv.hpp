// Host I/O of the classify path: FASTA/FASTQ reading with kseq semantics (reference: lib/mmseqs KSeqWrapper as used by
// KmerExtractor::loadChunkOfReads, KmerExtractor.cpp:429-481, and QueryIndexer::indexQueryFile, QueryIndexer.cpp:30-147) and
// the per-read TSV rows of Reporter::writeReadClassification (Reporter.cpp:35-80).
//
// The reference parses on one thread (its master thread is the bottleneck at 8 threads, SURVEY §8 A3') and formats on one
// thread; at tens of millions of reads per second from the GPU both have to be parallel.  The file is cut at record starts
// into one range per thread, every range is parsed with the same sequential grammar, and the rows of a batch are formatted
// by all threads into per-thread strings that are written in order.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/metabuli_b200.h"

namespace mblhost {

struct ReadSet {                       // SoA layout of mbl_batch + the names the Reporter prints
    std::vector<std::string> names;
    std::vector<char> bases;
    std::vector<uint64_t> offsets{0};
    std::function<void()> before_realloc;   // called before `bases` moves to a larger block (the host unpins the old one)
    size_t size() const { return names.size(); }
};

// whole file into memory (gz or plain; gzread handles both)
inline bool slurp_maybe_gz(const std::string& path, std::string& data) {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    std::vector<char> buf(1 << 24);
    int got;
    while ((got = gzread(g, buf.data(), (unsigned)buf.size())) > 0) data.append(buf.data(), (size_t)got);
    gzclose(g);
    return true;
}

// sequential kseq grammar over data[i0, i1): a record starts at a line whose first character is '>' or '@'; name = first
// whitespace-delimited token; sequence = the graphic characters of the following lines up to a line starting with '>', '@'
// or '+'; after a '+' line as many quality characters as the record has bases are skipped (for '>' records too, as kseq does).
// Returns false on an entry without a sequence or a name (QueryIndexer.cpp:50-53); *bad = its ordinal in the range.
inline bool parse_range(const std::string& data, size_t i0, size_t i1, ReadSet& out, size_t* bad) {
    size_t i = i0;
    const size_t n = i1;
    const char* d = data.data();
    auto line = [&](size_t& b, size_t& e) {
        b = i;
        const void* nl = i < n ? memchr(d + i, '\n', n - i) : nullptr;
        e = nl ? (size_t)((const char*)nl - d) : n;
        i = nl ? e + 1 : n;
        if (e > b && d[e - 1] == '\r') --e;
    };
    size_t b, e, entry = 0;
    while (i < n) {
        while (i < n && d[i] != '>' && d[i] != '@') line(b, e);
        if (i >= n) break;
        line(b, e);
        size_t p = b + 1;
        while (p < e && !isspace((unsigned char)d[p])) ++p;
        out.names.emplace_back(d + b + 1, p - (b + 1));
        const size_t start = out.bases.size();
        while (i < n && d[i] != '>' && d[i] != '@' && d[i] != '+') {
            line(b, e);
            const size_t at = out.bases.size();
            out.bases.insert(out.bases.end(), d + b, d + e);
            bool clean = true;
            for (size_t x = at; x < out.bases.size(); ++x) clean &= isgraph((unsigned char)out.bases[x]) != 0;
            if (!clean) {                                    // rare: blanks or control characters inside a sequence line
                size_t w = at;
                for (size_t x = at; x < out.bases.size(); ++x) if (isgraph((unsigned char)out.bases[x])) out.bases[w++] = out.bases[x];
                out.bases.resize(w);
            }
        }
        if (i < n && d[i] == '+') {                          // kseq reads qualities after a '+' line whatever the record's tag was
            line(b, e);
            size_t ql = 0;
            const size_t sl = out.bases.size() - start;
            while (i < n && ql < sl) { line(b, e); ql += e - b; }
        }
        ++entry;
        if (out.bases.size() == start || out.names.back().empty()) { if (bad) *bad = entry; return false; }
        out.offsets.push_back(out.bases.size());
    }
    return true;
}

// record starts at or after `from` that are safe cut points: FASTA: a line starting with '>'; FASTQ: a line starting with
// '@' whose second-next line starts with '+' (a quality line that starts with '@' is followed by a header, then bases)
inline size_t next_record_start(const std::string& data, size_t from, bool fastq) {
    const size_t n = data.size();
    const char* d = data.data();
    size_t i = from;
    if (i > 0) {                                              // move to a line start
        const void* nl = memchr(d + i - 1, '\n', n - (i - 1));
        if (!nl) return n;
        i = (size_t)((const char*)nl - d) + 1;
    }
    while (i < n) {
        if (!fastq && d[i] == '>') return i;
        if (fastq && d[i] == '@') {
            const void* l1 = memchr(d + i, '\n', n - i);
            const void* l2 = l1 ? memchr((const char*)l1 + 1, '\n', n - ((const char*)l1 + 1 - d)) : nullptr;
            if (l2 && (size_t)((const char*)l2 + 1 - d) < n && ((const char*)l2)[1] == '+') return i;
        }
        const void* nl = memchr(d + i, '\n', n - i);
        if (!nl) return n;
        i = (size_t)((const char*)nl - d) + 1;
    }
    return n;
}

// data[begin, end) — a whole number of records — parsed by up to T threads (record-aligned cuts, the sequential grammar per
// range) and APPENDED to out.  Returns false on an entry without sequence / name; *bad = its ordinal among the entries of the range.
inline bool parse_parallel(const std::string& data, size_t begin, size_t end, bool fastq, unsigned threads, ReadSet& out, size_t* bad) {
    unsigned T = threads ? threads : 1;
    if (end - begin < (1u << 22)) T = 1;
    std::vector<size_t> cut(T + 1, end);
    cut[0] = begin;
    for (unsigned t = 1; t < T; ++t) cut[t] = std::min(end, next_record_start(data, begin + (end - begin) / T * t, fastq));
    for (unsigned t = 1; t <= T; ++t) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    std::vector<ReadSet> part(T);
    std::vector<size_t> badv(T, 0);
    std::vector<char> ok(T, 1);
    auto work = [&](unsigned t) {
        part[t].bases.reserve((cut[t + 1] - cut[t]) / (fastq ? 2 : 1) + 64);
        ok[t] = parse_range(data, cut[t], cut[t + 1], part[t], &badv[t]) ? 1 : 0;
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    size_t n_reads = 0, n_bases = 0;
    for (unsigned t = 0; t < T; ++t) {
        if (!ok[t]) { if (bad) *bad = n_reads + badv[t]; return false; }
        n_reads += part[t].size(); n_bases += part[t].bases.size();
    }
    const size_t base0 = out.bases.size();
    out.names.reserve(out.names.size() + n_reads);
    out.bases.resize(base0 + n_bases);
    out.offsets.reserve(out.offsets.size() + n_reads);
    std::vector<size_t> base_at(T + 1, base0);
    for (unsigned t = 0; t < T; ++t) base_at[t + 1] = base_at[t] + part[t].bases.size();
    for (unsigned t = 0; t < T; ++t) {
        for (auto& s : part[t].names) out.names.emplace_back(std::move(s));
        for (size_t k = 1; k < part[t].offsets.size(); ++k) out.offsets.push_back(base_at[t] + part[t].offsets[k]);
    }
    auto copy = [&](unsigned t) { if (!part[t].bases.empty()) memcpy(out.bases.data() + base_at[t], part[t].bases.data(), part[t].bases.size()); };
    th.clear();
    for (unsigned t = 1; t < T; ++t) th.emplace_back(copy, t);
    copy(0);
    for (auto& x : th) x.join();
    return true;
}

inline bool sniff_fastq(const std::string& data) {
    size_t first = 0;
    while (first < data.size() && data[first] != '>' && data[first] != '@') {
        const void* nl = memchr(data.data() + first, '\n', data.size() - first);
        if (!nl) return false;
        first = (size_t)((const char*)nl - data.data()) + 1;
    }
    return first < data.size() && data[first] == '@';
}

// Whole-file parallel load (tests, small inputs).  Returns false with *err set on unreadable files or an entry without
// sequence / name.
inline bool load_fastx(const std::string& path, ReadSet& out, unsigned threads, std::string* err) {
    std::string data;
    if (!slurp_maybe_gz(path, data)) { if (err) *err = "cannot open " + path; return false; }
    out.names.clear(); out.bases.clear(); out.offsets.assign(1, 0);
    size_t bad = 0;
    if (!parse_parallel(data, 0, data.size(), sniff_fastq(data), threads, out, &bad)) {
        if (err) *err = std::to_string(bad) + "th entry has no sequence or name.";
        return false;
    }
    return true;
}

// Streaming reader: the file is decompressed and parsed in bounded chunks, cut at record starts, and handed out in batches of
// at most max_reads records — the reference's QuerySplit loop bounded by --max-ram (QueryIndexer.cpp:30-147,
// KmerExtractor.cpp:429-481) instead of a whole-file load: what is held at any time is one raw chunk, the records parsed from
// it that were not handed out yet, and the batches in flight.
class FastxStream {
public:
    ~FastxStream() { if (g_) gzclose(g_); }
    bool open(const std::string& path, std::string* err) {
        g_ = gzopen(path.c_str(), "rb");
        if (!g_) { if (err) *err = "cannot open " + path; return false; }
        gzbuffer(g_, 1 << 20);
        return true;
    }
    // out is cleared and receives the next min(max_reads, remaining) records; out.size() == 0 <=> end of file
    bool next(ReadSet& out, size_t max_reads, unsigned threads, std::string* err, size_t chunk_bytes = (size_t)256 << 20) {
        while (pend_.size() - pos_ < max_reads && !(eof_ && buf_.empty())) {
            if (!eof_) {
                const size_t old = buf_.size();
                buf_.resize(old + chunk_bytes);
                size_t got = 0;
                while (got < chunk_bytes) {
                    const int r = gzread(g_, &buf_[old + got], (unsigned)std::min<size_t>(chunk_bytes - got, 1u << 30));
                    if (r <= 0) { eof_ = true; break; }
                    got += (size_t)r;
                }
                buf_.resize(old + got);
            }
            if (!sniffed_ && !buf_.empty()) { fastq_ = sniff_fastq(buf_); sniffed_ = true; }
            size_t cut = buf_.size();
            if (!eof_) {                                     // keep the (possibly incomplete) last record for the next round
                cut = last_record_start(buf_, fastq_);
                if (cut == 0) continue;                       // a single record longer than the chunk: read on
            }
            size_t bad = 0;
            if (!parse_parallel(buf_, 0, cut, fastq_, threads, pend_, &bad)) {
                if (err) *err = std::to_string(parsed_ + bad) + "th entry has no sequence or name.";
                return false;
            }
            parsed_ = pend_.size() + consumed_;
            buf_.erase(0, cut);
        }
        out.names.clear(); out.bases.clear(); out.offsets.assign(1, 0);
        const size_t n = std::min(max_reads, pend_.size() - pos_);
        if (n) {
            const uint64_t b0 = pend_.offsets[pos_], b1 = pend_.offsets[pos_ + n];
            out.names.reserve(n);
            for (size_t i = 0; i < n; ++i) out.names.emplace_back(std::move(pend_.names[pos_ + i]));
            if (b1 - b0 > out.bases.capacity()) {             // grow with head room so that a steady stream of batches stops reallocating
                if (out.before_realloc) out.before_realloc();
                std::vector<char>().swap(out.bases);
                out.bases.reserve((size_t)(b1 - b0) + (size_t)(b1 - b0) / 8 + 4096);
            }
            out.bases.assign(pend_.bases.begin() + (ptrdiff_t)b0, pend_.bases.begin() + (ptrdiff_t)b1);
            out.offsets.reserve(n + 1);
            for (size_t i = 1; i <= n; ++i) out.offsets.push_back(pend_.offsets[pos_ + i] - b0);
            pos_ += n;
            if (pos_ == pend_.size()) {                       // everything parsed so far is handed out
                consumed_ += pos_;
                pend_.names.clear(); pend_.bases.clear(); pend_.offsets.assign(1, 0);
                pos_ = 0;
            }
        }
        return true;
    }

private:
    // the last safe record start of data (0 when there is none beyond the first)
    static size_t last_record_start(const std::string& data, bool fastq) {
        const size_t n = data.size();
        for (size_t window = 1u << 16;; window <<= 2) {
            const size_t from = n > window ? n - window : 1;
            size_t p = next_record_start(data, from, fastq);
            if (p < n) {
                for (size_t q = next_record_start(data, p + 1, fastq); q < n; q = next_record_start(data, q + 1, fastq)) p = q;
                return p;
            }
            if (from == 1) return 0;
        }
    }
    gzFile g_ = nullptr;
    std::string buf_;
    bool eof_ = false, fastq_ = false, sniffed_ = false;
    ReadSet pend_;
    size_t pos_ = 0, consumed_ = 0, parsed_ = 0;
};

// ---- Reporter::writeReadClassification rows (Reporter.cpp:43-79, printLineage 0) ------------------------------------------
// `ostream << float` prints like %g (6 significant digits, Q12).
inline void append_int(std::string& s, long long v) {
    char buf[24];
    int n = 0;
    bool neg = v < 0;
    unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) s.push_back('-');
    while (n) s.push_back(buf[--n]);
}

// Tax must provide: int32_t original(int32_t) const; const char* rank_name(int32_t internal_taxid) const; and, for
// lineage = true (--lineage 1), std::string lineage(int32_t internal_taxid) const (TaxonomyWrapper::taxLineage2)
template <class Tax>
void format_rows(const Tax& tax, const std::vector<std::string>& names, size_t name0, size_t n, const mbl_read_result* res,
                 const int32_t* pairs, unsigned threads, std::vector<std::string>& out, bool lineage = false) {
    unsigned T = threads ? threads : 1;
    if (n < 4096) T = 1;
    out.assign(T, std::string());
    auto work = [&](unsigned t) {
        const size_t a = n * t / T, b = n * (t + 1) / T;
        std::string& s = out[t];
        s.reserve((b - a) * 96);
        char num[48];
        for (size_t i = a; i < b; ++i) {
            const mbl_read_result& q = res[i];
            s.push_back(q.is_classified ? '1' : '0');
            s.push_back('\t');
            s += names[name0 + i];
            s.push_back('\t');
            append_int(s, tax.original(q.classification));
            s.push_back('\t');
            append_int(s, q.query_length);
            s.push_back('\t');
            s.append(num, (size_t)snprintf(num, sizeof num, "%g", (double)q.score));
            s.push_back('\t');
            if (q.is_classified) {
                s += tax.rank_name(q.classification);
                s.push_back('\t');
                if (lineage) { s += tax.lineage(q.classification); s.push_back('\t'); }
                for (uint32_t k = q.taxcnt_begin; k < q.taxcnt_begin + q.taxcnt_len; ++k) {
                    append_int(s, tax.original(pairs[2 * (size_t)k]));
                    s.push_back(':');
                    append_int(s, pairs[2 * (size_t)k + 1]);
                    s.push_back(' ');
                }
                s.push_back('\n');
            } else {
                s += lineage ? "-\t-\t-\t\n" : "-\t-\t\n";
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

// ---- Reporter::writeReportFile / writeReport (Reporter.cpp:117-193) with NcbiTaxonomy::getParentToChildren / getCladeCounts
// (NcbiTaxonomy.cpp:504-545): the Kraken-style <jobid>_report.tsv.  counts[t] = reads classified to internal taxid t
// (counts[0] = unclassified).  Children keep node order and are ordered by clade count with std::sort, as the reference does
// (SORT_SERIAL), so ties fall the same way; the walk starts at internal taxid 1 whatever the root is (Q11).
// Tax must provide: size_t n_nodes(); int32_t node_taxid(size_t i), node_parent(size_t i); bool exists(int32_t t);
// int32_t parent_of(int32_t t); int32_t max_taxid(); original(t); rank_name(t); const char* name(t).
// the taxonomy arrays of taxonomyDB in the shape write_report wants (per node: taxid, parent taxid, rank and name strings;
// D = taxid -> node, orig = internal -> original taxid or null)
struct ArrayTax {
    size_t nn; int32_t maxt; const int32_t *ntax, *npar, *D, *orig; const char* const* rank; const char* const* nm;
    size_t n_nodes() const { return nn; }
    int32_t node_taxid(size_t i) const { return ntax[i]; }
    int32_t node_parent(size_t i) const { return npar[i]; }
    int32_t max_taxid() const { return maxt; }
    bool exists(int32_t t) const { return t >= 0 && t <= maxt && D[t] != -1; }
    int32_t parent_of(int32_t t) const { return npar[D[t]]; }
    int32_t original(int32_t t) const { return orig ? orig[t] : t; }
    const char* rank_name(int32_t t) const { return rank[D[t]]; }
    const char* name(int32_t t) const { return nm[D[t]]; }
};

template <class Tax>
void write_report(const Tax& tax, const std::vector<uint64_t>& counts, uint64_t total_reads, std::string& out) {
    const size_t T = (size_t)tax.max_taxid() + 1;
    std::vector<uint64_t> clade(T, 0), own(T, 0);
    std::vector<char> present(T, 0);
    for (size_t t = 0; t < counts.size() && t < T; ++t) {
        if (!counts[t]) continue;
        own[t] = counts[t]; clade[t] += counts[t]; present[t] = 1;
        if (tax.exists((int32_t)t)) {
            int32_t x = (int32_t)t;
            while (tax.parent_of(x) != x && tax.exists(tax.parent_of(x))) {
                x = tax.parent_of(x);
                clade[(size_t)x] += counts[t]; present[(size_t)x] = 1;
            }
        }
    }
    std::vector<std::vector<int32_t>> children(T);
    for (size_t i = 0; i < tax.n_nodes(); ++i)
        if (tax.node_parent(i) != tax.node_taxid(i)) children[(size_t)tax.node_parent(i)].push_back(tax.node_taxid(i));
    char line[512];
    out += "#clade_proportion\tclade_count\ttaxon_count\trank\ttaxID\tname\n";
    if (clade[0] > 0) {
        snprintf(line, sizeof line, "%.4f\t%i\t%i\tno rank\t0\tunclassified\n", 100 * clade[0] / double(total_reads), (int)clade[0], (int)own[0]);
        out += line;
    }
    struct Frame { int32_t t; int depth; };
    // explicit stack: children are pushed in reverse so that they are written in sorted order
    std::vector<Frame> stack;
    if (T > 1) stack.push_back(Frame{1, 0});
    while (!stack.empty()) {
        const Frame f = stack.back();
        stack.pop_back();
        if (!present[(size_t)f.t] || clade[(size_t)f.t] == 0) continue;
        snprintf(line, sizeof line, "%.4f\t%i\t%i\t%s\t%i\t", 100 * clade[(size_t)f.t] / double(total_reads), (int)clade[(size_t)f.t],
                 (int)own[(size_t)f.t], tax.rank_name(f.t), (int)tax.original(f.t));
        out += line;
        out.append((size_t)(2 * f.depth), ' ');
        out += tax.name(f.t);
        out.push_back('\n');
        std::vector<int32_t> ch = children[(size_t)f.t];
        std::sort(ch.begin(), ch.end(), [&](int a, int b) {
            return (present[(size_t)a] ? clade[(size_t)a] : 0) > (present[(size_t)b] ? clade[(size_t)b] : 0);
        });
        size_t keep = 0;
        while (keep < ch.size() && present[(size_t)ch[keep]]) ++keep;          // the reference stops at the first child without reads
        for (size_t k = keep; k > 0; --k) stack.push_back(Frame{ch[k - 1], f.depth + 1});
    }
}

}  // namespace mblhost
