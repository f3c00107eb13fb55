// Host I/O of the classify path: FASTA/FASTQ reading with kseq semantics (reference: lib/mmseqs KSeqWrapper as used by
// KmerExtractor::loadChunkOfReads, KmerExtractor.cpp:429-481, and QueryIndexer::indexQueryFile, QueryIndexer.cpp:30-147) and
// the per-read TSV rows of Reporter::writeReadClassification (Reporter.cpp:35-80).
//
// The reference parses on one thread (its master thread is the bottleneck at 8 threads, SURVEY §8 A3') and formats on one
// thread; at tens of millions of reads per second from the GPU both have to be parallel.  The file is cut at record starts
// into one range per thread, every range is parsed with the same sequential grammar, and the rows of a batch are formatted
// by all threads into per-thread strings that are written in order.
#pragma once
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/metabuli_b200.h"
#include "fast_inflate.hpp"

namespace mblhost {

// growable byte array without value-initialisation (std::vector<char>::resize would zero-fill hundreds of megabytes per batch)
class ByteBuf {
public:
    ByteBuf() = default;
    ByteBuf(const ByteBuf& o) { append(o.p_, o.n_); }
    ByteBuf(ByteBuf&& o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr; o.n_ = o.cap_ = 0; }
    ByteBuf& operator=(const ByteBuf& o) { if (this != &o) { n_ = 0; append(o.p_, o.n_); } return *this; }
    ByteBuf& operator=(ByteBuf&& o) noexcept { if (this != &o) { free(p_); p_ = o.p_; n_ = o.n_; cap_ = o.cap_; o.p_ = nullptr; o.n_ = o.cap_ = 0; } return *this; }
    ~ByteBuf() { free(p_); }
    char* data() { return p_; }
    const char* data() const { return p_; }
    size_t size() const { return n_; }
    size_t capacity() const { return cap_; }
    bool empty() const { return n_ == 0; }
    char* begin() { return p_; }
    char* end() { return p_ + n_; }
    const char* begin() const { return p_; }
    const char* end() const { return p_ + n_; }
    char& operator[](size_t i) { return p_[i]; }
    const char& operator[](size_t i) const { return p_[i]; }
    void clear() { n_ = 0; }
    void release() { free(p_); p_ = nullptr; n_ = cap_ = 0; }
    void reserve(size_t c) {
        if (c <= cap_) return;
        char* q = static_cast<char*>(realloc(p_, c));
        if (!q) throw std::bad_alloc();
        p_ = q; cap_ = c;
    }
    void resize(size_t n) { if (n > cap_) reserve(n); n_ = n; }              // new bytes are NOT initialised
    void append(const char* src, size_t len) {
        if (!len) return;
        if (n_ + len > cap_) reserve(std::max(n_ + len, cap_ + cap_ / 2 + 64));
        memcpy(p_ + n_, src, len);
        n_ += len;
    }
    void insert(const char* at, const char* b, const char* e) { (void)at; append(b, (size_t)(e - b)); }   // at == end() only
    void assign(const char* b, const char* e) { n_ = 0; append(b, (size_t)(e - b)); }
private:
    char* p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

struct ReadSet {                       // SoA layout of mbl_batch + the names the Reporter prints
    std::vector<std::string> names;
    ByteBuf bases;
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

// kseq keeps the graphic characters of a sequence line (isgraph in the C locale: 33..126)
inline bool all_graphic(const char* p, size_t len) {
    unsigned bad = 0;
    for (size_t x = 0; x < len; ++x) bad |= (unsigned)((unsigned char)(p[x] - 33) > 93u);
    return bad == 0;
}

// sequential kseq grammar over d[i0, i1): a record starts at a line whose first character is '>' or '@'; name = first
// whitespace-delimited token; sequence = the graphic characters of the following lines up to a line starting with '>', '@'
// or '+'; after a '+' line as many quality characters as the record has bases are skipped (for '>' records too, as kseq does).
// Returns false on an entry without a sequence or a name (QueryIndexer.cpp:50-53); *bad = its ordinal in the range.
inline bool parse_range(const char* d, size_t i0, size_t i1, ReadSet& out, size_t* bad) {
    size_t i = i0;
    const size_t n = i1;
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
            if (all_graphic(d + b, e - b)) {
                out.bases.append(d + b, e - b);
            } else {                                         // rare: blanks or control characters inside a sequence line
                for (size_t x = b; x < e; ++x) if (isgraph((unsigned char)d[x])) out.bases.append(d + x, 1);
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
inline size_t next_record_start(const char* d, size_t n, size_t from, bool fastq) {
    size_t i = from;
    if (i > 0) {                                              // move to a line start
        if (i - 1 >= n) return n;
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

// d[begin, end) — a whole number of records — parsed by up to T threads (record-aligned cuts, the sequential grammar per range)
// into one ReadSet per range, APPENDED to parts in file order (empty ranges are dropped).  Returns false on an entry without
// sequence / name; *bad = its ordinal among the entries of [begin, end).
inline bool parse_parts(const char* d, size_t begin, size_t end, bool fastq, unsigned threads, std::vector<ReadSet>& parts, size_t* bad) {
    unsigned T = threads ? threads : 1;
    if (end - begin < (1u << 22)) T = 1;
    std::vector<size_t> cut(T + 1, end);
    cut[0] = begin;
    for (unsigned t = 1; t < T; ++t) cut[t] = std::min(end, next_record_start(d, end, begin + (end - begin) / T * t, fastq));
    for (unsigned t = 1; t <= T; ++t) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    std::vector<ReadSet> part(T);
    std::vector<size_t> badv(T, 0);
    std::vector<char> ok(T, 1);
    auto work = [&](unsigned t) {
        const size_t len = cut[t + 1] - cut[t];
        part[t].bases.reserve(len / (fastq ? 2 : 1) + 64);
        const size_t guess = len / (fastq ? 320 : 170) + 16;             // records of ~150 letters; vectors grow if it is off
        part[t].names.reserve(guess);
        part[t].offsets.reserve(guess + 1);
        ok[t] = parse_range(d, cut[t], cut[t + 1], part[t], &badv[t]) ? 1 : 0;
    };
    Workers th;
    th.reserve(T);
    for (unsigned t = 1; t < T; ++t) th.spawn([&work, t] { work(t); });
    work(0);
    th.join();
    size_t n_reads = 0;
    for (unsigned t = 0; t < T; ++t) {
        if (!ok[t]) { if (bad) *bad = n_reads + badv[t]; return false; }
        n_reads += part[t].size();
    }
    for (unsigned t = 0; t < T; ++t) if (part[t].size()) parts.emplace_back(std::move(part[t]));
    return true;
}

// reads [first, first + n) of the concatenation of parts[p0..] (starting at read r0 of parts[p0]) gathered into out by up to T
// threads: names are moved out of the parts, letters copied, offsets rebased.  out's previous content is replaced.
inline void gather_reads(std::vector<ReadSet>& parts, size_t p0, size_t r0, size_t n, unsigned threads, ReadSet& out) {
    struct Seg { size_t part, r0, r1, out_read; uint64_t out_base; };
    std::vector<Seg> segs;
    size_t left = n, p = p0, r = r0, out_read = 0;
    uint64_t out_base = 0;
    while (left) {
        const size_t take = std::min(left, parts[p].size() - r);
        if (take) {
            segs.push_back(Seg{p, r, r + take, out_read, out_base});
            out_base += parts[p].offsets[r + take] - parts[p].offsets[r];
            out_read += take; left -= take;
        }
        ++p; r = 0;
    }
    out.names.clear();
    out.names.resize(n);
    out.offsets.resize(n + 1);
    out.offsets[0] = 0;
    if (out_base > out.bases.capacity()) {                    // grow with head room so that a steady stream of batches stops reallocating
        if (out.before_realloc) out.before_realloc();
        out.bases.release();
        out.bases.reserve((size_t)out_base + (size_t)out_base / 8 + 4096);
    }
    out.bases.resize((size_t)out_base);
    // every segment is cut further so that all threads get work even when one part holds the whole batch
    unsigned T = std::max(1u, threads);
    if (n < 65536) T = 1;
    struct Task { size_t seg, r0, r1; };
    std::vector<Task> tasks;
    const size_t per = std::max<size_t>(4096, (n + T - 1) / T);
    for (size_t k = 0; k < segs.size(); ++k)
        for (size_t a = segs[k].r0; a < segs[k].r1; a += per) tasks.push_back(Task{k, a, std::min(segs[k].r1, a + per)});
    auto run = [&](size_t t0, size_t t1) {
        for (size_t k = t0; k < t1; ++k) {
            const Task& tk = tasks[k];
            const Seg& sg = segs[tk.seg];
            ReadSet& src = parts[sg.part];
            const uint64_t src0 = src.offsets[sg.r0];
            const size_t dst_read = sg.out_read + (tk.r0 - sg.r0);
            for (size_t i = tk.r0; i < tk.r1; ++i) {
                out.names[dst_read + (i - tk.r0)] = std::move(src.names[i]);
                out.offsets[dst_read + (i - tk.r0) + 1] = sg.out_base + (src.offsets[i + 1] - src0);
            }
            memcpy(out.bases.data() + sg.out_base + (src.offsets[tk.r0] - src0), src.bases.data() + src.offsets[tk.r0],
                   (size_t)(src.offsets[tk.r1] - src.offsets[tk.r0]));
        }
    };
    T = (unsigned)std::min<size_t>(T, std::max<size_t>(1, tasks.size()));
    Workers th;
    th.reserve(T);
    for (unsigned t = 1; t < T; ++t) th.spawn([&run, &tasks, t, T] { run(tasks.size() * t / T, tasks.size() * (t + 1) / T); });
    run(0, tasks.size() / T);
    th.join();
}

inline bool sniff_fastq(const char* d, size_t n) {
    size_t first = 0;
    while (first < n && d[first] != '>' && d[first] != '@') {
        const void* nl = memchr(d + first, '\n', n - first);
        if (!nl) return false;
        first = (size_t)((const char*)nl - d) + 1;
    }
    return first < n && d[first] == '@';
}

// Whole-file parallel load (tests, small inputs).  Returns false with *err set on unreadable files or an entry without
// sequence / name.
inline bool load_fastx(const std::string& path, ReadSet& out, unsigned threads, std::string* err) {
    std::string data;
    if (!slurp_maybe_gz(path, data)) { if (err) *err = "cannot open " + path; return false; }
    out.names.clear(); out.bases.clear(); out.offsets.assign(1, 0);
    size_t bad = 0;
    std::vector<ReadSet> parts;
    if (!parse_parts(data.data(), 0, data.size(), sniff_fastq(data.data(), data.size()), threads, parts, &bad)) {
        if (err) *err = std::to_string(bad) + "th entry has no sequence or name.";
        return false;
    }
    size_t n = 0;
    for (auto& p : parts) n += p.size();
    if (n) gather_reads(parts, 0, 0, n, threads, out);
    return true;
}

// Streaming reader: the file is read (and decompressed) in bounded chunks, cut at record starts, parsed by all threads, and
// handed out in batches of at most max_reads records — the reference's QuerySplit loop bounded by --max-ram
// (QueryIndexer.cpp:30-147, KmerExtractor.cpp:429-481) instead of a whole-file load.  While one chunk is parsed the next one is
// already being read by a helper thread (plain files through read(2), gzip through zlib), the parsed ranges stay where the
// parser threads wrote them and a batch is gathered from them by all threads: what is held at any time is two raw chunks, the
// records parsed but not handed out yet, and the batches in flight.
class FastxStream {
public:
    ~FastxStream() {
        if (io_.joinable()) io_.join();
        if (fd_ >= 0) ::close(fd_);
    }
    bool open(const std::string& path, std::string* err) {
        // gzip files go through the reader's own decoder (fast_inflate.hpp: one stream decoded by several threads, CRC-checked, no
        // fallback to zlib), anything else is read as it is
        if (GzInflater::looks_gzip(path)) { gz_ = true; return inf_.open(path, err); }
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) { if (err) *err = "cannot open " + path; return false; }
        return true;
    }
    // out is cleared and receives the next min(max_reads, remaining) records; out.size() == 0 <=> end of file
    bool next(ReadSet& out, size_t max_reads, unsigned threads, std::string* err, size_t chunk_bytes = (size_t)64 << 20) {
        io_threads_.store(std::max(1u, threads));
        while (pending_ < max_reads && !done_) {
            // the chunk being completed: what is left of the previous one (an incomplete record) + newly read bytes
            if (!started_) { start_read(0, chunk_bytes); started_ = true; }
            const size_t have = finish_read();                // bytes in raw_[cur_]
            if (io_failed_) { if (err) *err = gz_ ? inf_.error() : std::string("read error"); return false; }
            const char* d = raw_[cur_].data();
            if (!sniffed_ && have) { fastq_ = sniff_fastq(d, have); sniffed_ = true; }
            size_t cut = have;
            if (!eof_) {                                      // keep the (possibly incomplete) last record for the next round
                cut = last_record_start(d, have, fastq_);
                if (cut == 0) {                               // a single record longer than the chunk: read on into the same buffer
                    start_read_append(have, chunk_bytes);
                    continue;
                }
            }
            // hand the tail to the other buffer and let the helper thread fill the rest of it while this chunk is parsed
            const size_t tail = have - cut;
            const int nxt = cur_ ^ 1;
            if (!eof_) {
                raw_[nxt].resize(tail + chunk_bytes);
                if (tail) memcpy(raw_[nxt].data(), d + cut, tail);
                const int was = cur_;
                cur_ = nxt;
                start_read(tail, chunk_bytes);
                cur_ = was;
            }
            size_t bad = 0;
            const size_t before = parts_.size();
            if (!parse_parts(d, 0, cut, fastq_, threads, parts_, &bad)) {
                if (err) *err = std::to_string(parsed_ + bad) + "th entry has no sequence or name.";
                return false;
            }
            for (size_t k = before; k < parts_.size(); ++k) { pending_ += parts_[k].size(); parsed_ += parts_[k].size(); }
            if (eof_) done_ = true; else cur_ = nxt;
        }
        out.names.clear(); out.bases.clear(); out.offsets.assign(1, 0);
        const size_t n = std::min(max_reads, pending_);
        if (n) {
            gather_reads(parts_, part0_, read0_, n, threads, out);
            pending_ -= n;
            // advance the cursor; parts that are used up give their memory back
            size_t left = n;
            while (left) {
                const size_t take = std::min(left, parts_[part0_].size() - read0_);
                left -= take; read0_ += take;
                if (read0_ == parts_[part0_].size()) { parts_[part0_] = ReadSet(); ++part0_; read0_ = 0; }
            }
            if (part0_ == parts_.size()) { parts_.clear(); part0_ = 0; }
            else if (part0_ >= 256) { parts_.erase(parts_.begin(), parts_.begin() + (ptrdiff_t)part0_); part0_ = 0; }   // drop the used-up slots
        }
        return true;
    }

private:
    // reads up to want bytes to raw_[cur_] at offset at (helper thread); finish_read() joins it and returns the bytes the buffer holds
    void start_read(size_t at, size_t want) {
        ByteBuf& b = raw_[cur_];
        if (b.size() < at + want) b.resize(at + want);
        fill_at_ = at;
        io_ = std::thread([this, &b, at, want] {
            size_t got = 0;
            try {
            while (got < want) {
                const size_t ask = std::min<size_t>(want - got, 1u << 30);
                // gzip: whole groups of blocks, decoded by io_threads_ threads (the buffer grows when a group is larger than asked)
                const long long r = gz_ ? inf_.read_some(b, at + got, ask, io_threads_.load()) : (long long)::read(fd_, b.data() + at + got, ask);
                if (r < 0) { io_failed_ = true; eof_io_ = true; break; }
                if (r == 0) { eof_io_ = true; break; }
                got += (size_t)r;
            }
            } catch (...) { io_error_ = std::current_exception(); eof_io_ = true; }
            got_ = got;
        });
    }
    void start_read_append(size_t have, size_t want) {        // grow the current buffer (keeps its content) and read on
        ByteBuf& b = raw_[cur_];
        b.resize(have);                                       // size = valid bytes, so that the reallocation copies exactly those
        b.reserve(have + want);
        b.resize(have + want);
        start_read(have, want);
    }
    size_t finish_read() {
        if (io_.joinable()) io_.join();
        if (io_error_) { std::exception_ptr x = io_error_; io_error_ = nullptr; std::rethrow_exception(x); }
        if (eof_io_) eof_ = true;
        return fill_at_ + got_;
    }
    // the last safe record start of d[0, n) (0 when there is none beyond the first)
    static size_t last_record_start(const char* d, size_t n, bool fastq) {
        for (size_t window = 1u << 16;; window <<= 2) {
            const size_t from = n > window ? n - window : 1;
            size_t p = next_record_start(d, n, from, fastq);
            if (p < n) {
                for (size_t q = next_record_start(d, n, p + 1, fastq); q < n; q = next_record_start(d, n, q + 1, fastq)) p = q;
                return p;
            }
            if (from == 1) return 0;
        }
    }
    GzParallel inf_;
    std::atomic<unsigned> io_threads_{1};
    bool gz_ = false, io_failed_ = false;
    int fd_ = -1;
    ByteBuf raw_[2];
    int cur_ = 0;
    std::thread io_;
    std::exception_ptr io_error_;
    size_t fill_at_ = 0, got_ = 0;
    bool eof_io_ = false, eof_ = false, done_ = false, started_ = false, fastq_ = false, sniffed_ = false;
    std::vector<ReadSet> parts_;
    size_t part0_ = 0, read0_ = 0, pending_ = 0, parsed_ = 0;
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
inline char* put_int(char* p, long long v) {              // raw-pointer twin of append_int; the caller guarantees 21 bytes
    char buf[24];
    int n = 0;
    const bool neg = v < 0;
    unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) *p++ = '-';
    while (n) *p++ = buf[--n];
    return p;
}

// printf("%g") of a float (what `ostream << float` prints, Reporter.cpp:62), exact: for 1e-4 <= v < 10 — every score the
// classifier produces — the six significant digits come from integer arithmetic on the float's own mantissa (v * 10^s is a
// 54-bit integer over a power of two, rounded half-to-even like glibc does on the exact decimal expansion); anything else goes
// through snprintf.  tests/host: checked against snprintf on EVERY float of the fast range.  The caller guarantees 48 bytes.
inline char* put_float_g(char* p, float f) {
    if (f == 0.0f && !std::signbit(f)) { *p++ = '0'; return p; }
    if (!(f >= 1e-4f && f < 10.0f)) return p + snprintf(p, 48, "%g", (double)f);
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint64_t m = (u & 0x7FFFFFu) | 0x800000u;       // normal numbers only in this range
    const int e = (int)((u >> 23) & 0xFF) - 150;           // f = m * 2^e, e in [-37, -20]
    const double d = (double)f;                            // decade: no float equals a power of ten below 1, so these compares are exact
    int X = d >= 1.0 ? 0 : d >= 0.1 ? -1 : d >= 0.01 ? -2 : d >= 0.001 ? -3 : -4;
    static const uint64_t p10[10] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull};
    const uint64_t P = m * p10[5 - X];
    const int k = -e;
    uint64_t q = P >> k;
    const uint64_t r = P & ((1ull << k) - 1), half = 1ull << (k - 1);
    if (r > half || (r == half && (q & 1))) ++q;
    if (q >= 1000000ull) { q = 100000ull; ++X; }           // 0.9999996 -> 1
    char dig[6];
    for (int i = 5; i >= 0; --i) { dig[i] = (char)('0' + q % 10); q /= 10; }
    int last = 5;
    while (last > 0 && dig[last] == '0') --last;           // %g drops trailing zeros
    if (X >= 0) {                                          // X is 0 or 1 here
        for (int i = 0; i <= X; ++i) *p++ = dig[i];
        if (last > X) { *p++ = '.'; for (int i = X + 1; i <= last; ++i) *p++ = dig[i]; }
    } else {
        *p++ = '0'; *p++ = '.';
        for (int i = 0; i < -X - 1; ++i) *p++ = '0';
        for (int i = 0; i <= last; ++i) *p++ = dig[i];
    }
    return p;
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
        // rows are written through a raw pointer into a block that is grown ahead of every row by the row's upper bound
        ByteBuf buf;
        buf.reserve((b - a) * 72 + 4096);
        size_t len = 0;
        std::string lin;
        for (size_t i = a; i < b; ++i) {
            const mbl_read_result& q = res[i];
            const std::string& name = names[name0 + i];
            const char* rank = q.is_classified ? tax.rank_name(q.classification) : "";
            const size_t rank_len = strlen(rank);
            if (lineage && q.is_classified) lin = tax.lineage(q.classification); else lin.clear();
            const size_t need = name.size() + rank_len + lin.size() + 128 + 34 * (size_t)q.taxcnt_len;
            if (len + need > buf.capacity()) { buf.resize(len); buf.reserve(std::max(len + need, buf.capacity() + buf.capacity() / 2)); }
            char* p = buf.data() + len;
            *p++ = q.is_classified ? '1' : '0';
            *p++ = '\t';
            memcpy(p, name.data(), name.size()); p += name.size();
            *p++ = '\t';
            p = put_int(p, tax.original(q.classification));
            *p++ = '\t';
            p = put_int(p, q.query_length);
            *p++ = '\t';
            p = put_float_g(p, q.score);
            *p++ = '\t';
            if (q.is_classified) {
                memcpy(p, rank, rank_len); p += rank_len;
                *p++ = '\t';
                if (lineage) { memcpy(p, lin.data(), lin.size()); p += lin.size(); *p++ = '\t'; }
                for (uint32_t k = q.taxcnt_begin; k < q.taxcnt_begin + q.taxcnt_len; ++k) {
                    p = put_int(p, tax.original(pairs[2 * (size_t)k]));
                    *p++ = ':';
                    p = put_int(p, pairs[2 * (size_t)k + 1]);
                    *p++ = ' ';
                }
                *p++ = '\n';
            } else {
                const char* tail = lineage ? "-\t-\t-\t\n" : "-\t-\t\n";
                const size_t tl = lineage ? 7 : 5;
                memcpy(p, tail, tl); p += tl;
            }
            len = (size_t)(p - buf.data());
        }
        out[t].assign(buf.data(), len);
    };
    Workers th;
    th.reserve(T);
    for (unsigned t = 1; t < T; ++t) th.spawn([&work, t] { work(t); });
    work(0);
    th.join();
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
