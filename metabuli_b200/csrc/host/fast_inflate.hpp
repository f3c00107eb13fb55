// gzip (RFC 1952) / DEFLATE (RFC 1951) decoder for the streaming FASTA/FASTQ reader.
//
// Why not zlib: reads arrive gzip-compressed almost always, inflate is sequential per file, and zlib's inflate is what bounds the
// C++ host end to end on such input (~200 MB/s of text on the development machine = 1.1 M reads/s).  This decoder is
// written for that one job: the compressed file is mapped (no input refill logic), the bit buffer is 64 bits wide and refilled
// with one unaligned 8-byte load, literals/lengths are decoded from an 11-bit primary table (sub-tables for longer codes), up to
// three literals are emitted per refill, and matches are copied eight bytes at a time.  Output is produced into the caller's
// buffer in exact amounts (a match that does not fit is resumed by the next call), with a 32 KiB history kept across calls.
// Every member's CRC-32 and length are verified (zlib's crc32() is used for that), members may be concatenated (bgzip files),
// and any malformed stream is an error — there is no fallback.  On top of the serial decoder, GzParallel (end of this file)
// decodes ONE stream with several threads.  tests/test_host_io.py compares both with zlib on every compression level and
// strategy, stored / fixed / dynamic blocks, multi-member files, adversarial call patterns (1-byte outputs) and damaged streams.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdint>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <exception>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace mblhost {

// Worker threads must not let an exception (an allocation failure on a huge input) escape — that is std::terminate.  A pool keeps
// the first exception of every worker and hands it to the thread that joins.
class Workers {
public:
    template <class F>
    void spawn(F&& fn) {
        const size_t slot = err_.size();
        err_.emplace_back();
        th_.emplace_back([this, slot, fn]() mutable { try { fn(); } catch (...) { err_[slot] = std::current_exception(); } });
    }
    void reserve(size_t n) { err_.reserve(n); th_.reserve(n); }       // spawn() must not move err_ while workers run
    void join() {
        for (auto& t : th_) t.join();
        th_.clear();
        for (auto& e : err_) if (e) { std::exception_ptr x = e; err_.clear(); std::rethrow_exception(x); }
        err_.clear();
    }
    ~Workers() { for (auto& t : th_) if (t.joinable()) t.join(); }
private:
    std::vector<std::thread> th_;
    std::vector<std::exception_ptr> err_;
};

// growable array without value-initialisation (decoded symbols; a std::vector would zero-fill every growth)
template <class T>
class RawBuf {
public:
    RawBuf() = default;
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    ~RawBuf() { free(p_); }
    T* data() { return p_; }
    const T* data() const { return p_; }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    void clear() { n_ = 0; }
    void resize(size_t n) {                                       // keeps the first min(n, size()) elements; new ones are NOT initialised
        if (n > cap_) {
            const size_t c = std::max(n, cap_ + cap_ / 2);
            T* q = static_cast<T*>(realloc(p_, c * sizeof(T)));
            if (!q) throw std::bad_alloc();
            p_ = q; cap_ = c;
        }
        n_ = n;
    }
    T& operator[](size_t i) { return p_[i]; }
    const T& operator[](size_t i) const { return p_[i]; }
private:
    T* p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

class GzInflater {
public:
    GzInflater() { hist_.resize(kWindow); }
    GzInflater(const GzInflater&) = delete;
    GzInflater& operator=(const GzInflater&) = delete;
    ~GzInflater() { close(); }

    bool open(const std::string& path, std::string* err) {
        close();
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) { if (err) *err = "cannot open " + path; return false; }
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); if (err) *err = "cannot stat " + path; return false; }
        size_ = (size_t)st.st_size;
        if (size_) {
            map_ = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
            if (map_ == MAP_FAILED) { map_ = nullptr; ::close(fd); if (err) *err = "cannot map " + path; return false; }
            madvise(map_, size_, MADV_SEQUENTIAL);
        }
        ::close(fd);
        return attach(static_cast<const uint8_t*>(map_), size_);
    }
    // decode from memory the caller keeps alive (tests)
    bool attach(const uint8_t* data, size_t n) {
        in_ = data; end_ = data + n;
        bitbuf_ = 0; bitcnt_ = 0;
        state_ = kMemberHeader;
        hist_len_ = 0; pending_len_ = 0; stored_left_ = 0;
        failed_ = false; eof_ = false;
        error_ = nullptr;
        base_ = data;
        return true;
    }
    static bool looks_gzip(const std::string& path) {
        unsigned char m[2] = {0, 0};
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        const ssize_t r = ::read(fd, m, 2);
        ::close(fd);
        return r == 2 && m[0] == 0x1f && m[1] == 0x8b;
    }
    void close() {
        if (map_) munmap(map_, size_);
        map_ = nullptr; size_ = 0; in_ = end_ = nullptr;
    }
    bool eof() const { return eof_; }
    std::string error() const { return error_ ? std::string("gzip: ") + error_ : std::string(); }

    // Produces up to cap bytes at out; returns the number produced (less than cap only at the end of the input), or -1 on a
    // malformed stream (error() says why).
    long long read(char* out_c, size_t cap) {
        if (failed_) return -1;
        uint8_t* const out0 = reinterpret_cast<uint8_t*>(out_c);
        uint8_t* out = out0;
        uint8_t* const out_end = out0 + cap;
        while (out < out_end && !eof_) {
            switch (state_) {
                case kMemberHeader:
                    if (!skip_zero_padding()) { eof_ = true; break; }
                    if (!parse_header()) return fail_ret();
                    crc_ = crc32(0L, Z_NULL, 0); member_out_ = 0;
                    state_ = kBlockHeader;
                    break;
                case kBlockHeader:
                    if (!block_header()) return fail_ret();
                    break;
                case kStored: {
                    size_t n = std::min<size_t>(stored_left_, (size_t)(out_end - out));
                    if ((size_t)(end_ - in_) < n) { n = (size_t)(end_ - in_); if (!n) { set_err("truncated stored block"); return fail_ret(); } }
                    memcpy(out, in_, n);
                    account(out, n);
                    out += n; in_ += n; stored_left_ -= (uint32_t)n;
                    if (!stored_left_) state_ = final_ ? kTrailer : kBlockHeader;
                    break;
                }
                case kHuffman: {
                    uint8_t* const start = out;
                    const int r = huffman<uint8_t>(out0, out, out_end);
                    account(start, (size_t)(out - start));
                    if (r < 0) return fail_ret();
                    if (r == 1) state_ = final_ ? kTrailer : kBlockHeader;       // end of block
                    break;
                }
                case kTrailer:
                    if (!trailer()) return fail_ret();
                    state_ = kMemberHeader;
                    break;
            }
        }
        remember(out0, (size_t)(out - out0));
        return (long long)(out - out0);
    }

private:
    static constexpr size_t kWindow = 32768;
    static constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
    enum State { kMemberHeader, kBlockHeader, kStored, kHuffman, kTrailer };
    // table entry: op < 0x10: length / distance symbol with `op` extra bits, val = base; 0x10: literal val; 0x20: end of block;
    // 0x40 | b: pointer to a sub-table of 2^b entries at val; 0x80: invalid code
    struct Entry { uint16_t val; uint8_t bits; uint8_t op; };

    long long fail_ret() { failed_ = true; return -1; }
    void set_err(const char* m) { if (!error_) error_ = m; }

    void account(const uint8_t* p, size_t n) {
        if (!n) return;
        crc_ = crc32_big(crc_, p, n);
        member_out_ += n;
    }
    // history: the last 32 KiB of output BEFORE the buffer of the current read() call, so that a match can start in front of it;
    // updated once per call, after the call's output is complete (inside a call, distances are resolved against the buffer first)
    void remember(const uint8_t* p, size_t n) {
        if (!n) return;
        if (n >= kWindow) { memcpy(hist_.data(), p + n - kWindow, kWindow); hist_len_ = kWindow; return; }
        if (hist_len_ + n > kWindow) {
            const size_t drop = hist_len_ + n - kWindow;
            memmove(hist_.data(), hist_.data() + drop, hist_len_ - drop);
            hist_len_ -= drop;
        }
        memcpy(hist_.data() + hist_len_, p, n);
        hist_len_ += n;
    }
    static uLong crc32_big(uLong c, const uint8_t* p, size_t n) {
        while (n) { const uInt k = (uInt)std::min<size_t>(n, 1u << 30); c = crc32(c, p, k); p += k; n -= k; }
        return c;
    }

    // ---- bit reader -------------------------------------------------------------------------------------------------------
    inline void refill_fast() {                   // needs 8 readable bytes at in_
        uint64_t w;
        memcpy(&w, in_, 8);
        bitbuf_ |= w << bitcnt_;
        in_ += (63 - bitcnt_) >> 3;
        bitcnt_ |= 56;
    }
    inline void refill_safe() {
        while (bitcnt_ < 56 && in_ < end_) { bitbuf_ |= (uint64_t)*in_++ << bitcnt_; bitcnt_ += 8; }
    }
    inline bool need(int n) { if (bitcnt_ < n) { refill_safe(); if (bitcnt_ < n) { set_err("truncated stream"); return false; } } return true; }
    inline uint32_t take(int n) { const uint32_t v = (uint32_t)(bitbuf_ & ((1ull << n) - 1)); bitbuf_ >>= n; bitcnt_ -= n; return v; }
    void byte_align() {                            // drop the rest of the current byte, give whole unread bytes back to the input
        take(bitcnt_ & 7);
        in_ -= bitcnt_ >> 3;
        bitbuf_ = 0; bitcnt_ = 0;
    }

    // ---- gzip framing -----------------------------------------------------------------------------------------------------
    bool skip_zero_padding() {                      // some writers pad the end with zero bytes; false at the end of the input
        byte_align();
        while (in_ < end_ && *in_ == 0) ++in_;
        return in_ < end_;
    }
    bool parse_header() {
        if (end_ - in_ < 10 || in_[0] != 0x1f || in_[1] != 0x8b) { set_err("not a gzip member"); return false; }
        if (in_[2] != 8) { set_err("unknown compression method"); return false; }
        const int flg = in_[3];
        if (flg & 0xE0) { set_err("reserved flag bits set"); return false; }
        in_ += 10;
        if (flg & 4) {                              // FEXTRA
            if (end_ - in_ < 2) { set_err("truncated header"); return false; }
            const size_t xlen = in_[0] | ((size_t)in_[1] << 8);
            in_ += 2;
            if ((size_t)(end_ - in_) < xlen) { set_err("truncated header"); return false; }
            in_ += xlen;
        }
        for (int f : {8, 16})                       // FNAME, FCOMMENT: zero-terminated
            if (flg & f) {
                const void* z = memchr(in_, 0, (size_t)(end_ - in_));
                if (!z) { set_err("truncated header"); return false; }
                in_ = static_cast<const uint8_t*>(z) + 1;
            }
        if (flg & 2) { if (end_ - in_ < 2) { set_err("truncated header"); return false; } in_ += 2; }   // FHCRC
        return true;
    }
    bool trailer() {
        byte_align();
        if (end_ - in_ < 8) { set_err("truncated trailer"); return false; }
        uint32_t crc, isize;
        memcpy(&crc, in_, 4); memcpy(&isize, in_ + 4, 4);
        in_ += 8;
        if (crc != (uint32_t)crc_) { set_err("CRC mismatch"); return false; }
        if (isize != (uint32_t)member_out_) { set_err("length mismatch"); return false; }
        return true;
    }

    // ---- Huffman tables ---------------------------------------------------------------------------------------------------
    // Canonical code from lens[0..n): primary table of 2^tb entries at tab[0..), sub-tables appended behind it.  kind 0: literal /
    // length alphabet, 1: distance alphabet, 2: code-length alphabet.  Returns false on an over-subscribed or (for more than one
    // code) incomplete set of lengths.
    bool build(const uint8_t* lens, int n, int tb, int kind, std::vector<Entry>& tab) {
        static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        int count[16] = {0};
        for (int i = 0; i < n; ++i) ++count[lens[i]];
        count[0] = 0;
        int left = 1, used = 0;
        for (int l = 1; l <= 15; ++l) {
            left <<= 1;
            left -= count[l];
            if (left < 0) { set_err("over-subscribed Huffman code"); return false; }
            used += count[l];
        }
        if (left > 0 && used > 1) { set_err("incomplete Huffman code"); return false; }
        const Entry invalid{0, 1, 0x80};
        tab.assign((size_t)1 << tb, invalid);
        if (!used) return true;
        auto entry_of = [&](int sym, int bits) -> Entry {
            if (kind == 2) return Entry{(uint16_t)sym, (uint8_t)bits, 0x10};
            if (kind == 1) {
                if (sym >= 30) return Entry{0, (uint8_t)bits, 0x80};
                return Entry{dist_base[sym], (uint8_t)bits, dist_extra[sym]};
            }
            if (sym < 256) return Entry{(uint16_t)sym, (uint8_t)bits, 0x10};
            if (sym == 256) return Entry{0, (uint8_t)bits, 0x20};
            if (sym >= 286) return Entry{0, (uint8_t)bits, 0x80};
            return Entry{len_base[sym - 257], (uint8_t)bits, len_extra[sym - 257]};
        };
        auto reverse = [](uint32_t c, int l) { uint32_t r = 0; for (int i = 0; i < l; ++i) { r = (r << 1) | (c & 1); c >>= 1; } return r; };
        // first code of every length
        uint32_t next[16];
        {
            uint32_t code = 0;
            for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
            // count[0] was zeroed above, so next[1] = 0 as the standard demands
        }
        // pass 1: the longest code below every primary prefix that needs a sub-table
        std::vector<uint8_t>& sub_bits = sub_bits_;
        sub_bits.assign((size_t)1 << tb, 0);
        {
            uint32_t nx[16];
            memcpy(nx, next, sizeof nx);
            for (int sym = 0; sym < n; ++sym) {
                const int l = lens[sym];
                if (!l) continue;
                const uint32_t rev = reverse(nx[l]++, l);
                if (l > tb) { uint8_t& b = sub_bits[rev & (((uint32_t)1 << tb) - 1)]; b = std::max<uint8_t>(b, (uint8_t)(l - tb)); }
            }
        }
        for (size_t p = 0; p < sub_bits.size(); ++p)
            if (sub_bits[p]) {
                const size_t at = tab.size();
                if (at > 0xFFFF) { set_err("Huffman table too large"); return false; }
                tab[p] = Entry{(uint16_t)at, (uint8_t)tb, (uint8_t)(0x40 | sub_bits[p])};
                tab.resize(at + ((size_t)1 << sub_bits[p]), invalid);
            }
        // pass 2: fill
        for (int sym = 0; sym < n; ++sym) {
            const int l = lens[sym];
            if (!l) continue;
            const uint32_t rev = reverse(next[l]++, l);
            if (l <= tb) {
                const Entry e = entry_of(sym, l);
                for (uint32_t k = rev; k < ((uint32_t)1 << tb); k += (uint32_t)1 << l) tab[k] = e;
            } else {
                const uint32_t p = rev & (((uint32_t)1 << tb) - 1);
                const int sb = tab[p].op & 0x0F;
                const Entry e = entry_of(sym, l - tb);
                for (uint32_t k = rev >> tb; k < ((uint32_t)1 << sb); k += (uint32_t)1 << (l - tb)) tab[tab[p].val + k] = e;
            }
        }
        return true;
    }

    bool block_header() {
        if (!need(3)) return false;
        final_ = take(1) != 0;
        const uint32_t type = take(2);
        if (type == 0) {
            byte_align();
            if (end_ - in_ < 4) { set_err("truncated stored block"); return false; }
            const uint32_t len = in_[0] | ((uint32_t)in_[1] << 8), nlen = in_[2] | ((uint32_t)in_[3] << 8);
            if ((len ^ nlen) != 0xFFFF) { set_err("stored block length check failed"); return false; }
            in_ += 4;
            stored_left_ = len;
            state_ = len ? kStored : (final_ ? kTrailer : kBlockHeader);
            return true;
        }
        if (type == 1) {
            if (fixed_lit_.empty()) {
                uint8_t l[288];
                for (int i = 0; i < 144; ++i) l[i] = 8;
                for (int i = 144; i < 256; ++i) l[i] = 9;
                for (int i = 256; i < 280; ++i) l[i] = 7;
                for (int i = 280; i < 288; ++i) l[i] = 8;
                uint8_t d[32];
                for (int i = 0; i < 32; ++i) d[i] = 5;
                if (!build(l, 288, kLitBits, 0, fixed_lit_) || !build(d, 32, kDistBits, 1, fixed_dist_)) return false;
            }
            lit_ = fixed_lit_.data(); dist_ = fixed_dist_.data();
            state_ = kHuffman;
            return true;
        }
        if (type == 3) { set_err("reserved block type"); return false; }
        if (!need(14)) return false;
        const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
        if (hlit > 286 || hdist > 30) { set_err("too many length or distance codes"); return false; }
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t pl[19] = {0};
        for (int i = 0; i < hclen; ++i) { if (!need(3)) return false; pl[order[i]] = (uint8_t)take(3); }
        if (!build(pl, 19, kPreBits, 2, pre_)) return false;
        uint8_t lens[286 + 30 + 138];
        int i = 0;
        while (i < hlit + hdist) {
            refill_safe();
            if (bitcnt_ < 1) { set_err("truncated stream"); return false; }
            const Entry e = pre_[bitbuf_ & ((1u << kPreBits) - 1)];
            if (e.op != 0x10 || e.bits > bitcnt_) { set_err("invalid code-length code"); return false; }
            take(e.bits);
            const int sym = e.val;
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            int rep;
            uint8_t v = 0;
            if (sym == 16) {
                if (!i) { set_err("repeat without a previous length"); return false; }
                if (!need(2)) return false;
                v = lens[i - 1]; rep = 3 + (int)take(2);
            } else if (sym == 17) {
                if (!need(3)) return false;
                rep = 3 + (int)take(3);
            } else {
                if (!need(7)) return false;
                rep = 11 + (int)take(7);
            }
            if (i + rep > hlit + hdist) { set_err("code lengths overrun"); return false; }
            memset(lens + i, v, (size_t)rep);
            i += rep;
        }
        if (!lens[256]) { set_err("no end-of-block code"); return false; }
        if (!build(lens, hlit, kLitBits, 0, dyn_lit_) || !build(lens + hlit, hdist, kDistBits, 1, dyn_dist_)) return false;
        lit_ = dyn_lit_.data(); dist_ = dyn_dist_.data();
        state_ = kHuffman;
        return true;
    }

    // copies len symbols from dist back; the source may lie before this call's buffer: bytes come from the history, and in the
    // 16-bit mode of the parallel decoder (the 32 KiB in front of the segment are not known yet) they become markers
    // 256 + position in that window, resolved once the previous segment is complete
    template <class Sym>
    inline bool copy_match(Sym* const out0, Sym*& out, uint32_t len, uint32_t dist) {
        const size_t have = (size_t)(out - out0);
        if (dist > have) {
            size_t back = dist - have;                            // symbols the source starts before out0
            if (sizeof(Sym) == 2) {
                if (back > kWindow) { set_err("distance reaches before the window"); return false; }
                while (len && back) { *out++ = (Sym)(256 + (kWindow - back)); --back; --len; }
            } else {
                if (back > hist_len_) { set_err("distance reaches before the start of the data"); return false; }
                const uint8_t* h = hist_.data() + hist_len_ - back;
                while (len && h < hist_.data() + hist_len_) { *out++ = (Sym)*h++; --len; }
            }
            // whatever is left continues at out0 (dist symbols behind out again)
        }
        const Sym* src = out - dist;
        while (len--) *out++ = *src++;
        return true;
    }

    // Decodes symbols of the current block into [out, out_end).  Returns 1 at the end of the block, 0 when the output is full,
    // -1 on error.
    template <class Sym>
    int huffman(Sym* const out0, Sym*& out, Sym* const out_end) {
        constexpr uint32_t kWord = 8 / sizeof(Sym);                // symbols per 8-byte copy
        const Entry* const lit = lit_;
        const Entry* const dst = dist_;
        const uint32_t lit_mask = (1u << kLitBits) - 1, dist_mask = (1u << kDistBits) - 1;
        if (pending_len_) {                                       // a match the previous call could not finish
            const uint32_t n = (uint32_t)std::min<size_t>(pending_len_, (size_t)(out_end - out));
            if (!copy_match<Sym>(out0, out, n, pending_dist_)) return -1;
            pending_len_ -= n;
            if (pending_len_) return 0;
        }
        // fast loop: room for three literals and a whole match (plus the copy's overshoot), sixteen readable input bytes
        while (out_end - out >= 3 + 258 + 16 && end_ - in_ >= 16) {
            refill_fast();
            Entry e = lit[bitbuf_ & lit_mask];
            if (e.op == 0x10) {                                   // up to three literals per refill (3 x 15 bits)
                bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                *out++ = (Sym)e.val;
                e = lit[bitbuf_ & lit_mask];
                if (e.op == 0x10) {
                    bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                    *out++ = (Sym)e.val;
                    e = lit[bitbuf_ & lit_mask];
                    if (e.op == 0x10) {
                        bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                        *out++ = (Sym)e.val;
                        continue;
                    }
                }
                refill_fast();
            }
            if (e.op & 0x40) {                                    // code longer than the primary table
                bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                e = lit[e.val + (bitbuf_ & ((1u << (e.op & 0x0F)) - 1))];
                if (e.op == 0x10) { bitbuf_ >>= e.bits; bitcnt_ -= e.bits; *out++ = (Sym)e.val; continue; }
            }
            bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
            if (e.op >= 0x10) {
                if (e.op == 0x20) return 1;
                set_err("invalid literal/length code");
                return -1;
            }
            const uint32_t len = e.val + (uint32_t)(bitbuf_ & ((1u << e.op) - 1));
            bitbuf_ >>= e.op; bitcnt_ -= e.op;
            // at most 15 + 5 bits used since the last refill (>= 56): 36 left, the distance code needs up to 15 + 13
            Entry d = dst[bitbuf_ & dist_mask];
            if (d.op & 0x40) {
                bitbuf_ >>= d.bits; bitcnt_ -= d.bits;
                d = dst[d.val + (bitbuf_ & ((1u << (d.op & 0x0F)) - 1))];
            }
            if (d.op >= 0x10) { set_err("invalid distance code"); return -1; }
            bitbuf_ >>= d.bits; bitcnt_ -= d.bits;
            const uint32_t dist = d.val + (uint32_t)(bitbuf_ & ((1u << d.op) - 1));
            bitbuf_ >>= d.op; bitcnt_ -= d.op;
            if (dist > (size_t)(out - out0)) {
                if (!copy_match<Sym>(out0, out, len, dist)) return -1;
            } else if (dist >= kWord) {                            // eight bytes at a time; writes up to 2 words past the match
                const Sym* s = out - dist;
                Sym* o = out;
                Sym* const oe = out + len;
                uint64_t w;
                memcpy(&w, s, 8); memcpy(o, &w, 8);
                memcpy(&w, s + kWord, 8); memcpy(o + kWord, &w, 8);    // most matches in sequence data are shorter than 16
                if (len > 2 * kWord) {
                    s += 2 * kWord; o += 2 * kWord;
                    do { memcpy(&w, s, 8); memcpy(o, &w, 8); s += kWord; o += kWord; } while (o < oe);
                }
                out = oe;
            } else {
                const Sym* s = out - dist;
                Sym* const oe = out + len;
                while (out < oe) *out++ = *s++;
            }
        }
        // careful loop: one symbol at a time, every bound checked
        while (out < out_end) {
            refill_safe();
            if (bitcnt_ < 1) { set_err("truncated stream"); return -1; }
            Entry e = lit[bitbuf_ & lit_mask];
            int used = 0;
            if (e.op & 0x40) {
                used = e.bits;
                e = lit[e.val + ((bitbuf_ >> used) & ((1u << (e.op & 0x0F)) - 1))];
            }
            if (e.op & 0x80) { set_err("invalid literal/length code"); return -1; }
            if (used + e.bits > bitcnt_) { set_err("truncated stream"); return -1; }
            take(used + e.bits);
            if (e.op == 0x10) { *out++ = (Sym)e.val; continue; }
            if (e.op == 0x20) return 1;
            if (!need(e.op)) return -1;
            const uint32_t len = e.val + take(e.op);
            refill_safe();
            if (bitcnt_ < 1) { set_err("truncated stream"); return -1; }
            Entry d = dst[bitbuf_ & dist_mask];
            used = 0;
            if (d.op & 0x40) {
                used = d.bits;
                d = dst[d.val + ((bitbuf_ >> used) & ((1u << (d.op & 0x0F)) - 1))];
            }
            if (d.op >= 0x10) { set_err("invalid distance code"); return -1; }
            if (used + d.bits > bitcnt_) { set_err("truncated stream"); return -1; }
            take(used + d.bits);
            if (!need(d.op)) return -1;
            const uint32_t dist = d.val + take(d.op);
            const uint32_t n = (uint32_t)std::min<size_t>(len, (size_t)(out_end - out));
            if (!copy_match<Sym>(out0, out, n, dist)) return -1;
            if (n < len) { pending_len_ = len - n; pending_dist_ = dist; return 0; }
        }
        return 0;
    }


    // ---- block-granular interface of the parallel reader (GzParallel) ----------------------------------------------------------
    uint64_t tell_bits() const { return (uint64_t)(in_ - base_) * 8 - (uint64_t)bitcnt_; }
    bool seek_bits(uint64_t pos) {
        in_ = base_ + (pos >> 3);
        bitbuf_ = 0; bitcnt_ = 0;
        pending_len_ = 0; stored_left_ = 0;
        state_ = kBlockHeader;
        const int skip = (int)(pos & 7);
        if (in_ > end_) return false;
        if (skip) { if (!need(skip)) return false; take(skip); }
        return true;
    }
    enum Stop { kAtStop = 0, kPastEnd = 1, kFinalBlock = 2, kFailed = 3 };
    // Decodes whole blocks from the current block boundary, appending symbols to out, until the position equals one of the
    // sorted stop positions (checked at block boundaries only), reaches end_bit, or the member's last block is done.
    template <class Sym>
    Stop run_blocks(RawBuf<Sym>& out, const uint64_t* stops, size_t n_stops, uint64_t end_bit, size_t* stop_index) {
        size_t used = out.size();
        size_t next_stop = 0;
        for (bool first = true;; first = false) {
            const uint64_t pos = tell_bits();
            while (next_stop < n_stops && stops[next_stop] < pos) ++next_stop;
            if (!first && next_stop < n_stops && stops[next_stop] == pos) { out.resize(used); *stop_index = next_stop; return kAtStop; }
            if (!first && pos >= end_bit) { out.resize(used); return kPastEnd; }
            if (!block_header()) { out.resize(used); return kFailed; }
            if (state_ == kStored) {
                if ((size_t)(end_ - in_) < stored_left_) { set_err("truncated stored block"); out.resize(used); return kFailed; }
                out.resize(used + stored_left_);
                for (uint32_t i = 0; i < stored_left_; ++i) out[used + i] = (Sym)in_[i];
                used += stored_left_; in_ += stored_left_; stored_left_ = 0;
            } else if (state_ == kHuffman) {
                for (;;) {
                    if (out.size() < used + (1u << 20)) out.resize(used + (4u << 20));
                    Sym* o = out.data() + used;
                    const int r = huffman<Sym>(out.data(), o, out.data() + out.size());
                    used = (size_t)(o - out.data());
                    if (r < 0) { out.resize(used); return kFailed; }
                    if (r == 1) break;
                }
            }
            state_ = kBlockHeader;
            if (final_) { out.resize(used); return kFinalBlock; }
        }
    }
    // Is there a non-final dynamic block at bit position pos whose first symbols decode to text?  (candidate test of the block
    // search; a false positive is caught later because the previous segment will not end exactly there)
    bool probe(uint64_t pos) {
        const uint8_t* p = base_ + (pos >> 3);
        if (end_ - p < 64) return false;
        uint64_t w;
        memcpy(&w, p, 8);
        w >>= (pos & 7);
        if ((w & 7) != 4) return false;                           // BFINAL 0, BTYPE 2 (dynamic)
        const uint32_t hlit = (uint32_t)((w >> 3) & 31) + 257, hdist = (uint32_t)((w >> 8) & 31) + 1;
        if (hlit > 286 || hdist > 30) return false;
        error_ = nullptr;
        if (!seek_bits(pos) || !block_header() || state_ != kHuffman) { error_ = nullptr; return false; }
        uint16_t buf[1024 + 300];
        uint16_t* o = buf;
        const int r = huffman<uint16_t>(buf, o, buf + 1024);
        if (r < 0) { error_ = nullptr; return false; }
        if (o - buf < 64 && r != 1) return false;
        for (const uint16_t* q = buf; q < o; ++q) {
            const uint16_t c = *q;
            if (c >= 256) continue;                               // from the unknown window
            if (!((c >= 32 && c < 127) || c == '\n' || c == '\r' || c == '\t')) return false;
        }
        return true;
    }

    void* map_ = nullptr;
    size_t size_ = 0;
    const uint8_t *in_ = nullptr, *end_ = nullptr, *base_ = nullptr;
    uint64_t bitbuf_ = 0;
    int bitcnt_ = 0;
    State state_ = kMemberHeader;
    bool final_ = false, failed_ = false, eof_ = false;
    uint32_t stored_left_ = 0, pending_len_ = 0, pending_dist_ = 0;
    std::vector<Entry> fixed_lit_, fixed_dist_, dyn_lit_, dyn_dist_, pre_;
    std::vector<uint8_t> sub_bits_;
    const Entry *lit_ = nullptr, *dist_ = nullptr;
    std::vector<uint8_t> hist_;
    size_t hist_len_ = 0;
    uLong crc_ = 0;
    uint64_t member_out_ = 0;
    const char* error_ = nullptr;
    friend class GzParallel;
};


// Parallel decoding of ONE gzip stream (the idea of pugz, Kerbiriou & Chikhi 2019, for text): a group of the compressed file is
// cut into segments; a dynamic block start is searched near every cut (bit by bit: header fields, complete Huffman codes, text-
// like first symbols); every segment is decoded from its block start by its own thread — the first one continues the stream with
// the known history, the others do not know the 32 KiB in front of them and emit 16-bit symbols in which a reference into that
// window is a marker; a segment ends exactly where the next one starts (a searched position that no segment ends on was a false
// positive and its segment is dropped; the one before it simply runs on).  Then the windows are resolved in order (32 KiB per
// segment), the bulk in parallel, CRC-32s are combined, and the serial decoder's state is moved to the end of the group.  Anything
// unusual inside a group (member end, stored or fixed blocks where a cut falls, no block start found) only means fewer segments;
// the serial decoder handles the rest.  The output is identical by construction and the member's CRC-32 still has to match.
class GzParallel {
public:
    bool open(const std::string& path, std::string* err) { return main_.open(path, err); }
    bool attach(const uint8_t* data, size_t n) { return main_.attach(data, n); }
    bool eof() const { return main_.eof(); }
    std::string error() const { return main_.error(); }
    // groups decoded in parallel / segments that took part in them / block starts that turned out to be false positives
    size_t groups() const { return groups_; }
    size_t segments() const { return segments_; }
    size_t false_starts() const { return false_starts_; }
    size_t bgzf_groups() const { return bgzf_groups_; }

    // Appends decoded bytes to dst[at..) (dst is grown as needed): whole blocks, about `want` bytes when the stream allows, at least
    // one byte unless the input is at its end.  Returns the number of bytes, 0 at the end of the input, -1 on a malformed stream.
    // Between calls the stream always stands at a block or member boundary.
    template <class Buf>
    long long read_some(Buf& dst, size_t at, size_t want, unsigned threads) {
        const size_t seg_bytes = (size_t)2 << 20;                 // compressed bytes per segment
        const unsigned T = std::max(1u, threads);
        GzInflater& m = main_;
        for (;;) {
            if (m.failed_) return -1;
            if (m.eof_) return 0;
            if (m.state_ == GzInflater::kMemberHeader) {
                if (!m.skip_zero_padding()) { m.eof_ = true; return 0; }
                if (T >= 2) {                                     // BGZF (bgzip): members carry their own size, decode a run of them side by side
                    const long long r = bgzf_group(dst, at, want, T);
                    if (r == -2) continue;                        // only empty blocks (the BGZF end marker) were consumed
                    if (r != 0) return r;
                }
                if (!m.parse_header()) { m.failed_ = true; return -1; }
                m.crc_ = crc32(0L, Z_NULL, 0); m.member_out_ = 0; m.hist_len_ = 0;
                m.state_ = GzInflater::kBlockHeader;
            } else if (m.state_ == GzInflater::kTrailer) {
                if (!m.trailer()) { m.failed_ = true; return -1; }
                small_members_ = m.member_out_ < ((uint64_t)8 << 20);   // bgzip-style files: members of 64 KiB, nothing to split
                m.state_ = GzInflater::kMemberHeader;
                continue;
            }
            const uint64_t c0 = m.tell_bits();
            const size_t left = (size_t)(m.end_ - m.base_) - (size_t)(c0 >> 3);
            const size_t n_seg = std::min<size_t>(T, left / seg_bytes);
            if (n_seg >= 2 && !small_members_) {
                const long long r = group(dst, at, c0, n_seg, seg_bytes);
                if (r < 0) return r;
                if (r > 0) return r;
                if (m.state_ != GzInflater::kBlockHeader) continue;   // an empty last block ended the member
            }
            // serial: the tail of the file, one thread, or a place where no second block start was found
            serial_.clear();
            size_t dummy = 0;
            const uint64_t end_bit = c0 + (uint64_t)std::max<size_t>(want / 4, 1u << 16) * 8;
            const GzInflater::Stop how = m.run_blocks<uint8_t>(serial_, nullptr, 0, end_bit, &dummy);
            if (how == GzInflater::kFailed) { m.failed_ = true; return -1; }
            if (how == GzInflater::kFinalBlock) m.state_ = GzInflater::kTrailer;
            if (serial_.empty()) continue;                        // blocks without output (sync flushes)
            if (dst.size() < at + serial_.size()) dst.resize(at + serial_.size());
            memcpy(dst.data() + at, serial_.data(), serial_.size());
            m.account(serial_.data(), serial_.size());
            m.remember(serial_.data(), serial_.size());
            return (long long)serial_.size();
        }
    }

private:
    // size of the BGZF block at p (RFC 1952 extra subfield 'B','C': BSIZE = block size - 1), 0 when p is not a BGZF header
    static size_t bgzf_block_size(const uint8_t* p, const uint8_t* end) {
        if (end - p < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
        const size_t xlen = p[10] | ((size_t)p[11] << 8);
        if ((size_t)(end - p) < 12 + xlen) return 0;
        for (size_t i = 0; i + 4 <= xlen;) {
            const uint8_t* f = p + 12 + i;
            const size_t slen = f[2] | ((size_t)f[3] << 8);
            if (f[0] == 'B' && f[1] == 'C' && slen == 2 && i + 6 <= xlen) return (size_t)(f[4] | ((size_t)f[5] << 8)) + 1;
            i += 4 + slen;
        }
        return 0;
    }
    // A run of BGZF blocks (up to ~want bytes of output) decoded by T threads, every block with its own CRC check, outputs
    // appended in order.  0 when the stream is not BGZF here, -2 when blocks without output were consumed, -1 on error.
    template <class Buf>
    long long bgzf_group(Buf& dst, size_t at, size_t want, unsigned T) {
        GzInflater& m = main_;
        std::vector<std::pair<const uint8_t*, size_t>> blocks;
        const uint8_t* p = m.in_;
        size_t comp = 0;
        while (p < m.end_ && comp < std::max<size_t>(want / 3, 1u << 20)) {
            const size_t bs = bgzf_block_size(p, m.end_);
            if (!bs || (size_t)(m.end_ - p) < bs) break;
            blocks.emplace_back(p, bs);
            p += bs; comp += bs;
        }
        if (blocks.size() < 2) return 0;
        T = (unsigned)std::min<size_t>(T, blocks.size());
        std::vector<RawBuf<uint8_t>> part(T);
        std::vector<const char*> perr(T, nullptr);
        Workers th;
        th.reserve(T);
        auto work = [&](unsigned t) {
            GzInflater d;
            RawBuf<uint8_t>& o = part[t];
            size_t used = 0;
            for (size_t b = blocks.size() * t / T; b < blocks.size() * (t + 1) / T; ++b) {
                d.attach(blocks[b].first, blocks[b].second);
                for (;;) {                                         // a BGZF block holds at most 64 KiB
                    o.resize(used + (1u << 17));
                    const long long r = d.read(reinterpret_cast<char*>(o.data() + used), 1u << 17);
                    if (r < 0) { perr[t] = d.error_ ? d.error_ : "bad block"; o.resize(used); return; }
                    used += (size_t)r;
                    if (r < (1 << 17)) break;
                }
            }
            o.resize(used);
        };
        for (unsigned t = 1; t < T; ++t) th.spawn([&work, t] { work(t); });
        work(0);
        th.join();
        for (unsigned t = 0; t < T; ++t) if (perr[t]) { m.set_err(perr[t]); m.failed_ = true; return -1; }
        size_t n_out = 0;
        for (auto& o : part) n_out += o.size();
        m.in_ = p;                                                // behind the run, at the next member header (or the end)
        m.bitbuf_ = 0; m.bitcnt_ = 0;
        if (!n_out) return -2;                                    // only empty blocks (the BGZF end marker): progress without output
        if (dst.size() < at + n_out) dst.resize(at + n_out);
        size_t o = at;
        for (auto& x : part) { if (x.size()) memcpy(dst.data() + o, x.data(), x.size()); o += x.size(); }
        ++bgzf_groups_;
        return (long long)n_out;
    }

    struct Seg {
        uint64_t start = 0, end = 0;
        RawBuf<uint16_t> sym;                                     // segments 1..: symbols with markers
        RawBuf<uint8_t> bytes;                                    // segment 0: plain bytes
        GzInflater::Stop how = GzInflater::kFailed;
        size_t stop_index = 0;
        bool found = false;
    };

    template <class Buf>
    long long group(Buf& dst, size_t at, uint64_t c0, size_t n_seg, size_t seg_bytes) {
        const uint8_t* base = main_.base_;
        const size_t total = (size_t)(main_.end_ - base);
        if (seg_.size() < n_seg) { seg_.clear(); for (size_t j = 0; j < n_seg; ++j) seg_.emplace_back(new Seg()); }   // buffers are reused
        struct SegView { std::vector<std::unique_ptr<Seg>>& v; Seg& operator[](size_t j) { return *v[j]; } } seg{seg_};
        for (size_t j = 0; j < n_seg; ++j) { Seg& x = seg[j]; x.start = x.end = 0; x.how = GzInflater::kFailed; x.stop_index = 0; x.found = false; x.sym.clear(); x.bytes.clear(); }
        if (dec_.size() < n_seg) dec_.resize(n_seg);
        for (auto& d : dec_) if (!d) d.reset(new GzInflater());
        seg[0].start = c0; seg[0].found = true;
        // 1. block starts near the cuts
        {
            Workers th;
            th.reserve(n_seg);
            for (size_t j = 1; j < n_seg; ++j) th.spawn([&, j] {
                GzInflater& d = *dec_[j];
                d.attach(base, total);
                const uint64_t from = ((c0 >> 3) + j * seg_bytes) * 8, to = from + (uint64_t)seg_bytes * 4;   // half a segment
                for (uint64_t pos = from; pos < to; ++pos)
                    if (d.probe(pos)) { seg[j].start = pos; seg[j].found = true; break; }
            });
            th.join();
        }
        std::vector<uint64_t> stops;
        for (size_t j = 1; j < n_seg; ++j) if (seg[j].found) stops.push_back(seg[j].start);
        if (stops.empty()) return 0;
        const uint64_t end_bit = ((c0 >> 3) + n_seg * seg_bytes) * 8;
        // 2. decode the segments
        {
            Workers th;
            th.reserve(n_seg);
            for (size_t j = 1; j < n_seg; ++j) if (seg[j].found) th.spawn([&, j] {
                GzInflater& d = *dec_[j];
                d.error_ = nullptr;
                if (!d.seek_bits(seg[j].start)) { seg[j].how = GzInflater::kFailed; return; }
                seg[j].sym.clear();
                seg[j].how = d.run_blocks<uint16_t>(seg[j].sym, stops.data(), stops.size(), end_bit, &seg[j].stop_index);
                seg[j].end = d.tell_bits();
            });
            seg[0].bytes.clear();
            seg[0].how = main_.run_blocks<uint8_t>(seg[0].bytes, stops.data(), stops.size(), end_bit, &seg[0].stop_index);
            seg[0].end = main_.tell_bits();
            th.join();
        }
        if (seg[0].how == GzInflater::kFailed) { main_.failed_ = true; return -1; }
        // 3. the chain of segments that really follow each other
        std::vector<size_t> chain{0};
        for (size_t cur = 0; seg[cur].how == GzInflater::kAtStop;) {
            const uint64_t pos = stops[seg[cur].stop_index];
            size_t nxt = 0;
            for (size_t j = 1; j < n_seg; ++j) if (seg[j].found && seg[j].start == pos) nxt = j;
            if (!nxt || seg[nxt].how == GzInflater::kFailed) {
                // the segment behind this stop is unusable (cannot happen for a true block start of a well-formed stream): go on serially
                break;
            }
            chain.push_back(nxt);
            cur = nxt;
        }
        Seg& last = seg[chain.back()];
        // 4. sizes, windows, bytes
        size_t n_out = seg[0].bytes.size();
        for (size_t k = 1; k < chain.size(); ++k) n_out += seg[chain[k]].sym.size();
        if (dst.size() < at + n_out) dst.resize(at + n_out);
        uint8_t* out = reinterpret_cast<uint8_t*>(dst.data()) + at;
        if (!seg[0].bytes.empty()) memcpy(out, seg[0].bytes.data(), seg[0].bytes.size());
        // windows in order: the last 32 KiB in front of every 16-bit segment (history of the stream + what the group has produced)
        std::vector<std::vector<uint8_t>> window(chain.size());
        std::vector<size_t> off(chain.size() + 1, 0);
        off[1] = seg[0].bytes.size();
        for (size_t k = 1; k < chain.size(); ++k) off[k + 1] = off[k] + seg[chain[k]].sym.size();
        bool bad_marker = false;
        std::vector<size_t> invalid_below(chain.size(), 0);       // window positions below this hold no data (start of a member)
        for (size_t k = 1; k < chain.size(); ++k) {
            std::vector<uint8_t>& w = window[k];
            w.assign(GzInflater::kWindow, 0);
            // bytes in front of segment k: out[0, off[k]) preceded by the stream's history
            const size_t have = off[k];
            size_t valid;                                          // how many of the window's last positions hold real data
            if (have >= GzInflater::kWindow) { memcpy(w.data(), out + have - GzInflater::kWindow, GzInflater::kWindow); valid = GzInflater::kWindow; }
            else {
                const size_t from_hist = std::min(main_.hist_len_, GzInflater::kWindow - have);
                memcpy(w.data() + GzInflater::kWindow - have - from_hist, main_.hist_.data() + main_.hist_len_ - from_hist, from_hist);
                memcpy(w.data() + GzInflater::kWindow - have, out, have);
                valid = have + from_hist;
            }
            // resolve the tail of segment k now (the next window needs it); the bulk follows in parallel
            const RawBuf<uint16_t>& sy = seg[chain[k]].sym;
            const size_t n = sy.size(), tail = std::min<size_t>(n, GzInflater::kWindow);
            for (size_t i = n - tail; i < n; ++i) {
                const uint16_t c = sy[i];
                if (c >= 256 && (size_t)(c - 256) < GzInflater::kWindow - valid) bad_marker = true;
                out[off[k] + i] = c < 256 ? (uint8_t)c : w[c - 256];
            }
            invalid_below[k] = GzInflater::kWindow - valid;
        }
        std::vector<char> bad_bulk(chain.size(), 0);
        std::vector<uLong> part_crc(chain.size(), 0);
        {
            Workers th;
            th.reserve(chain.size());
            for (size_t k = 1; k < chain.size(); ++k) th.spawn([&, k] {
                const RawBuf<uint16_t>& sy = seg[chain[k]].sym;
                const uint8_t* w = window[k].data();
                const size_t n = sy.size(), bulk = n - std::min<size_t>(n, GzInflater::kWindow);
                uint8_t* o = out + off[k];
                const size_t lo = invalid_below[k];
                bool bad = false;
                for (size_t i = 0; i < bulk; ++i) {
                    const uint16_t c = sy[i];
                    if (c >= 256 && (size_t)(c - 256) < lo) bad = true;
                    o[i] = c < 256 ? (uint8_t)c : w[c - 256];
                }
                bad_bulk[k] = bad;
                part_crc[k] = GzInflater::crc32_big(crc32(0L, Z_NULL, 0), o, n);      // the tail was resolved above
            });
            part_crc[0] = GzInflater::crc32_big(crc32(0L, Z_NULL, 0), out, off[1]);
            th.join();
        }
        for (char b : bad_bulk) bad_marker |= b != 0;
        if (bad_marker) { main_.set_err("distance reaches before the start of the data"); main_.failed_ = true; return -1; }
        // 5. the serial decoder continues behind the group
        if (chain.size() > 1) {
            if (!main_.seek_bits(last.end)) { main_.failed_ = true; return -1; }
        }
        if (last.how == GzInflater::kFinalBlock) main_.state_ = GzInflater::kTrailer;
        else if (last.how == GzInflater::kFailed) { main_.failed_ = true; main_.error_ = dec_[chain.back()]->error_; return -1; }
        for (size_t k = 0; k < chain.size(); ++k)                 // CRC-32 of the group from the segments' own CRCs
            if (off[k + 1] > off[k]) main_.crc_ = crc32_combine(main_.crc_, part_crc[k], (z_off_t)(off[k + 1] - off[k]));
        main_.member_out_ += n_out;
        main_.remember(out, n_out);
        if (chain.size() > 1) {
            ++groups_; segments_ += chain.size();
            for (size_t j = 1; j <= chain.back(); ++j) if (seg[j].found && std::find(chain.begin(), chain.end(), j) == chain.end()) ++false_starts_;
        }
        return (long long)n_out;
    }

    GzInflater main_;
    std::vector<std::unique_ptr<GzInflater>> dec_;
    RawBuf<uint8_t> serial_;
    std::vector<std::unique_ptr<Seg>> seg_;
    size_t groups_ = 0, segments_ = 0, false_starts_ = 0, bgzf_groups_ = 0;
    bool small_members_ = false;
};

}  // namespace mblhost
