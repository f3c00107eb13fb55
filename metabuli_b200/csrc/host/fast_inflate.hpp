// gzip (RFC 1952) / DEFLATE (RFC 1951) decoder for the streaming FASTA/FASTQ reader.
//
// Why not zlib: reads arrive gzip-compressed almost always, inflate is sequential per file, and zlib's inflate is what bounds the
// C++ host end to end on such input (~200 MB/s of text on the development machine = 1.1 M reads/s).  This decoder is
// written for that one job: the compressed file is mapped (no input refill logic), the bit buffer is 64 bits wide and refilled
// with one unaligned 8-byte load, literals/lengths are decoded from an 11-bit primary table (sub-tables for longer codes), up to
// three literals are emitted per refill, and matches are copied eight bytes at a time.  Output is produced into the caller's
// buffer in exact amounts (a match that does not fit is resumed by the next call), with a 32 KiB history kept across calls.
// Every member's CRC-32 and length are verified (zlib's crc32() is used for that), members may be concatenated (bgzip files),
// and any malformed stream is an error — there is no fallback.  tests/test_host_io.py compares it with zlib on every compression
// level and strategy, stored / fixed / dynamic blocks, multi-member files and adversarial call patterns (1-byte outputs).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace mblhost {

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
        error_.clear();
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
    const std::string& error() const { return error_; }

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
                    const int r = huffman(out0, out, out_end);
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
    void set_err(const char* m) { if (error_.empty()) error_ = std::string("gzip: ") + m; }

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

    // copies len bytes from dist back; the source may lie before this call's buffer (history)
    inline bool copy_match(uint8_t* const out0, uint8_t*& out, uint32_t len, uint32_t dist) {
        const size_t have = (size_t)(out - out0);
        if (dist > have) {
            const size_t back = dist - have;                      // bytes the source starts before out0
            if (back > hist_len_) { set_err("distance reaches before the start of the data"); return false; }
            const uint8_t* h = hist_.data() + hist_len_ - back;
            while (len && h < hist_.data() + hist_len_) { *out++ = *h++; --len; }
            // whatever is left continues at out0 (dist bytes behind out again)
        }
        const uint8_t* src = out - dist;
        while (len--) *out++ = *src++;
        return true;
    }

    // Decodes symbols of the current block into [out, out_end).  Returns 1 at the end of the block, 0 when the output is full,
    // -1 on error.
    int huffman(uint8_t* const out0, uint8_t*& out, uint8_t* const out_end) {
        const Entry* const lit = lit_;
        const Entry* const dst = dist_;
        const uint32_t lit_mask = (1u << kLitBits) - 1, dist_mask = (1u << kDistBits) - 1;
        if (pending_len_) {                                       // a match the previous call could not finish
            const uint32_t n = (uint32_t)std::min<size_t>(pending_len_, (size_t)(out_end - out));
            if (!copy_match(out0, out, n, pending_dist_)) return -1;
            pending_len_ -= n;
            if (pending_len_) return 0;
        }
        // fast loop: room for three literals and a whole match (plus the copy's overshoot), sixteen readable input bytes
        while (out_end - out >= 3 + 258 + 16 && end_ - in_ >= 16) {
            refill_fast();
            Entry e = lit[bitbuf_ & lit_mask];
            if (e.op == 0x10) {                                   // up to three literals per refill (3 x 15 bits)
                bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                *out++ = (uint8_t)e.val;
                e = lit[bitbuf_ & lit_mask];
                if (e.op == 0x10) {
                    bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                    *out++ = (uint8_t)e.val;
                    e = lit[bitbuf_ & lit_mask];
                    if (e.op == 0x10) {
                        bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                        *out++ = (uint8_t)e.val;
                        continue;
                    }
                }
                refill_fast();
            }
            if (e.op & 0x40) {                                    // code longer than the primary table
                bitbuf_ >>= e.bits; bitcnt_ -= e.bits;
                e = lit[e.val + (bitbuf_ & ((1u << (e.op & 0x0F)) - 1))];
                if (e.op == 0x10) { bitbuf_ >>= e.bits; bitcnt_ -= e.bits; *out++ = (uint8_t)e.val; continue; }
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
                if (!copy_match(out0, out, len, dist)) return -1;
            } else if (dist >= 8) {                                // eight bytes at a time; writes up to 15 bytes past the match
                const uint8_t* s = out - dist;
                uint8_t* o = out;
                uint8_t* const oe = out + len;
                uint64_t w;
                memcpy(&w, s, 8); memcpy(o, &w, 8);
                memcpy(&w, s + 8, 8); memcpy(o + 8, &w, 8);       // most matches in sequence data are shorter than 16
                if (len > 16) {
                    s += 16; o += 16;
                    do { memcpy(&w, s, 8); memcpy(o, &w, 8); s += 8; o += 8; } while (o < oe);
                }
                out = oe;
            } else {
                const uint8_t* s = out - dist;
                uint8_t* const oe = out + len;
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
            if (e.op == 0x10) { *out++ = (uint8_t)e.val; continue; }
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
            if (!copy_match(out0, out, n, dist)) return -1;
            if (n < len) { pending_len_ = len - n; pending_dist_ = dist; return 0; }
        }
        return 0;
    }

    void* map_ = nullptr;
    size_t size_ = 0;
    const uint8_t *in_ = nullptr, *end_ = nullptr;
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
    std::string error_;
};

}  // namespace mblhost
