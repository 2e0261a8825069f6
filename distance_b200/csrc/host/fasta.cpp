#include "fasta.hpp"

#include <atomic>
#include <cctype>
#include <chrono>
#include <cerrno>
#include <cstring>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace host {

namespace {

// encoding.rs:7-38: the 32 accepted characters (15 IUPAC letters in both cases, '-', '?')
struct ValidTable {
    bool ok[256];
    ValidTable() {
        std::memset(ok, 0, sizeof ok);
        for (const char* p = "AGCTRMWSKYVHDBN"; *p; p++) {
            ok[(uint8_t)*p] = true;
            ok[(uint8_t)(*p + 32)] = true;
        }
        ok[(uint8_t)'-'] = true;
        ok[(uint8_t)'?'] = true;
    }
};
const ValidTable g_valid;

// Rust `{:?}` of a String: quotes, backslashes and control characters escaped.
std::string rust_debug_str(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\n': o += "\\n"; break;
        case '\r': o += "\\r"; break;
        case '\t': o += "\\t"; break;
        default:
            if (c < 0x20 || c == 0x7f) {
                char b[16];
                snprintf(b, sizeof b, "\\u{%x}", c);
                o += b;
            } else {
                o += (char)c;
            }
        }
    }
    return o + "\"";
}

// `*nuc as char` (fastaio.rs:112): a byte >= 0x80 becomes the Latin-1 code point, printed as UTF-8.
std::string byte_as_char(uint8_t c) {
    std::string o;
    if (c < 0x80) {
        o += (char)c;
    } else {
        o += (char)(0xC0 | (c >> 6));
        o += (char)(0x80 | (c & 0x3F));
    }
    return o;
}

const char* rust_errno_kind(int e) {
    switch (e) {
    case ENOENT: return "NotFound";
    case EACCES: case EPERM: return "PermissionDenied";
    case EISDIR: return "IsADirectory";
    case EEXIST: return "AlreadyExists";
    case EPIPE: return "BrokenPipe";
    case ENOSPC: return "StorageFull";
    case EINVAL: return "InvalidInput";
    default: return "Uncategorized";
    }
}

std::string invalid_nuc_message(const std::string& id, uint8_t c) {  // fastaio.rs:89-91
    return "Invalid nucleotide character in record '" + id + "': '" + byte_as_char(c) + "'";
}

}  // namespace

DistanceError message_error(const std::string& msg) { return DistanceError("Message(" + rust_debug_str(msg) + ")"); }

DistanceError io_error_custom(const std::string& msg) {
    return DistanceError("IOError(Custom { kind: Other, error: " + rust_debug_str(msg) + " })");
}

DistanceError io_error_os(int err) {
    return DistanceError("IOError(Os { code: " + std::to_string(err) + ", kind: " + rust_errno_kind(err) +
                         ", message: " + rust_debug_str(std::strerror(err)) + " })");
}

bool valid_nucleotide(uint8_t c) { return g_valid.ok[c]; }

void validate_record(const std::string& id, const uint8_t* seq, uint64_t len) {
    for (uint64_t i = 0; i < len; i++)
        if (!g_valid.ok[seq[i]]) throw message_error(invalid_nuc_message(id, seq[i]));
}

FastaReader::FastaReader(int fd, bool validate) : fd_(fd), buf_(8u << 20), validate_(validate) {}

FastaReader::FastaReader(const char* data, size_t size, bool validate)
    : fd_(-1), mem_(data), mem_size_(size), buf_(1u << 20), validate_(validate) {}

bool FastaReader::fill() {
    if (eof_) return false;
    pos_ = end_ = 0;
    if (fd_ < 0) {
        const size_t n = std::min(buf_.size(), mem_size_ - mem_pos_);
        if (n == 0) { eof_ = true; return false; }
        std::memcpy(buf_.data(), mem_ + mem_pos_, n);
        mem_pos_ += n;
        end_ = n;
        return true;
    }
    for (;;) {
        ssize_t r = ::read(fd_, buf_.data(), buf_.size());
        if (r < 0) {
            if (errno == EINTR) continue;
            throw io_error_os(errno);
        }
        if (r == 0) { eof_ = true; return false; }
        end_ = (size_t)r;
        return true;
    }
}

bool FastaReader::read_line_raw() {
    line_.clear();
    bool got = false;
    for (;;) {
        if (pos_ == end_ && !fill()) return got;
        got = true;
        const char* p = buf_.data() + pos_;
        const char* nl = (const char*)std::memchr(p, '\n', end_ - pos_);
        if (nl) {
            line_.append(p, nl - p);
            pos_ = (size_t)(nl - buf_.data()) + 1;
            return true;
        }
        line_.append(p, end_ - pos_);
        pos_ = end_;
    }
}

// rust-bio 1.6.0 (io/fasta.rs, Reader::read) reads every line into a String -- so a line that is not valid UTF-8 fails with
// io::ErrorKind::InvalidData --, trims it with str::trim_end() and splits the header at the first char::is_whitespace.  Both
// use the Unicode White_Space property: U+0009..000D, U+0020, U+0085, U+00A0, U+1680, U+2000..200A, U+2028, U+2029,
// U+202F, U+205F, U+3000.
static size_t ws3(unsigned char a, unsigned char b, unsigned char c) {   // a three-byte white-space character?
    if (a == 0xE1) return (b == 0x9A && c == 0x80) ? 3 : 0;                                   // U+1680
    if (a == 0xE3) return (b == 0x80 && c == 0x80) ? 3 : 0;                                   // U+3000
    if (a != 0xE2) return 0;
    if (b == 0x80) return (c >= 0x80 && c <= 0x8A) || c == 0xA8 || c == 0xA9 || c == 0xAF ? 3 : 0;   // U+2000..200A, 2028, 2029, 202F
    return (b == 0x81 && c == 0x9F) ? 3 : 0;                                                  // U+205F
}
static bool ascii_ws(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }
// bytes of the white-space character that starts at p[0] (n bytes are readable), or 0
static size_t ws_at_start(const char* p, size_t n) {
    const unsigned char a = (unsigned char)p[0];
    if (a < 0x80) return ascii_ws(a) ? 1 : 0;
    if (a == 0xC2) return n >= 2 && ((unsigned char)p[1] == 0x85 || (unsigned char)p[1] == 0xA0) ? 2 : 0;   // U+0085, U+00A0
    return n >= 3 ? ws3(a, (unsigned char)p[1], (unsigned char)p[2]) : 0;
}
// bytes of the white-space character that ends at p[n - 1], or 0
static size_t ws_at_end(const char* p, size_t n) {
    const unsigned char z = (unsigned char)p[n - 1];
    if (z < 0x80) return ascii_ws(z) ? 1 : 0;
    if (n >= 2 && (unsigned char)p[n - 2] == 0xC2) return z == 0x85 || z == 0xA0 ? 2 : 0;
    return n >= 3 ? ws3((unsigned char)p[n - 3], (unsigned char)p[n - 2], z) : 0;
}
static size_t trimmed_len(const std::string& s) {   // str::trim_end
    size_t n = s.size();
    while (n > 0) {
        const size_t k = ws_at_end(s.data(), n);
        if (k == 0) break;
        n -= k;
    }
    return n;
}
// core::str::from_utf8: no overlong forms, no surrogates, nothing above U+10FFFF
static bool valid_utf8(const char* p, size_t n) {
    size_t i = 0;
    while (i < n) {
        const unsigned char a = (unsigned char)p[i];
        if (a < 0x80) { i++; continue; }
        auto cont = [&](size_t k, unsigned char lo = 0x80, unsigned char hi = 0xBF) {
            return i + k < n && (unsigned char)p[i + k] >= lo && (unsigned char)p[i + k] <= hi;
        };
        if (a >= 0xC2 && a <= 0xDF) { if (!cont(1)) return false; i += 2; }
        else if (a == 0xE0) { if (!cont(1, 0xA0, 0xBF) || !cont(2)) return false; i += 3; }
        else if ((a >= 0xE1 && a <= 0xEC) || a == 0xEE || a == 0xEF) { if (!cont(1) || !cont(2)) return false; i += 3; }
        else if (a == 0xED) { if (!cont(1, 0x80, 0x9F) || !cont(2)) return false; i += 3; }
        else if (a == 0xF0) { if (!cont(1, 0x90, 0xBF) || !cont(2) || !cont(3)) return false; i += 4; }
        else if (a >= 0xF1 && a <= 0xF3) { if (!cont(1) || !cont(2) || !cont(3)) return false; i += 4; }
        else if (a == 0xF4) { if (!cont(1, 0x80, 0x8F) || !cont(2) || !cont(3)) return false; i += 4; }
        else return false;
    }
    return true;
}
static bool has_high_byte(const char* p, size_t n) {
    for (size_t i = 0; i < n; i++)
        if ((unsigned char)p[i] >= 0x80) return true;
    return false;
}

bool FastaReader::read_line() {
    const bool got = read_line_raw();
    // BufRead::read_line: io::Error::INVALID_UTF8, a SimpleMessage (its Debug form is `Error { kind, message }`)
    if (got && has_high_byte(line_.data(), line_.size()) && !valid_utf8(line_.data(), line_.size()))
        throw DistanceError("IOError(Error { kind: InvalidData, message: \"stream did not contain valid UTF-8\" })");
    return got;
}

bool FastaReader::next(std::string& id, std::vector<uint8_t>& seq) {
    if (!have_line_) {
        if (!read_line()) return false;  // EOF
        have_line_ = true;
    }
    if (line_.empty() || line_[0] != '>') throw io_error_custom("Expected > at record start.");
    const size_t hl = trimmed_len(line_);
    size_t sp = 1;
    while (sp < hl && ws_at_start(line_.data() + sp, hl - sp) == 0) sp++;   // splitn(2, char::is_whitespace)
    id.assign(line_, 1, sp - 1);
    const bool has_desc = sp < hl;
    const size_t seq0 = seq.size();
    have_line_ = false;
    for (;;) {
        if (!read_line()) break;
        if (!line_.empty() && line_[0] == '>') { have_line_ = true; break; }
        const size_t n = trimmed_len(line_);
        seq.insert(seq.end(), line_.begin(), line_.begin() + n);
    }
    // fastaio.rs:111-113 runs on the complete record: an unreadable line (of this record or the next header) is reported first
    if (validate_) validate_record(id, seq.data() + seq0, seq.size() - seq0);
    if (id.empty() && !has_desc && seq.size() == seq0) return false;  // rust-bio: an empty record ends the iteration
    return true;
}

namespace {

// The whole input in memory.  Regular files: a huge-page backed buffer filled by parallel pread() calls (page-cache
// copies scale with threads, where first-touch faults on an mmap of the file serialise); pipes / stdin: read to EOF.
struct InputBytes {
    const char* data = nullptr;
    size_t size = 0;
    void* big = nullptr;
    std::vector<char> owned;
    InputBytes(int fd, int threads) {
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 && lseek(fd, 0, SEEK_CUR) == 0) {
            const size_t bytes = ((size_t)st.st_size + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
            if (posix_memalign(&big, 2u << 20, bytes) != 0) throw std::bad_alloc();
            madvise(big, bytes, MADV_HUGEPAGE);
            const size_t total = (size_t)st.st_size;
            int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
            T = (int)std::min<size_t>((size_t)std::min(T, 64), std::max<size_t>(1, total / (8u << 20)));
            std::atomic<int> err{0};
            std::atomic<size_t> got{0};
            auto rd = [&](int t) {
                size_t lo = total / T * t, hi = t == T - 1 ? total : total / T * (t + 1);
                while (lo < hi) {
                    ssize_t r = ::pread(fd, static_cast<char*>(big) + lo, std::min<size_t>(hi - lo, 64u << 20), (off_t)lo);
                    if (r < 0) { if (errno == EINTR) continue; err.store(errno); return; }
                    if (r == 0) break;   // the file shrank
                    lo += (size_t)r;
                    got += (size_t)r;
                }
            };
            std::vector<std::thread> th;
            for (int t = 1; t < T; t++) th.emplace_back(rd, t);
            rd(0);
            for (auto& x : th) x.join();
            if (err.load()) throw io_error_os(err.load());
            data = static_cast<const char*>(big);
            size = got.load() == total ? total : 0;
            if (size == total) return;
            std::free(big); big = nullptr;   // the file changed under us: fall through to the sequential read
            lseek(fd, 0, SEEK_SET);
        }
        owned.resize(1u << 20);
        size_t used = 0;
        for (;;) {
            if (used == owned.size()) owned.resize(owned.size() * 2);
            ssize_t r = ::read(fd, owned.data() + used, owned.size() - used);
            if (r < 0) {
                if (errno == EINTR) continue;
                throw io_error_os(errno);
            }
            if (r == 0) break;
            used += (size_t)r;
        }
        data = owned.data(); size = used;
    }
    ~InputBytes() { if (big) std::free(big); }
    InputBytes(const InputBytes&) = delete;
    InputBytes& operator=(const InputBytes&) = delete;
};

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// One record starting at `p` (which points at '>'): header, then sequence lines up to the next line that starts with
// '>' or `end`.  Calls line(ptr, trimmed_len) for every sequence line.  Returns the start of the next record (or end).
template <typename F>
inline const char* walk_record(const char* p, const char* end, const char*& id0, size_t& id_len, bool& has_desc, F&& line) {
    const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
    const char* hend = nl ? nl : end;
    const char* t = hend;
    while (t > p && is_space((unsigned char)t[-1])) t--;
    const char* q = p + 1;
    while (q < t && !std::isspace((unsigned char)*q)) q++;
    id0 = p + 1; id_len = (size_t)(q - (p + 1)); has_desc = q < t;
    if (has_high_byte(p, (size_t)(hend - p))) id_len = 0;   // Unicode white space / UTF-8 validity: the sequential reader decides
    p = nl ? nl + 1 : end;
    while (p < end && *p != '>') {
        nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
        const char* lend = nl ? nl : end;
        const char* e = lend;
        while (e > p && is_space((unsigned char)e[-1])) e--;
        line(p, (size_t)(e - p));
        p = nl ? nl + 1 : end;
    }
    return p;
}

// The parallel pass over whole records.  false = "not clean": the caller re-reads the same bytes sequentially for the
// exact error / edge case.  width: in = the width every record must have (0 = take the first record's), out = the width.
// dst_for(n) returns where the n x width bytes go (nullptr = refuse: too many records for the caller's buffer).
// The ids are appended to `ids`; n_out = records parsed.
template <typename DstFor>
bool parse_clean(const char* data, size_t size, int threads, uint64_t& width, DstFor&& dst_for, std::vector<std::string>& ids,
                 uint64_t& n_out) {
    if (size == 0 || data[0] != '>') return false;
    const char* end = data + size;
    if (width == 0) {   // width = length of the first record
        const char* id0; size_t idl; bool desc;
        walk_record(data, end, id0, idl, desc, [&](const char*, size_t n) { width += n; });
        if (width == 0) return false;
    }
    int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    size_t seg_bytes = 4u << 20;   // at least this much input per thread (DG_PARSE_SEG overrides it: tests cut tiny inputs)
    if (const char* e = std::getenv("DG_PARSE_SEG")) seg_bytes = std::max<size_t>(1, (size_t)std::strtoull(e, nullptr, 10));
    T = (int)std::min<size_t>((size_t)std::min(T, 64), std::max<size_t>(1, size / seg_bytes));
    // segment starts: the first '>' at a line start at or after size * t / T
    std::vector<const char*> seg(T + 1, end);
    seg[0] = data;
    for (int t = 1; t < T; t++) {
        const char* p = data + size / T * t;
        const char* found = end;
        while (p < end) {
            const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
            if (!nl || nl + 1 >= end) break;
            if (nl[1] == '>') { found = nl + 1; break; }
            p = nl + 1;
        }
        seg[t] = found;
    }
    for (int t = 1; t <= T; t++) if (seg[t] < seg[t - 1]) seg[t] = seg[t - 1];
    // pass 1: records per segment, every record exactly `width` long and with a non-empty id
    std::vector<uint64_t> count(T, 0);
    std::atomic<bool> clean{true};
    auto pass1 = [&](int t) {
        uint64_t n = 0;
        const char* p = seg[t];
        const char* e = seg[t + 1];
        while (p < e && clean.load(std::memory_order_relaxed)) {
            const char* id0; size_t idl; bool desc;
            uint64_t len = 0;
            p = walk_record(p, e, id0, idl, desc, [&](const char*, size_t k) { len += k; });
            if (len != width || idl == 0) { clean.store(false); return; }
            n++;
        }
        count[t] = n;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(pass1, t);
        pass1(0);
        for (auto& x : th) x.join();
    }
    if (!clean.load()) return false;
    std::vector<uint64_t> first(T + 1, 0);
    for (int t = 0; t < T; t++) first[t + 1] = first[t] + count[t];
    const uint64_t n = first[T];
    if (n == 0) return false;
    uint8_t* dst = dst_for(n);
    if (!dst) return false;
    const size_t id0_index = ids.size();
    ids.resize(id0_index + n);
    // pass 2: ids, validation (encoding.rs:7-38) and the copy
    auto pass2 = [&](int t) {
        uint64_t r = first[t];
        const char* p = seg[t];
        const char* e = seg[t + 1];
        while (p < e && clean.load(std::memory_order_relaxed)) {
            const char* id0; size_t idl; bool desc;
            uint8_t* out = dst + r * width;
            unsigned bad = 0;
            p = walk_record(p, e, id0, idl, desc, [&](const char* s, size_t k) {
                for (size_t i = 0; i < k; i++) bad |= (unsigned)!g_valid.ok[(uint8_t)s[i]];
                std::memcpy(out, s, k);
                out += k;
            });
            if (bad) { clean.store(false); return; }
            ids[id0_index + r].assign(id0, idl);
            r++;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(pass2, t);
        pass2(0);
        for (auto& x : th) x.join();
    }
    if (!clean.load()) {
        ids.resize(id0_index);
        return false;
    }
    n_out = n;
    return true;
}

bool parse_parallel(const char* data, size_t size, int threads, Alignment& a) {
    uint64_t width = 0, n = 0;
    const bool ok = parse_clean(data, size, threads, width, [&](uint64_t recs) {
        // uninitialised, 2 MB aligned, transparent huge pages requested: the pages are first touched by the parser's threads
        void* mem = nullptr;
        const size_t bytes = ((size_t)(recs * width) + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
        if (posix_memalign(&mem, 2u << 20, bytes) != 0) throw std::bad_alloc();
        madvise(mem, bytes, MADV_HUGEPAGE);
        a.fast_.reset(static_cast<uint8_t*>(mem));
        return a.fast_.get();
    }, a.ids, n);
    if (!ok) { a = Alignment(); return false; }
    a.width = width;
    return true;
}

}  // namespace

Alignment load_fasta(int fd, int threads) {
    const bool trace = std::getenv("DG_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    InputBytes in(fd, threads);
    const auto t1 = std::chrono::steady_clock::now();
    Alignment a;
    const bool ok = parse_parallel(in.data, in.size, threads, a);
    if (trace)
        fprintf(stderr, "[load_fasta] %zu bytes: input %.3f s, parallel parse %.3f s (%s)\n", in.size,
                std::chrono::duration<double>(t1 - t0).count(),
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(), ok ? "ok" : "fell back");
    if (ok) return a;
    a = Alignment();
    FastaReader rd(in.data, in.size);
    std::string id;
    bool first = true;
    for (;;) {
        const size_t before = a.seqs.size();
        if (!rd.next(id, a.seqs)) break;
        const uint64_t len = a.seqs.size() - before;
        if (first) {
            a.width = len;
            first = false;
        } else if (len != a.width) {  // fastaio.rs:188-190
            throw message_error("Different length sequences in alignment(s): " + std::to_string(len) + " vs " +
                                std::to_string(a.width));
        }
        a.ids.push_back(id);
    }
    if (a.ids.empty()) throw message_error("Empty FASTA file");  // fastaio.rs:195-197
    return a;
}

// ---- streamed file: blocks of whole records, parsed in parallel straight into the caller's (pinned) batch buffer ---------
StreamBlockParser::StreamBlockParser(int fd, uint64_t width, int threads) : fd_(fd), width_(width), threads_(threads) {
    struct stat st;
    regular_ = fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
    if (regular_) {
        const off_t cur = lseek(fd, 0, SEEK_CUR);
        regular_ = cur >= 0;
        offset_ = regular_ ? (uint64_t)cur : 0;
        file_size_ = (uint64_t)st.st_size;
    }
}

// Append up to `want` bytes of input to raw_; sets eof_ at the end.  Regular files: parallel pread().
void StreamBlockParser::fill(size_t want) {
    if (eof_ || want == 0) return;
    const size_t old = raw_len_;
    if (raw_cap_ < old + want) {
        const size_t cap = std::max(old + want, raw_cap_ + raw_cap_ / 2);
        std::unique_ptr<char[]> bigger(new char[cap]);   // not value-initialised
        if (old) std::memcpy(bigger.get(), raw_.get(), old);
        raw_ = std::move(bigger);
        raw_cap_ = cap;
    }
    size_t got = 0;
    if (regular_) {
        const size_t avail = offset_ < file_size_ ? (size_t)std::min<uint64_t>(want, file_size_ - offset_) : 0;
        int T = threads_ > 0 ? threads_ : (int)std::max(1u, std::thread::hardware_concurrency());
        T = (int)std::min<size_t>((size_t)std::min(T, 32), std::max<size_t>(1, avail / (8u << 20)));
        std::atomic<int> err{0};
        std::atomic<size_t> done{0};
        auto rd = [&](int t) {
            size_t lo = avail / T * t, hi = t == T - 1 ? avail : avail / T * (t + 1);
            while (lo < hi) {
                ssize_t r = ::pread(fd_, raw_.get() + old + lo, hi - lo, (off_t)(offset_ + lo));
                if (r < 0) { if (errno == EINTR) continue; err.store(errno); return; }
                if (r == 0) break;
                lo += (size_t)r; done += (size_t)r;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(rd, t);
        if (avail) rd(0);
        for (auto& x : th) x.join();
        if (err.load()) throw io_error_os(err.load());
        got = done.load() == avail ? avail : 0;
        if (done.load() != avail) regular_ = false;   // the file changed under us: continue with plain reads from here
        else offset_ += got;
        if (got < want && regular_) eof_ = true;
        if (!regular_) lseek(fd_, (off_t)offset_, SEEK_SET);
    }
    if (!regular_) {
        while (got < want) {
            ssize_t r = ::read(fd_, raw_.get() + old + got, want - got);
            if (r < 0) { if (errno == EINTR) continue; throw io_error_os(errno); }
            if (r == 0) { eof_ = true; break; }
            got += (size_t)r;
        }
    }
    raw_len_ = old + got;
}

uint64_t StreamBlockParser::next(uint8_t* dst, uint64_t max_records, std::vector<std::string>& ids) {
    if (done_) return 0;
    // every record occupies at least width + 2 input bytes ('>', '\n', the sequence): a block of at most
    // max_records * (width + 2) bytes holds at most max_records whole records
    const size_t first_target = (size_t)(max_records * (width_ + 2));
    size_t target = first_target;
    size_t cut = 0;
    for (;;) {
        if (raw_len_ < target) fill(target - raw_len_);
        if (raw_len_ == 0) { done_ = true; return 0; }
        if (raw_len_ > first_target) {
            // the block had to grow beyond the size that bounds its record count (records longer than width + 2 bytes:
            // long headers, wrapped lines, a width mismatch ahead): count the record starts and take at most max_records
            uint64_t starts = 0;
            size_t boundary = 0;   // start of record number max_records (0-based), if there is one
            for (size_t p = 1; p < raw_len_; p++) {
                const char* q = static_cast<const char*>(std::memchr(raw_.get() + p, '>', raw_len_ - p));
                if (!q) break;
                p = (size_t)(q - raw_.get());
                if (raw_[p - 1] == '\n' && ++starts == max_records) { boundary = p; break; }
            }
            if (boundary) { cut = boundary; break; }           // max_records whole records (record 0 starts at byte 0)
            if (eof_) { cut = raw_len_; break; }               // fewer: the rest of the input
            if (starts > 0) {                                  // some whole records and a partial one: cut at the last start
                for (size_t p = raw_len_ - 1; p > 0; p--)
                    if (raw_[p] == '>' && raw_[p - 1] == '\n') { cut = p; break; }
                break;
            }
        } else {
            if (eof_) { cut = raw_len_; break; }
            // the last record start inside the block: everything before it is whole records
            cut = 0;
            for (size_t p = raw_len_ - 1; p > 0; p--)
                if (raw_[p] == '>' && raw_[p - 1] == '\n') { cut = p; break; }
            if (cut > 0) break;
        }
        target *= 2;   // not even one whole record yet: take more
    }
    uint64_t n = 0;
    uint64_t w = width_;
    const size_t ids_before = ids.size();
    const bool clean = parse_clean(raw_.get(), cut, threads_, w, [&](uint64_t recs) { return recs <= max_records ? dst : nullptr; },
                                   ids, n);
    if (!clean) {
        // the sequential reader over the same bytes, with stream_fasta's order of checks: width (fastaio.rs:246-248), then
        // the nucleotides (:250-254)
        ids.resize(ids_before);
        FastaReader rd(raw_.get(), cut, /*validate=*/false);
        std::vector<uint8_t> rec;
        std::string id;
        n = 0;
        for (;;) {
            rec.clear();
            if (!rd.next(id, rec)) {
                if (!rd.at_end()) done_ = true;   // an empty record ends the iteration (rust-bio), whatever follows
                break;
            }
            if (rec.size() != width_)
                throw message_error("Different length sequences in alignment(s): " + std::to_string(rec.size()) + " vs " +
                                    std::to_string(width_));
            validate_record(id, rec.data(), rec.size());
            if (n >= max_records) throw message_error("internal: stream block holds more records than the batch");
            std::memcpy(dst + n * width_, rec.data(), width_);
            ids.push_back(id);
            n++;
        }
    }
    std::memmove(raw_.get(), raw_.get() + cut, raw_len_ - cut);   // carry the partial last record over
    raw_len_ -= cut;
    if (eof_ && raw_len_ == 0) done_ = true;
    return n;
}

void check_same_width(const Alignment& a, const Alignment& b) {  // fastaio.rs:206-208
    if (a.width != b.width)
        throw message_error("Different length sequences in alignment(s): " + std::to_string(a.width) + " vs " +
                            std::to_string(b.width));
}

}  // namespace host
