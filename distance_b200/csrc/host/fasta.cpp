#include "fasta.hpp"

#include <cctype>
#include <cerrno>
#include <cstring>
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

bool FastaReader::fill() {
    if (eof_) return false;
    pos_ = end_ = 0;
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

bool FastaReader::read_line() {
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

static size_t trimmed_len(const std::string& s) {
    size_t n = s.size();
    while (n > 0) {
        unsigned char c = (unsigned char)s[n - 1];
        if (c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f') n--; else break;
    }
    return n;
}

bool FastaReader::next(std::string& id, std::vector<uint8_t>& seq) {
    if (!have_line_) {
        if (!read_line()) return false;  // EOF
        have_line_ = true;
    }
    if (line_.empty() || line_[0] != '>') throw io_error_custom("Expected > at record start.");
    const size_t hl = trimmed_len(line_);
    size_t sp = 1;
    while (sp < hl && !std::isspace((unsigned char)line_[sp])) sp++;
    id.assign(line_, 1, sp - 1);
    const bool has_desc = sp < hl;
    const size_t seq0 = seq.size();
    have_line_ = false;
    for (;;) {
        if (!read_line()) break;
        if (!line_.empty() && line_[0] == '>') { have_line_ = true; break; }
        const size_t n = trimmed_len(line_);
        if (validate_) validate_record(id, (const uint8_t*)line_.data(), n);  // fastaio.rs:111-113
        seq.insert(seq.end(), line_.begin(), line_.begin() + n);
    }
    if (id.empty() && !has_desc && seq.size() == seq0) return false;  // rust-bio: an empty record ends the iteration
    return true;
}

Alignment load_fasta(int fd) {
    Alignment a;
    FastaReader rd(fd);
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

void check_same_width(const Alignment& a, const Alignment& b) {  // fastaio.rs:206-208
    if (a.width != b.width)
        throw message_error("Different length sequences in alignment(s): " + std::to_string(a.width) + " vs " +
                            std::to_string(b.width));
}

}  // namespace host
