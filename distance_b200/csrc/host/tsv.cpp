#include "tsv.hpp"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unistd.h>

#include "fasta.hpp"

namespace host {

void format_u32(uint32_t v, std::string& out) {
    char b[12];
    int n = 0;
    do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) out.push_back(b[--n]);
}

// Exact `{:.12}`: d = m * 2^e with a 53-bit m; d * 10^12 = (m * 10^12) >> -e, and m * 10^12 < 2^93 fits
// unsigned __int128, so the quotient and the discarded remainder are exact and round-half-even is
// decided on the true binary value, the same digits Rust's (and glibc's) exact algorithms print.
void format_float12(double d, std::string& out) {
    if (std::isnan(d)) { out += "NaN"; return; }
    if (std::isinf(d)) { out += d > 0 ? "inf" : "-inf"; return; }
    if (std::signbit(d)) { out.push_back('-'); d = -d; }
    int e;
    const double fr = std::frexp(d, &e);  // d = fr * 2^e, fr in [0.5, 1)
    if (d != 0.0 && (e > 30 || e < -1000)) {
        // huge or subnormal-range values: the generic exact path of the C library
        char b[400];
        const int n = snprintf(b, sizeof b, "%.12f", d);
        out.append(b, (size_t)n);
        return;
    }
    unsigned __int128 q = 0;
    if (d != 0.0) {
        const uint64_t m = (uint64_t)std::ldexp(fr, 53);  // 53-bit integer mantissa
        const int sh = 53 - e;                            // d = m * 2^-sh, sh in [23, 1053]
        const unsigned __int128 prod = (unsigned __int128)m * 1000000000000ull;
        if (sh >= 128) {
            q = 0;  // prod < 2^93 < 2^127 = half an ulp of the last place: rounds to 0
        } else {
            q = prod >> sh;
            const unsigned __int128 rem = prod & ((((unsigned __int128)1) << sh) - 1);
            const unsigned __int128 half = ((unsigned __int128)1) << (sh - 1);
            if (rem > half || (rem == half && (q & 1))) q += 1;
        }
    }
    const uint64_t ip = (uint64_t)(q / 1000000000000ull);
    uint64_t fp = (uint64_t)(q % 1000000000000ull);
    char b[24];
    int n = 0;
    uint64_t v = ip;
    do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) out.push_back(b[--n]);
    out.push_back('.');
    char f[12];
    for (int i = 11; i >= 0; i--) { f[i] = (char)('0' + fp % 10); fp /= 10; }
    out.append(f, 12);
}

TsvWriter::TsvWriter(int fd, int threads) : fd_(fd), threads_(std::max(1, threads)) {}

void TsvWriter::write_all(const char* p, size_t n) {
    while (n) {
        ssize_t w = ::write(fd_, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            if (errno == EPIPE) std::exit(0);  // handle_broken_pipe, lib.rs:598-608
            throw io_error_os(errno);
        }
        p += w;
        n -= (size_t)w;
    }
}

void TsvWriter::write_header() {
    static const char h[] = "sequence1\tsequence2\tdistance\n";  // lib.rs:613
    write_all(h, sizeof h - 1);
}

void TsvWriter::flush() {}

namespace {

struct Chunk {
    uint64_t row0, row1;   // major rows [row0, row1)
    uint64_t data_off;     // index of the first result of row0 within the panel
    std::string text;
};

inline void emit(std::string& t, const std::string& a, const std::string& b, const dg_panel& p, uint64_t k) {
    t += a;
    t.push_back('\t');
    t += b;
    t.push_back('\t');
    if (p.result_kind == DG_RESULT_U32) format_u32(static_cast<const uint32_t*>(p.data)[k], t);
    else if (p.result_kind == DG_RESULT_U16) format_u32(static_cast<const uint16_t*>(p.data)[k], t);
    else format_float12(static_cast<const double*>(p.data)[k], t);
    t.push_back('\n');
}

}  // namespace

void TsvWriter::write_panel(const dg_panel& p) {
    if (p.n_results == 0) return;
    const std::vector<std::string>& id1 = *ids1_;
    const std::vector<std::string>& id2 = *ids2_;
    // cut the panel's rows into ~4 chunks per thread of roughly equal result count
    const uint64_t rows = p.row_end - p.row_begin;
    const uint64_t want = std::max<uint64_t>(1, std::min<uint64_t>(rows, (uint64_t)threads_ * 4));
    const uint64_t per = (p.n_results + want - 1) / want;
    std::vector<Chunk> chunks;
    {
        uint64_t off = 0, r = p.row_begin;
        while (r < p.row_end) {
            Chunk c;
            c.row0 = r;
            c.data_off = off;
            uint64_t acc = 0;
            while (r < p.row_end && (acc < per || acc == 0)) {
                acc += p.mode == DG_MODE_SQUARE ? (p.n_cols - 1 - r) : p.n_cols;
                r++;
            }
            c.row1 = r;
            off += acc;
            chunks.push_back(std::move(c));
        }
    }
    auto work = [&](size_t first, size_t step) {
        for (size_t ci = first; ci < chunks.size(); ci += step) {
            Chunk& c = chunks[ci];
            std::string& t = c.text;
            uint64_t k = c.data_off;
            if (p.mode == DG_MODE_SQUARE) {
                const uint64_t n = p.n_cols;
                for (uint64_t i = c.row0; i < c.row1; i++)
                    for (uint64_t j = i + 1; j < n; j++) emit(t, id1[i], id1[j], p, k++);
            } else if (p.mode == DG_MODE_RECT) {
                for (uint64_t i = c.row0; i < c.row1; i++)
                    for (uint64_t j = 0; j < p.n_cols; j++) emit(t, id1[i], id2[j], p, k++);
            } else {  // STREAM: rows = streamed records, columns = loaded records (lib.rs:322-331)
                for (uint64_t r = c.row0; r < c.row1; r++)
                    for (uint64_t i = 0; i < p.n_cols; i++) emit(t, id1[i], id2[r], p, k++);
            }
        }
    };
    const size_t nt = std::min<size_t>((size_t)threads_, chunks.size());
    if (nt <= 1) {
        work(0, 1);
    } else {
        std::vector<std::thread> th;
        for (size_t t = 0; t < nt; t++) th.emplace_back(work, t, nt);
        for (auto& t : th) t.join();
    }
    for (auto& c : chunks) write_all(c.text.data(), c.text.size());
    lines_ += p.n_results;
}

}  // namespace host
