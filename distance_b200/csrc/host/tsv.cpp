#include "tsv.hpp"
#include "measures.hpp"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <mutex>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

#include "fasta.hpp"

namespace host {

namespace {

// "00" "01" ... "99"
struct Digits2 {
    char t[200];
    Digits2() {
        for (int i = 0; i < 100; i++) { t[2 * i] = (char)('0' + i / 10); t[2 * i + 1] = (char)('0' + i % 10); }
    }
};
const Digits2 kD2;

// Every uint16 as text followed by '\n' in 8 bytes; byte 7 = length incl. the newline.  Counts of the n / n_high
// measures are small numbers that repeat, so the hot lines of this 512 KB table stay in L1.
struct U16Lines {
    uint64_t w[65536];
    U16Lines() {
        for (uint32_t v = 0; v < 65536; v++) {
            char b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            char tmp[8];
            int n = 0;
            uint32_t x = v;
            do { tmp[n++] = (char)('0' + x % 10); x /= 10; } while (x);
            for (int i = 0; i < n; i++) b[i] = tmp[n - 1 - i];
            b[n] = '\n';
            b[7] = (char)(n + 1);
            std::memcpy(&w[v], b, 8);
        }
    }
};
const U16Lines& u16_lines() {
    static const U16Lines t;
    return t;
}

inline char* put_u64(char* p, uint64_t v) {
    char b[20];
    int n = 0;
    while (v >= 100) { const uint64_t q = v / 100; const uint32_t r = (uint32_t)(v - q * 100); v = q; b[n++] = kD2.t[2 * r + 1]; b[n++] = kD2.t[2 * r]; }
    if (v >= 10) { b[n++] = kD2.t[2 * v + 1]; b[n++] = kD2.t[2 * v]; }
    else b[n++] = (char)('0' + v);
    while (n) *p++ = b[--n];
    return p;
}

}  // namespace

char* put_u32(char* p, uint32_t v) { return put_u64(p, v); }

// Exact `{:.12}`: d = m * 2^e with a 53-bit m; d * 10^12 = (m * 10^12) >> -e, and m * 10^12 < 2^93 fits
// unsigned __int128, so the quotient and the discarded remainder are exact and round-half-even is
// decided on the true binary value, the same digits Rust's (and glibc's) exact algorithms print.
char* put_float12(char* p, double d) {
    if (std::isnan(d)) { std::memcpy(p, "NaN", 3); return p + 3; }
    if (std::isinf(d)) {
        if (d > 0) { std::memcpy(p, "inf", 3); return p + 3; }
        std::memcpy(p, "-inf", 4);
        return p + 4;
    }
    if (std::signbit(d)) { *p++ = '-'; d = -d; }
    uint64_t bits;
    std::memcpy(&bits, &d, 8);
    const int be = (int)(bits >> 52);                 // biased exponent (sign already cleared)
    if (d != 0.0 && (be == 0 || be > 1023 + 29)) {
        // subnormals and values >= 2^30: the generic exact path of the C library (never a real distance)
        return p + snprintf(p, 400, "%.12f", d);
    }
    unsigned __int128 q = 0;
    if (d != 0.0) {
        const uint64_t m = (bits & ((1ull << 52) - 1)) | (1ull << 52);   // 53-bit integer mantissa: d = m * 2^(be - 1075)
        const int sh = 1075 - be;                                        // in [23, 1074]
        const unsigned __int128 prod = (unsigned __int128)m * 1000000000000ull;
        if (sh < 128) {   // else prod < 2^93 < 2^127 = half an ulp of the last place: rounds to 0
            q = prod >> sh;
            const unsigned __int128 rem = prod & ((((unsigned __int128)1) << sh) - 1);
            const unsigned __int128 half = ((unsigned __int128)1) << (sh - 1);
            if (rem > half || (rem == half && (q & 1))) q += 1;
        }
    }
    uint64_t ip, fp;
    if ((uint64_t)(q >> 64) == 0) {
        const uint64_t q64 = (uint64_t)q;
        ip = q64 / 1000000000000ull;
        fp = q64 - ip * 1000000000000ull;
    } else {
        ip = (uint64_t)(q / 1000000000000ull);
        fp = (uint64_t)(q % 1000000000000ull);
    }
    p = put_u64(p, ip);
    *p++ = '.';
    // twelve digits, two at a time
    const uint32_t hi = (uint32_t)(fp / 1000000ull), lo = (uint32_t)(fp - (uint64_t)hi * 1000000ull);
    const uint32_t h0 = hi / 10000, h1 = hi / 100 % 100, h2 = hi % 100;
    const uint32_t l0 = lo / 10000, l1 = lo / 100 % 100, l2 = lo % 100;
    std::memcpy(p, kD2.t + 2 * h0, 2); std::memcpy(p + 2, kD2.t + 2 * h1, 2); std::memcpy(p + 4, kD2.t + 2 * h2, 2);
    std::memcpy(p + 6, kD2.t + 2 * l0, 2); std::memcpy(p + 8, kD2.t + 2 * l1, 2); std::memcpy(p + 10, kD2.t + 2 * l2, 2);
    return p + 12;
}

void format_u32(uint32_t v, std::string& out) {
    char b[16];
    out.append(b, (size_t)(put_u32(b, v) - b));
}

void format_float12(double d, std::string& out) {
    char b[420];
    out.append(b, (size_t)(put_float12(b, d) - b));
}

// ---- id table: every id followed by '\t' in one arena, copied with fixed-size moves -------------------------------
void IdTable::sync(const std::vector<std::string>& ids) {
    for (size_t k = off_.size(); k < ids.size(); k++) {
        const std::string& s = ids[k];
        off_.push_back(used_);
        len_.push_back((uint32_t)s.size() + 1);
        if (arena_.size() < used_ + s.size() + 1 + kSlack) arena_.resize(std::max<size_t>(arena_.size() * 2, used_ + s.size() + 1 + kSlack));
        std::memcpy(arena_.data() + used_, s.data(), s.size());
        arena_[used_ + s.size()] = '\t';
        used_ += s.size() + 1;
        max_len_ = std::max<size_t>(max_len_, s.size() + 1);
    }
}

// ---- worker pool ------------------------------------------------------------------------------------------------
struct TsvWriter::Pool {
    std::mutex mu;
    std::condition_variable cv_job, cv_done, cv_turn;
    uint64_t generation = 0;
    bool stop = false;
    int running = 0;
    // the current panel
    const dg_panel* panel = nullptr;
    std::vector<uint64_t> row_start;          // index of the first result of every major row of the panel (+ total)
    uint64_t n_chunks = 0;
    std::atomic<uint64_t> next_chunk{0};
    uint64_t write_turn = 0;
    std::string error;
    std::vector<std::thread> threads;
};

TsvWriter::TsvWriter(int fd, int threads) : fd_(fd), threads_(std::max(1, threads)), pool_(new Pool) {
    // A regular file that is not in append mode takes positioned writes: the turn only hands out the offset, and the
    // page-cache copies of different chunks run in parallel.  Pipes, terminals, /dev/null and O_APPEND files keep the
    // ordered write() calls.
    struct stat st;
    const int fl = fcntl(fd, F_GETFL);
    const off_t cur = lseek(fd, 0, SEEK_CUR);
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && fl >= 0 && !(fl & O_APPEND) && cur >= 0) {
        positioned_ = true;
        file_off_ = (uint64_t)cur;
    }
    for (int t = 1; t < threads_; t++) pool_->threads.emplace_back([this] { worker_loop(); });
}

TsvWriter::~TsvWriter() {
    {
        std::lock_guard<std::mutex> lk(pool_->mu);
        pool_->stop = true;
    }
    pool_->cv_job.notify_all();
    for (auto& t : pool_->threads) t.join();
    delete pool_;
}

void TsvWriter::write_all(const char* p, size_t n) {
    while (n) {
        ssize_t w = ::write(fd_, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            if (errno == EPIPE) std::_Exit(0);  // handle_broken_pipe, lib.rs:598-608
            throw io_error_os(errno);
        }
        p += w;
        n -= (size_t)w;
    }
}

void TsvWriter::write_at(const char* p, size_t n, uint64_t off) {
    while (n) {
        ssize_t w = ::pwrite(fd_, p, n, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            throw io_error_os(errno);
        }
        p += w;
        off += (uint64_t)w;
        n -= (size_t)w;
    }
}

void TsvWriter::write_header() {
    static const char h[] = "sequence1\tsequence2\tdistance\n";  // lib.rs:613
    if (positioned_) {
        write_at(h, sizeof h - 1, file_off_);
        file_off_ += sizeof h - 1;
    } else {
        write_all(h, sizeof h - 1);
    }
}

void TsvWriter::flush() {
    if (positioned_) lseek(fd_, (off_t)file_off_, SEEK_SET);   // leave the descriptor where sequential writes would have
}

namespace {
constexpr uint64_t kChunk = 1 << 16;   // results per formatting chunk (a few MB of text)

inline char* copy_id(char* p, const char* src, uint32_t len) {
    if (len <= 32) {   // arena and output buffer both carry >= 32 bytes of slack
        std::memcpy(p, src, 16);
        std::memcpy(p + 16, src + 16, 16);
    } else {
        std::memcpy(p, src, len);
    }
    return p + len;
}

// DG_RESULT_COUNTS16: four uint16 counts per pair -> the f64 of measures.rs with the host's libm (measures.hpp)
inline double counts_value(int measure, const void* data, uint64_t k, const uint32_t* q_acgt, const uint32_t* t_acgt) {
    const uint16_t* c = static_cast<const uint16_t*>(data) + 4 * k;
    switch (measure) {
    case DG_MEASURE_RAW: return raw_from_counts(c[0], c[1]);
    case DG_MEASURE_JC69: return jc69_from_counts(c[0], c[1]);
    case DG_MEASURE_K80: return k80_from_counts(c[0], c[1], c[2]);
    default: return tn93_from_counts(c[0], c[1], c[2], c[3], q_acgt, t_acgt);
    }
}

template <int KIND>
inline char* put_value(char* p, const void* data, uint64_t k) {
    if (KIND == DG_RESULT_U16) {
        const uint64_t w = u16_lines().w[static_cast<const uint16_t*>(data)[k]];
        std::memcpy(p, &w, 8);
        return p + (w >> 56);
    }
    if (KIND == DG_RESULT_U32) {
        const uint32_t v = static_cast<const uint32_t*>(data)[k];
        if (v < 65536) {
            const uint64_t w = u16_lines().w[v];
            std::memcpy(p, &w, 8);
            return p + (w >> 56);
        }
        p = put_u32(p, v);
    } else {
        p = put_float12(p, static_cast<const double*>(data)[k]);
    }
    *p++ = '\n';
    return p;
}
}  // namespace

// Text of results [k0, k1) of the panel into buf (resized as needed); returns the byte count.
template <int KIND>
size_t TsvWriter::format_chunk(const dg_panel& p, const std::vector<uint64_t>& row_start, uint64_t k0, uint64_t k1,
                               std::vector<char>& buf) {
    constexpr bool kFloat = KIND == DG_RESULT_F64 || KIND == DG_RESULT_COUNTS16;
    const size_t line_max = ids1_tab_.max_len() + ids2_tab_.max_len() + (kFloat ? 420 : 16) + 64;
    const size_t typical = ids1_tab_.max_len() + ids2_tab_.max_len() + (kFloat ? 24 : 8);
    if (buf.size() < (k1 - k0) * typical + line_max) buf.resize((k1 - k0) * typical + line_max);
    char* out = buf.data();
    char* limit = buf.data() + buf.size() - line_max;
    auto grow = [&](char*& q) {   // only lines with huge float text (>= 2^30) can outgrow the typical size
        const size_t used = (size_t)(q - buf.data());
        buf.resize(buf.size() * 2 + line_max);
        out = buf.data();
        limit = buf.data() + buf.size() - line_max;
        q = out + used;
    };
    // the major row that holds k0
    uint64_t r = (uint64_t)(std::upper_bound(row_start.begin(), row_start.end(), k0) - row_start.begin()) - 1;
    char* q = out;
    uint64_t k = k0;
    while (k < k1) {
        while (row_start[r + 1] <= k) r++;   // rows without results (the last square row) are skipped
        const uint64_t row = p.row_begin + r;
        const uint64_t in_row = k - row_start[r];
        const uint64_t row_end_k = std::min<uint64_t>(k1, row_start[r + 1]);
        if (p.mode == DG_MODE_STREAM) {
            // rows = streamed records, columns = loaded records; TSV: id1 = loaded id, id2 = streamed id (lib.rs:322-331)
            const char* b = ids2_tab_.ptr(row);
            const uint32_t bl = ids2_tab_.len(row);
            for (uint64_t i = in_row; k < row_end_k; k++, i++) {
                if (q > limit) grow(q);
                q = copy_id(q, ids1_tab_.ptr(i), ids1_tab_.len(i));
                q = copy_id(q, b, bl);
                q = put_value<KIND>(q, p.data, k);
            }
        } else {
            const IdTable& cols = p.mode == DG_MODE_SQUARE ? ids1_tab_ : ids2_tab_;
            const char* a = ids1_tab_.ptr(row);
            const uint32_t al = ids1_tab_.len(row);
            for (uint64_t j = (p.mode == DG_MODE_SQUARE ? row + 1 : 0) + in_row; k < row_end_k; k++, j++) {
                if (q > limit) grow(q);
                q = copy_id(q, a, al);
                q = copy_id(q, cols.ptr(j), cols.len(j));
                if (KIND == DG_RESULT_COUNTS16) {
                    // row record = the reference's `query`, column record = `target` (lib.rs:430-434)
                    const uint32_t* qa = acgt_rows_ ? acgt_rows_ + 4 * row : nullptr;
                    const uint32_t* ta = acgt_cols_ ? acgt_cols_ + 4 * j : nullptr;
                    q = put_float12(q, counts_value(counts_measure_, p.data, k, qa, ta));
                    *q++ = '\n';
                } else {
                    q = put_value<KIND>(q, p.data, k);
                }
            }
        }
    }
    return (size_t)(q - out);
}

// Grab chunks in order, format each into a private buffer, write it when its turn comes: formatting runs on every
// thread, the write() calls stay in output order.
void TsvWriter::run_chunks() {
    Pool& P = *pool_;
    std::vector<char> buf;
    for (;;) {
        const uint64_t c = P.next_chunk.fetch_add(1);
        if (c >= P.n_chunks) break;
        const dg_panel& p = *P.panel;
        const uint64_t k0 = c * kChunk, k1 = std::min<uint64_t>(p.n_results, k0 + kChunk);
        size_t n = 0;
        std::string err;
        try {
            if (p.result_kind == DG_RESULT_U16) n = format_chunk<DG_RESULT_U16>(p, P.row_start, k0, k1, buf);
            else if (p.result_kind == DG_RESULT_U32) n = format_chunk<DG_RESULT_U32>(p, P.row_start, k0, k1, buf);
            else if (p.result_kind == DG_RESULT_COUNTS16) n = format_chunk<DG_RESULT_COUNTS16>(p, P.row_start, k0, k1, buf);
            else n = format_chunk<DG_RESULT_F64>(p, P.row_start, k0, k1, buf);
        } catch (const std::exception& e) {
            err = e.what();
        }
        std::unique_lock<std::mutex> lk(P.mu);
        P.cv_turn.wait(lk, [&] { return P.write_turn == c; });
        if (positioned_) {
            // my turn = my offset; the copy into the file happens outside the turn, in parallel with the other chunks'
            const uint64_t off = file_off_;
            const bool go = err.empty() && P.error.empty();
            if (go) file_off_ += n;
            P.write_turn = c + 1;
            lk.unlock();
            P.cv_turn.notify_all();
            if (go) {
                try {
                    write_at(buf.data(), n, off);
                } catch (const DistanceError& e) {
                    err = e.what();
                }
            }
            if (!err.empty()) {
                std::lock_guard<std::mutex> lk2(P.mu);
                if (P.error.empty()) P.error = err;
            }
            continue;
        }
        if (err.empty() && P.error.empty()) {
            lk.unlock();
            try {
                write_all(buf.data(), n);
            } catch (const DistanceError& e) {
                err = e.what();
            }
            lk.lock();
        }
        if (!err.empty() && P.error.empty()) P.error = err;
        P.write_turn = c + 1;
        lk.unlock();
        P.cv_turn.notify_all();
    }
}

void TsvWriter::worker_loop() {
    Pool& P = *pool_;
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(P.mu);
            P.cv_job.wait(lk, [&] { return P.stop || P.generation != seen; });
            if (P.stop) return;
            seen = P.generation;
        }
        run_chunks();
        {
            std::lock_guard<std::mutex> lk(P.mu);
            P.running--;
        }
        P.cv_done.notify_one();
    }
}

void TsvWriter::write_panel(const dg_panel& p) {
    if (p.n_results == 0) return;
    ids1_tab_.sync(*ids1_);
    if (ids2_ != ids1_) ids2_tab_.sync(*ids2_);
    Pool& P = *pool_;
    // first result of every major row
    const uint64_t rows = p.row_end - p.row_begin;
    P.row_start.resize(rows + 1);
    uint64_t off = 0;
    for (uint64_t r = 0; r < rows; r++) {
        P.row_start[r] = off;
        off += p.mode == DG_MODE_SQUARE ? (p.n_cols - 1 - (p.row_begin + r)) : p.n_cols;
    }
    P.row_start[rows] = off;
    if (off != p.n_results) throw message_error("TSV writer: panel geometry does not match its result count");
    {
        std::lock_guard<std::mutex> lk(P.mu);
        P.panel = &p;
        P.n_chunks = (p.n_results + kChunk - 1) / kChunk;
        P.next_chunk.store(0);
        P.write_turn = 0;
        P.error.clear();
        P.running = (int)P.threads.size();
        P.generation++;
    }
    P.cv_job.notify_all();
    run_chunks();   // the calling thread formats too
    {
        std::unique_lock<std::mutex> lk(P.mu);
        P.cv_done.wait(lk, [&] { return P.running == 0; });
    }
    if (!P.error.empty()) throw DistanceError(P.error);
    lines_ += p.n_results;
}

}  // namespace host
