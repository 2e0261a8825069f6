// tsv.hpp -- ordered TSV writer, replaces gather_write (lib.rs:612-644).
//
// Output text is the reference's: header `sequence1\tsequence2\tdistance`, one line per pair,
// integers as `{}` (lib.rs:627), floats as `{:.12}` (lib.rs:631): exact decimal expansion rounded
// half-to-even at 12 places, `NaN`, `inf`, `-inf`, and the sign of -0.0 kept (Rust Display for f64).
// Panels arrive serially and in output order from the engine's sink, so no reorder map is needed.
// A panel is cut into chunks of 65,536 results; a persistent pool of `threads` workers (the calling thread is one
// of them) grabs chunks in order, formats each into a private buffer (ids copied from a packed arena with fixed-size
// moves, counts through a 65,536-entry text table, floats through an exact 128-bit fixed-point conversion) and issues
// its write() when its turn comes, so formatting is parallel and the output stays in order.  When the output is a
// regular file (not O_APPEND) the turn only hands out the file offset and the pwrite() calls themselves run in parallel.
// A BrokenPipe on the output ends the process with status 0 (lib.rs:598-608).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/distance_gpu.h"

namespace host {

// Appends the reference's `{:.12}` text of d to out.
void format_float12(double d, std::string& out);
void format_u32(uint32_t v, std::string& out);
// The same into a raw buffer (needs up to 400 / 10 bytes); returns the end.
char* put_float12(char* p, double d);
char* put_u32(char* p, uint32_t v);

// ids, each followed by a TAB, back to back in one arena (with slack so short ids can be copied with two 16-byte moves)
class IdTable {
public:
    static constexpr size_t kSlack = 64;
    void sync(const std::vector<std::string>& ids);   // appends the ids not seen yet (stream mode grows its list)
    void clear() { arena_.clear(); off_.clear(); len_.clear(); used_ = 0; max_len_ = 0; }
    const char* ptr(uint64_t k) const { return arena_.data() + off_[k]; }
    uint32_t len(uint64_t k) const { return len_[k]; }   // incl. the TAB
    size_t max_len() const { return max_len_; }

private:
    std::vector<char> arena_;
    std::vector<size_t> off_;
    std::vector<uint32_t> len_;
    size_t used_ = 0, max_len_ = 0;
};

class TsvWriter {
public:
    // ids1 / ids2: SQUARE -> both = the alignment's ids; RECT -> file 0 / file 1;
    // STREAM -> ids1 = loaded ids, ids2 = the ids of the streamed records pushed so far (appended by the caller).
    TsvWriter(int fd, int threads);
    ~TsvWriter();
    TsvWriter(const TsvWriter&) = delete;
    TsvWriter& operator=(const TsvWriter&) = delete;
    void write_header();
    void set_ids(const std::vector<std::string>* ids1, const std::vector<std::string>* ids2) {
        ids1_ = ids1; ids2_ = ids2;
        ids1_tab_.clear(); ids2_tab_.clear();
    }
    // DG_RESULT_COUNTS16 panels (raw / jc69 / k80 / tn93 as integer counts): which measure to evaluate on the host, and for
    // tn93 the A,T,G,C counts (4 per record) of the row / column alignment (fastaio.rs:53-66)
    void set_counts_measure(int measure, const uint32_t* acgt_rows, const uint32_t* acgt_cols) {
        counts_measure_ = measure; acgt_rows_ = acgt_rows; acgt_cols_ = acgt_cols;
    }
    void write_panel(const dg_panel& p);  // throws DistanceError on an io error other than BrokenPipe
    void flush();
    uint64_t lines() const { return lines_; }

private:
    struct Pool;
    void write_all(const char* p, size_t n);
    void write_at(const char* p, size_t n, uint64_t off);
    void worker_loop();
    void run_chunks();
    template <int KIND>
    size_t format_chunk(const dg_panel& p, const std::vector<uint64_t>& row_start, uint64_t k0, uint64_t k1, std::vector<char>& buf);
    int fd_;
    int threads_;
    Pool* pool_;
    bool positioned_ = false;   // regular, non-append output: chunks are written with pwrite at offsets handed out in order
    uint64_t file_off_ = 0;
    IdTable ids1_tab_, ids2_tab_;
    const std::vector<std::string>* ids1_ = nullptr;
    const std::vector<std::string>* ids2_ = nullptr;
    uint64_t lines_ = 0;
    int counts_measure_ = -1;
    const uint32_t* acgt_rows_ = nullptr;
    const uint32_t* acgt_cols_ = nullptr;
};

}  // namespace host
