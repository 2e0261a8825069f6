// tsv.hpp -- ordered TSV writer, replaces gather_write (lib.rs:612-644).
//
// Output text is the reference's: header `sequence1\tsequence2\tdistance`, one line per pair,
// integers as `{}` (lib.rs:627), floats as `{:.12}` (lib.rs:631): exact decimal expansion rounded
// half-to-even at 12 places, `NaN`, `inf`, `-inf`, and the sign of -0.0 kept (Rust Display for f64).
// Panels arrive serially and in output order from the engine's sink, so no reorder map is needed;
// each panel is formatted by `threads` workers into per-chunk buffers that are written in order.
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

class TsvWriter {
public:
    // ids1 / ids2: SQUARE -> both = the alignment's ids; RECT -> file 0 / file 1;
    // STREAM -> ids1 = loaded ids, ids2 = the ids of the streamed records pushed so far (appended by the caller).
    TsvWriter(int fd, int threads);
    void write_header();
    void set_ids(const std::vector<std::string>* ids1, const std::vector<std::string>* ids2) { ids1_ = ids1; ids2_ = ids2; }
    void write_panel(const dg_panel& p);  // throws DistanceError on an io error other than BrokenPipe
    void flush();
    uint64_t lines() const { return lines_; }

private:
    void write_all(const char* p, size_t n);
    int fd_;
    int threads_;
    const std::vector<std::string>* ids1_ = nullptr;
    const std::vector<std::string>* ids2_ = nullptr;
    uint64_t lines_ = 0;
};

}  // namespace host
