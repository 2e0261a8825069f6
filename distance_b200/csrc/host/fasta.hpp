// fasta.hpp -- FASTA tokeniser + record validation for the `distance` host.
//
// Replaces fastaio.rs:174-286 (load_fasta / load_fastas / stream_fasta) and the part of rust-bio's
// bio::io::fasta::Reader (bio 1.6.0, Cargo.lock:216-217; source not in /root/reference) that those
// call.  rust-bio semantics restated from its published behaviour:
//   * a record starts at a line beginning with '>'; anything else where a header is expected is the
//     io error "Expected > at record start."
//   * id = header text after '>' up to the first whitespace, desc = the rest (trailing whitespace cut)
//   * seq = the following lines up to the next '>' line or EOF, each with trailing whitespace (incl.
//     '\r') removed and concatenated; blank lines inside a record are skipped by that rule
//   * iteration ends at EOF or at a completely empty record
// The reference's own tests only pin two-line, LF, upper-case records (fastaio.rs:344-350,
// lib.rs:906-914); multi-line / CRLF behaviour is "parity unpinned" (SURVEY.md 8c).
//
// Sequence bytes are validated against the table of encoding.rs:4-41 while they are copied, so the
// reference's error order is kept (fastaio.rs:111-113 before :188-190 when loading; :246-248 before
// :250-254 when streaming).  The bytes stay ASCII: the LUT itself runs on the device
// (DG_INPUT_ASCII), which also reproduces the stream-mode tn93 upper-case count quirk
// (fastaio.rs:139-142).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace host {

// Mirrors DistanceError (lib.rs:21-39) as far as the process exit text goes: main prints
// `Error: <Debug of the variant>` and exits 1 (main.rs:4-15).
struct DistanceError : std::runtime_error {
    explicit DistanceError(const std::string& debug_text) : std::runtime_error(debug_text) {}
};
DistanceError message_error(const std::string& msg);            // Message("...")
DistanceError io_error_custom(const std::string& msg);          // IOError(Custom { kind: Other, error: "..." })
DistanceError io_error_os(int err);                             // IOError(Os { code, kind, message })

struct Alignment {
    std::vector<std::string> ids;
    uint64_t width = 0;
    uint64_t n() const { return ids.size(); }
    const uint8_t* data() const { return fast_ ? fast_.get() : seqs.data(); }   // n x width raw ASCII bytes (validated)

    std::vector<uint8_t> seqs;            // filled by the sequential reader ...
    struct FreeDeleter { void operator()(uint8_t* p) const { std::free(p); } };
    std::unique_ptr<uint8_t, FreeDeleter> fast_;   // ... or by the parallel loader (uninitialised, huge-page backed when possible)
};

class FastaReader {
public:
    // validate = false: bytes are copied unchecked (stream mode checks the width first, fastaio.rs:246-254)
    explicit FastaReader(int fd, bool validate = true);
    // the same over bytes already in memory
    FastaReader(const char* data, size_t size, bool validate = true);
    // Reads the next record: id into `id`, sequence bytes APPENDED to `seq`.  Returns false at the
    // end of the input.  Throws DistanceError on malformed input or an invalid nucleotide.
    bool next(std::string& id, std::vector<uint8_t>& seq);
    // every input byte has been consumed (false after next() stopped early at an empty record)
    bool at_end() const { return eof_ && pos_ == end_ && !have_line_; }

private:
    bool read_line();      // fills line_ (without the '\n'); false at EOF with nothing read; throws on invalid UTF-8
    bool read_line_raw();
    bool fill();
    int fd_;
    const char* mem_ = nullptr;   // in-memory source (fd_ < 0)
    size_t mem_size_ = 0, mem_pos_ = 0;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
    std::string line_;
    bool have_line_ = false;
    bool validate_ = true;
};

// load_fasta (fastaio.rs:174-200).  The whole input is taken into memory (mmap for regular files) and parsed by
// `threads` workers (0 = all cores), each on a run of whole records: record boundaries are found, every record's
// width is checked and its bytes are validated and copied straight to n x width.  Anything but a clean alignment
// (a width mismatch, an invalid byte, a malformed or empty record ...) abandons the parallel pass and re-reads the
// same bytes with the sequential FastaReader, which raises the reference's error for the first offence in file order.
Alignment load_fasta(int fd, int threads = 0);
// stream_fasta (fastaio.rs:215-286) in blocks: reads the streamed file a batch at a time (parallel pread for regular
// files), cuts the block at the last record start, and parses its whole records with the same parallel pass as
// load_fasta, writing the sequence bytes straight into the caller's batch buffer (the pinned staging buffer of
// dg_stream_buffer).  A block that is not clean (width mismatch, invalid byte, malformed record) is re-read by the
// sequential FastaReader with the reference's order of checks, so the first offence in file order raises its error.
class StreamBlockParser {
public:
    StreamBlockParser(int fd, uint64_t width, int threads = 0);
    // Up to max_records whole records into dst (max_records x width bytes); their ids are appended.  0 = end of input.
    uint64_t next(uint8_t* dst, uint64_t max_records, std::vector<std::string>& ids);

private:
    void fill(size_t want);
    int fd_;
    uint64_t width_;
    int threads_;
    bool regular_ = false, eof_ = false, done_ = false;
    uint64_t offset_ = 0, file_size_ = 0;
    std::unique_ptr<char[]> raw_;   // uninitialised storage: carry-over of the previous block + the new bytes
    size_t raw_cap_ = 0, raw_len_ = 0;
};

// the cross-file width check of load_fastas (fastaio.rs:202-212)
void check_same_width(const Alignment& a, const Alignment& b);

bool valid_nucleotide(uint8_t c);
// Throws the Message error of fastaio.rs:89-91 for the first invalid byte of seq[0..len), if any.
void validate_record(const std::string& id, const uint8_t* seq, uint64_t len);

}  // namespace host
