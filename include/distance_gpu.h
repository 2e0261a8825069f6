/*
 * distance_gpu.h -- C ABI of the B200 (sm_100a) pairwise-comparison engine.
 *
 * This is the drop-in boundary for the hot path of benjamincjackson/distance
 * v0.3.1.  The reference has no FFI of its own: its only seam is the per-pair
 * Rust fn pointer returned by get_distance_function (src/lib.rs:477-488) and
 * called at src/lib.rs:434 (load) and src/lib.rs:325 (stream).  A per-pair seam
 * is useless for a GPU, so the boundary sits one level up and replaces what
 * lies between "records are encoded" (src/lib.rs:217-241, src/fastaio.rs:250-254)
 * and "ordered Distances reach gather_write" (src/lib.rs:612-637), i.e. the
 * bodies of load() (src/lib.rs:367-474) and stream() (src/lib.rs:269-365).
 * INTEGRATION.md shows the Rust `extern "C"` block and build.rs a maintainer
 * would add.
 *
 * Conventions
 *   - plain C: opaque context, pointers + sizes, no C++ types, no exceptions.
 *   - every call returns int: DG_OK (0) or a negative DG_ERR_*; the text is in
 *     dg_last_error().  Nothing aborts the process.
 *   - the caller owns every input buffer; the callee has copied what it needs
 *     by the time a call returns.  Result panels live in callee-owned pinned
 *     memory and are valid only until the sink callback returns.
 *   - a context is driven from ONE host thread.  Sink callbacks run on that
 *     thread, serially, in the reference's output order, so the writer needs
 *     no reorder map (the reference reorders by batch idx, src/lib.rs:616-637).
 *   - results never depend on the reference's -t / -b (pinned by its tests
 *     src/lib.rs:947-999, 1026-1066, 1092-1132).
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with DG_ERR_CUDA.
 *
 * Host guarantees before calling (the reference checks them while parsing):
 * all sequences of all inputs have the same width (src/fastaio.rs:188-190,
 * 206-208, 246-248) and every input holds >= 1 record (src/fastaio.rs:195-197,
 * 281-283).  Paradis input bytes must be one of the 17 codes of
 * src/encoding.rs:7-38 (checked on device: DG_ERR_INVALID_CODE).
 */
#ifndef DISTANCE_GPU_H
#define DISTANCE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_ABI_VERSION 2

#if defined(__GNUC__)
#define DG_API __attribute__((visibility("default")))
#else
#define DG_API
#endif

/* status codes */
#define DG_OK 0
#define DG_ERR_INVALID_ARG (-1)
#define DG_ERR_CUDA (-2)
#define DG_ERR_STATE (-3)
#define DG_ERR_INVALID_CODE (-4) /* invalid nucleotide byte; see dg_invalid_site */
#define DG_ERR_SINK (-5)         /* the sink callback returned non-zero */
#define DG_ERR_NOMEM (-6)

/* -m values, in the order of the CLI's value_parser (src/lib.rs:108) and the
 * dispatch of get_distance_function (src/lib.rs:477-488). */
typedef enum {
    DG_MEASURE_N = 0,      /* snp_consensus, src/measures.rs:28-53 (== n_high for every input) */
    DG_MEASURE_N_HIGH = 1, /* snp,           src/measures.rs:14-23 */
    DG_MEASURE_RAW = 2,    /* raw,           src/measures.rs:56-69 */
    DG_MEASURE_JC69 = 3,   /* jc69,          src/measures.rs:72-77 */
    DG_MEASURE_K80 = 4,    /* k80,           src/measures.rs:80-113 */
    DG_MEASURE_TN93 = 5    /* tn93,          src/measures.rs:116-193 */
} dg_measure;

/* What the `codes` buffers hold. */
typedef enum {
    DG_INPUT_PARADIS = 0, /* EncodedFastaRecord.seq bytes (src/fastaio.rs:16) */
    DG_INPUT_ASCII = 1,   /* raw FASTA sequence bytes; the LUT of src/encoding.rs:4-41 runs on device */
    DG_INPUT_NIBBLE = 2   /* ABI 2: two sites per byte (site 2k in the low nibble of byte k, site 2k+1 in the high one), a row is
                             (width + 1) / 2 bytes.  A nibble = the possibility bits of the Paradis code (its high nibble:
                             A = 8, G = 4, C = 2, T = 1, R = 12 ... N = 15); "known" is implied by a single bit, and N, '-' and
                             '?' (which no measure tells apart: src/measures.rs:17, 60-62, 158) all travel as 15.  Nibble 0 is
                             invalid.  Half the bytes over PCIe for a parser that packs while it validates. */
} dg_input_kind;

typedef enum {
    DG_RESULT_U32 = 0, /* n / n_high: the count; FloatInt::Int, printed `{}`   (src/lib.rs:626-628) */
    DG_RESULT_F64 = 1, /* raw/jc69/k80/tn93: IEEE bits; FloatInt::Float `{:.12}` (src/lib.rs:630-632) */
    DG_RESULT_U16 = 2, /* n / n_high with DG_OPT_RESULT_U16: the same count in 16 bits (halves the D2H bytes) */
    DG_RESULT_U8 = 3   /* ABI 2, n / n_high with DG_OPT_RESULT_U8: the count in 8 bits; 255 = "255 or more: look the pair up in
                          dg_panel.overflow" (closely related genomes differ at a few dozen sites: a quarter of the uint32 bytes) */
} dg_result_kind;

/* DG_RESULT_COUNTS16 (= 4, ABI 2; DG_OPT_RESULT_COUNTS): raw / jc69 / k80 / tn93 panels of resident runs and sessions hold,
 * per pair, the four uint16 counts the distance is a function of instead of the f64 itself (same 8 bytes):
 *   raw, jc69 : { n = #different, same, 0, 0 }                       d = n + same         (src/measures.rs:56-77)
 *   k80       : { same, ts + tv, tv, 0 }                             L = same + ts + tv   (src/measures.rs:80-113)
 *   tn93      : { count_L, count_d, count_P1, count_P2 }                                  (src/measures.rs:156-175)
 * A host that evaluates measures.rs's f64 expressions on them with its own libm (Rust's f64::ln is the platform's log)
 * prints text that is byte-identical to the reference's; the device's fused f64 epilogue (CUDA log: <= 1 ulp off) stays
 * the default.  Streamed batches, the LOP3 engine and widths above 65,535 always deliver DG_RESULT_F64. */
#define DG_RESULT_COUNTS16 4

/* DG_RESULT_U8 panels: the pairs whose count does not fit 8 bits, in no particular order. */
typedef struct {
    uint32_t index; /* position in dg_panel.data */
    uint32_t value; /* the count */
} dg_overflow;

typedef enum {
    DG_MODE_SQUARE = 0, /* one alignment, pairs i<j row-major: generate_pairs_square    src/lib.rs:502-547 */
    DG_MODE_RECT = 1,   /* two alignments, i major j minor:   generate_pairs_rectangle src/lib.rs:551-596 */
    DG_MODE_STREAM = 2  /* streamed record major, loaded record minor: src/lib.rs:322-325 */
} dg_mode;

/* One contiguous slab of results in the reference's output order.
 *   SQUARE: major rows are sequence1 = record i of the alignment; row i holds the
 *           (n - 1 - i) results for j = i+1 .. n-1, rows concatenated.
 *   RECT:   major rows are records of input 0, each row = n_cols results over input 1.
 *   STREAM: major rows are STREAMED records (global index since dg_stream_begin), each row =
 *           n_cols results over the loaded records.  TSV columns are still
 *           id1 = loaded id, id2 = streamed id (src/lib.rs:326-331). */
typedef struct {
    int32_t mode;        /* dg_mode */
    int32_t result_kind; /* dg_result_kind */
    uint64_t row_begin;  /* first major row in this panel */
    uint64_t row_end;    /* one past the last major row */
    uint64_t n_cols;     /* RECT/STREAM: results per row; SQUARE: n (row i has n-1-i results) */
    uint64_t n_results;  /* total results in `data` */
    const void *data;    /* uint32_t / double / uint16_t / uint8_t [n_results] per result_kind; pinned host memory */
    const dg_overflow *overflow; /* ABI 2: DG_RESULT_U8 only, else NULL */
    uint64_t n_overflow;
} dg_panel;

/* Return 0 to continue, non-zero to abort the run (-> DG_ERR_SINK). */
typedef int (*dg_sink_fn)(void *user, const dg_panel *panel);

typedef struct dg_ctx dg_ctx;

/* run flags */
#define DG_RUN_DEVICE_ONLY 1u /* compute every panel but skip the D2H copy and the sink (kernel-only timing) */
#define DG_RUN_REPACK 2u      /* re-run pack_planes from the device-resident codes first (needs DG_OPT_KEEP_CODES) */

/* dg_set_option keys */
#define DG_OPT_PANEL_BYTES 1 /* target bytes of one result panel (default 256 MiB) */
#define DG_OPT_KEEP_CODES 2  /* keep the raw code bytes on the device after dg_load_resident (0/1) */
#define DG_OPT_TILE_VARIANT 3 /* tuning: 0 = default tile shape per measure family, >0 = alternatives */
#define DG_OPT_ENGINE 4      /* 0 = auto (per shape / ambiguity load), 1 = LOP3+POPC bit-plane tiles, 2 = tcgen05 kind::i8
                                GEMM (int8 planes, int32 accumulation), 3 = tcgen05 kind::mxf4 GEMM (the same planes as
                                E2M1 nibbles, unit scale factors, fp32 accumulation: exact for these integer sums, twice
                                the MAC rate).  Set it before dg_load_resident: it decides which operands are built. */
#define DG_OPT_RESULT_U16 5  /* 0/1: deliver n / n_high panels as uint16_t (DG_RESULT_U16).  A count never exceeds
                                the width, so this is lossless for width <= 65535 (else DG_ERR_INVALID_ARG). */

#define DG_OPT_RESULT_U8 9   /* 0/1: deliver n / n_high panels of dg_run_* and the dg_square_* / dg_rect_* sessions as uint8_t
                                (DG_RESULT_U8) with an overflow list for the counts >= 255.  A panel with more than 16,384 such
                                pairs arrives as DG_RESULT_U16 instead (check dg_panel.result_kind per panel).  Needs width <= 65535. */

#define DG_OPT_RESULT_COUNTS 10 /* 0/1: raw / jc69 / k80 / tn93 panels carry DG_RESULT_COUNTS16 tuples where possible (check
                                   dg_panel.result_kind per panel) */

#define DG_OPT_PIPE_PANELS 6 /* dg_square_* sessions: how many result panels (per part) the triangle is cut into
                                (default 8; smaller panels start the D2H stream earlier, larger ones fill whole rounds of the CTA pairs better) */
#define DG_OPT_PIPE_CHUNK_BYTES 7 /* dg_square_* sessions: target bytes of one upload chunk (0 = automatic:
                                     max(24 MiB, alignment bytes / 40)); chunks are whole multiples of 128 records */
#define DG_OPT_REPACK_OVERLAP 8 /* 0/1 (default 1): kernel-only square runs with DG_RUN_REPACK re-pack the operand planes
                                  chunk by chunk, overlapped with the tiles (panels then run in descending row order);
                                  0 = pack everything first (dg_timings.pack_ms then times the pack kernel alone) */

typedef struct {
    double pack_ms;       /* pack_planes kernels, CUDA events on the launching stream */
    double count_ms;      /* count-tile kernels (incl. fused epilogue), CUDA events */
    double h2d_ms;        /* host-side wall time spent in H2D copies of codes */
    double total_ms;      /* wall time of the last dg_run_* / stream session */
    double run_ms;        /* device time of the last dg_run_*: CUDA events on the compute stream from the
                             first enqueue to the last kernel's end (max over the context's devices) */
    uint64_t pack_launches;
    uint64_t count_launches;
    uint64_t pairs;       /* pairs computed by count kernels since the last reset */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t engine;      /* engine of the last run: 1 = LOP3+POPC bit-plane tiles, 2 = tcgen05 int8 GEMM, 3 = tcgen05 fp4 GEMM */
    double sm_mhz;        /* ABI 2: SM clock the tensor-engine launches since the last reset actually ran at (clock64 /
                             globaltimer around each launch, first device); 0 when no such launch ran.  nvidia-smi's
                             200 ms samples miss the few-ms dips of a power-capped burst; this does not. */
} dg_timings;

DG_API int dg_abi_version(void);
/* Number of visible CUDA devices, or a negative DG_ERR_*. */
DG_API int dg_device_count(void);

/* Create a context on the given CUDA devices (gpu_ids == NULL -> devices 0..n_gpus-1).
 * `width` = alignment width in sites; fixed for the life of the context. */
DG_API int dg_create(const int *gpu_ids, int n_gpus, int measure, uint64_t width, dg_ctx **out);
DG_API void dg_destroy(dg_ctx *ctx);
/* Message of the last failing call on this context (ctx == NULL: of the last failing
 * dg_create on this thread).  Never NULL. */
DG_API const char *dg_last_error(const dg_ctx *ctx);
DG_API int dg_set_option(dg_ctx *ctx, int key, int64_t value);

/* Upload alignment `which` (0 or 1) = Setup.loaded_fastas[which] (src/lib.rs:135, 217):
 * n x width bytes row-major.  The planes are packed on every device of the context.
 * acgt_counts: n x 4 in the order A,T,G,C (EncodedFastaRecord.count_A/T/G/C, src/fastaio.rs:17-20)
 * or NULL -> computed on device from the codes exactly like count_bases (src/fastaio.rs:53-66,
 * case-insensitive).  Only tn93 reads them. */
DG_API int dg_load_resident(dg_ctx *ctx, int which, const uint8_t *codes, uint64_t n, int input_kind,
                     const uint64_t *acgt_counts);

/* The same for codes that already sit in DEVICE memory of CUDA device `src_device` (e.g. gathered from the other
 * ranks over NVLink / NCCL, or produced by a GPU parser): every device of the context takes its replica with
 * cudaMemcpyPeerAsync, so no byte crosses PCIe.  The caller's work on d_codes must be complete (synchronised)
 * before the call; the buffer may be reused when the call returns. */
DG_API int dg_load_resident_device(dg_ctx *ctx, int which, const uint8_t *d_codes, int src_device, uint64_t n,
                                   int input_kind, const uint64_t *acgt_counts);

/* When a call failed with DG_ERR_INVALID_CODE: the first offending (record, site, byte), from which
 * the host formats "Invalid nucleotide character in record '<id>': '<c>'" (src/fastaio.rs:89-91). */
DG_API int dg_invalid_site(const dg_ctx *ctx, uint64_t *record, uint64_t *site, uint8_t *byte);

/* All-vs-all within alignment 0 (replaces load() with one input, src/lib.rs:390-399). */
DG_API int dg_run_square(dg_ctx *ctx, dg_sink_fn sink, void *user, uint32_t flags);
/* Alignment 0 x alignment 1 (replaces load() with two inputs, src/lib.rs:401-409). */
DG_API int dg_run_rect(dg_ctx *ctx, dg_sink_fn sink, void *user, uint32_t flags);
/* Multi-process sharding: run only the panels that dg_plan_parts assigns to `part` (mode SQUARE or RECT).
 * Panels are independent, so ranks need no collective; each rank's sink sees its own panels in
 * increasing order. */
DG_API int dg_run_part(dg_ctx *ctx, int mode, uint32_t part, uint32_t n_parts, dg_sink_fn sink, void *user,
                uint32_t flags);

/* The panel plan dg_run_square / dg_run_rect / dg_run_part follow, without touching a device (pure
 * host arithmetic; lets a multi-process launcher and CPU tests see the sharding).  mode SQUARE:
 * n_rows = n_cols = n.  Writes up to `cap` panels (row_begin, row_end, n_results) and returns the
 * total number of panels, or a negative DG_ERR_*.  Which part owns panel k: dg_plan_parts.  Assumes the default
 * result width (uint32 / double); with DG_OPT_RESULT_U16 use dg_plan_ctx. */
DG_API int64_t dg_plan_panels(int measure, int mode, uint64_t n_rows, uint64_t n_cols, uint64_t panel_bytes,
                              int tile_variant, uint64_t *row_begin, uint64_t *row_end, uint64_t *n_results,
                              uint64_t cap);

/* ABI 2: the part that owns each panel of a plan (n_results[k] = the panel's result count, as dg_plan_* report them):
 * largest panels first, each to the part with the least work so far -- deterministic, so every rank derives the same
 * shares.  (Round-robin dealing left the largest of 8 shares 12 - 20 % above the mean.) */
DG_API int dg_plan_parts(const uint64_t *n_results, uint64_t count, uint32_t n_parts, uint32_t *part_of);

/* The same plan for THIS context: its loaded alignments, panel bytes, tile variant and result width
 * (uint16 panels hold twice the rows of uint32 ones).  Returns the number of panels or a negative DG_ERR_*. */
DG_API int64_t dg_plan_ctx(dg_ctx *ctx, int mode, uint64_t *row_begin, uint64_t *row_end, uint64_t *n_results,
                           uint64_t cap);

/* Pipelined all-vs-all (the same work as dg_load_resident + dg_run_part on alignment 0, overlapped): the caller
 * hands the alignment over in chunks, HIGHEST records first.  Row i of the upper triangle (src/lib.rs:512-513)
 * only needs the records j > i, so as soon as the records [lo, n) are on the device every result panel whose rows
 * start at or above lo runs: the PCIe upload, packing + tiles and the D2H of finished panels overlap.
 *   - dg_square_begin plans the session: this part's panels (dealt by dg_plan_parts over the session's
 *     plan, as in dg_run_part) and the chunk sequence, which is the same for every part.
 *   - dg_square_next returns the record range [*lo, *hi) the next dg_square_push must deliver (*lo == *hi: done).
 *   - dg_square_push takes that chunk: `codes` points at the first byte of record lo; src_device < 0 = host memory
 *     (pinned memory makes the copy asynchronous), otherwise the CUDA device that holds it (peer copy, e.g. a chunk
 *     another rank uploaded and broadcast over NVLink).  ready_event: NULL, or a cudaEvent_t of this process that
 *     the copy must wait for (e.g. recorded after the collective that fills the buffer), so the caller never has to
 *     block the host on its own stream.  Copies are asynchronous: a pinned or device buffer must
 *     stay untouched until DG_SQUARE_LOOKAHEAD further pushes, or dg_square_end, have returned (so a staging ring of
 *     DG_SQUARE_LOOKAHEAD + 1 chunk buffers is enough); pageable host memory may be reused at once.
 *   - the sink runs inside push / end calls, serially, once per panel, in COMPLETION order (descending rows), not
 *     in the reference's output order: each dg_panel says which rows it holds, so a consumer that needs the order of
 *     src/lib.rs:616-637 places panels by row_begin (or uses dg_load_resident + dg_run_square, which deliver in order).
 *   - when dg_square_end returns, alignment 0 is resident exactly as after dg_load_resident.
 * acgt_counts as in dg_load_resident.  One device per context.  Any failing call closes the session.  An invalid
 * nucleotide byte fails the push / end call that notices it (DG_ERR_INVALID_CODE) before any panel that depends on
 * it is delivered; dg_invalid_site then names the lowest (record, site) among the chunks seen so far, which need
 * not be the first in file order. */
#define DG_SQUARE_LOOKAHEAD 8
DG_API int dg_square_begin(dg_ctx *ctx, uint64_t n, int input_kind, const uint64_t *acgt_counts, uint32_t part,
                           uint32_t n_parts, dg_sink_fn sink, void *user);
DG_API int dg_square_next(dg_ctx *ctx, uint64_t *lo, uint64_t *hi);
/* The whole chunk sequence of the open session (push order): writes up to `cap` ranges, returns the number of chunks
 * or a negative DG_ERR_*.  Lets a launcher enqueue every upload / collective before the first push. */
DG_API int64_t dg_square_plan(dg_ctx *ctx, uint64_t *lo, uint64_t *hi, uint64_t cap);
DG_API int dg_square_push(dg_ctx *ctx, const uint8_t *codes, int src_device, uint64_t lo, uint64_t hi,
                          void *ready_event);
DG_API int dg_square_end(dg_ctx *ctx);
/* The whole session for an alignment in host memory: begin, push every chunk from `codes` (n x width), end. */
DG_API int dg_run_square_host(dg_ctx *ctx, const uint8_t *codes, uint64_t n, int input_kind,
                              const uint64_t *acgt_counts, uint32_t part, uint32_t n_parts, dg_sink_fn sink, void *user);

/* The two-file counterpart (replaces load() with two inputs, src/lib.rs:401-409, like dg_run_rect): alignment 1 is
 * resident (dg_load_resident(ctx, 1, ...)), alignment 0 arrives in chunks, LOWEST records first, through the same
 * dg_square_next / dg_square_plan / dg_square_push / dg_square_end calls.  A panel (rows of alignment 0 x all of
 * alignment 1) runs as soon as its rows have landed, so panels complete in ASCENDING order: the sink sees the
 * reference's output order (generate_pairs_rectangle, src/lib.rs:551-596) while the upload is still going on.
 * When dg_square_end returns, alignment 0 is resident as after dg_load_resident(ctx, 0, ...). */
DG_API int dg_rect_begin(dg_ctx *ctx, uint64_t n, int input_kind, const uint64_t *acgt_counts, uint32_t part,
                         uint32_t n_parts, dg_sink_fn sink, void *user);
DG_API int dg_run_rect_host(dg_ctx *ctx, const uint8_t *codes, uint64_t n, int input_kind,
                            const uint64_t *acgt_counts, uint32_t part, uint32_t n_parts, dg_sink_fn sink, void *user);

/* -s streaming (replaces stream(), src/lib.rs:269-365): alignment 0 is resident, batches of the
 * streamed alignment are pushed in file order.  Batches are staged through double-buffered pinned
 * memory and copied with cudaMemcpyAsync while earlier batches compute.  The sink may run during
 * any push/end call, always in streamed-record order. */
DG_API int dg_stream_begin(dg_ctx *ctx, dg_sink_fn sink, void *user, uint64_t max_batch);
DG_API int dg_stream_push(dg_ctx *ctx, const uint8_t *codes, uint64_t n_batch, int input_kind,
                   const uint64_t *acgt_counts);
/* Zero-copy producer: the pinned staging buffer the NEXT dg_stream_push will use (capacity in records =
 * the session's batch size).  A parser that writes its records straight into *buf and then calls
 * dg_stream_push(ctx, *buf, n <= capacity, ...) skips the host-side staging copy.  The pointer is valid
 * until that push; the sink may run inside this call. */
DG_API int dg_stream_buffer(dg_ctx *ctx, uint8_t **buf, uint64_t *capacity_records);
DG_API int dg_stream_end(dg_ctx *ctx);

/* Debug / parity: raw integer counts of every pair (a in alignment which_a, b in which_b), from
 * the same tiled kernels with the f64 epilogue replaced by a store.  out = n_a x n_b x 4 uint32:
 *   n, n_high : {diff, 0, 0, 0}
 *   raw, jc69 : {diff (n), same, 0, 0}            d = same + diff      (src/measures.rs:59-66)
 *   k80       : {same, ts + tv, tv, 0}            L = same + ts + tv   (src/measures.rs:85-107)
 *   tn93      : {count_L, count_d, P1, P2}                             (src/measures.rs:156-175) */
DG_API int dg_debug_counts(dg_ctx *ctx, int which_a, int which_b, uint32_t *out);
/* Debug / parity: packed planes of alignment `which` as seen by device 0.
 * core/aux = n x words x 4 uint32 ({pA,pG,pC,pT} / {K,C,M,0} per 32-site word),
 * acgt = n x 4 (A,T,G,C).  Any pointer may be NULL. *words_out = 32-bit words per sequence. */
DG_API int dg_debug_planes(dg_ctx *ctx, int which, uint32_t *core, uint32_t *aux, uint64_t *acgt,
                    uint64_t *words_out);

DG_API int dg_get_timings(const dg_ctx *ctx, dg_timings *out);
DG_API int dg_reset_timings(dg_ctx *ctx);

/* Page-locked host memory for callers that want zero-copy staging of their inputs. */
DG_API void *dg_alloc_pinned(size_t bytes);
DG_API void dg_free_pinned(void *p);

#ifdef __cplusplus
}
#endif
#endif /* DISTANCE_GPU_H */
