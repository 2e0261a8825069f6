// dg_api.cu -- host side of libdistance_gpu: context, plane residency, the panel / tile scheduler
// that replaces the reference's thread + batch scheduler (lib.rs:269-596), -s streaming through
// double-buffered pinned batches, and the extern "C" entry points of include/distance_gpu.h.
//
// No CPU fallback lives here: every compute entry point needs a CUDA device.
#include "../../include/distance_gpu.h"
#include "kernels.cuh"
#include "tc_engine.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace {

using namespace dg;

thread_local std::string g_create_error;

double wall_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct DgError {
    int code;
    std::string msg;
};

[[noreturn]] void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw DgError{code, buf};
}

#define CUDA_CHECK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess)                                                             \
            fail(e_ == cudaErrorMemoryAllocation ? DG_ERR_NOMEM : DG_ERR_CUDA,             \
                 "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,   \
                 cudaGetErrorString(e_));                                                  \
    } while (0)

// ---- tile shapes -------------------------------------------------------------------------------
struct TileShape { int tm, tn; };

TileShape tile_shape(int fam, int variant) {
    switch (fam) {
    case FAM_SNP:  return variant == 1 ? TileShape{128, 128} : TileShape{128, 64};
    case FAM_RAW:  return variant == 1 ? TileShape{64, 64} : TileShape{64, 128};
    default:       return TileShape{64, 64};
    }
}

template <int FAM, int RM, int RN, bool COUNTS, int MINB>
void launch_tile(const CountParams& p, dim3 grid, cudaStream_t s) {
    using Cfg = TileCfg<FAM, RM, RN>;
    auto kern = count_tile_kernel<FAM, RM, RN, COUNTS, MINB>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(p);
    CUDA_CHECK(cudaGetLastError());
}

template <bool COUNTS>
void launch_count(int fam, int variant, const CountParams& p, dim3 grid, cudaStream_t s) {
    switch (fam) {
    case FAM_SNP:
        if (variant == 1) launch_tile<FAM_SNP, 8, 8, COUNTS, 2>(p, grid, s);
        else launch_tile<FAM_SNP, 8, 4, COUNTS, 2>(p, grid, s);  // measured faster on B200 (profiles/)
        break;
    case FAM_RAW:
        if (variant == 1) launch_tile<FAM_RAW, 4, 4, COUNTS, 2>(p, grid, s);
        else launch_tile<FAM_RAW, 4, 8, COUNTS, 2>(p, grid, s);
        break;
    case FAM_K80: launch_tile<FAM_K80, 4, 4, COUNTS, 2>(p, grid, s); break;
    default: launch_tile<FAM_TN93, 4, 4, COUNTS, 2>(p, grid, s); break;
    }
}

// ---- per-device state --------------------------------------------------------------------------
struct PlaneSet {  // one alignment packed on one device
    uint64_t n = 0, n_pad = 0;
    uint4* core = nullptr;
    uint4* aux = nullptr;
    uint32_t* acgt = nullptr;
    uint8_t* codes = nullptr;  // kept only with DG_OPT_KEEP_CODES
    uint8_t* nib = nullptr;    // DG_INPUT_NIBBLE: the uploaded nibble rows (unpacked into `codes` on the device)
    uint64_t nib_cap = 0;      // bytes
    int input_kind = 0;
    bool acgt_from_host = false;
    uint64_t cap_pad = 0;     // resident sets: record capacity of codes / acgt / tc_ops (buffers are reused across loads)
    bool lop3_ready = false;  // core / aux hold the current alignment (built lazily: only when the LOP3 engine runs)
    bool tc_ready = false;    // tc_ops / pp hold the current alignment
    bool pp_stale = false;    // loaded by a dg_square_* session: pp (the per-site index) is rebuilt on first use
    // tcgen05 engine (DG_OPT_ENGINE = 2): int8 one-hot operand planes, N-like counts, partial-code index
    int8_t* tc_ops = nullptr;
    bool tc_fp4 = false;      // tc_ops hold E2M1 nibbles (engine 3) instead of int8 (engine 2)
    uint64_t tc_wp8 = 0;
    tc::PpIndex pp;
    CUtensorMap map_a{}, map_b{};  // box 128 rows / box 256 rows over tc_ops
};

struct Slot {  // one stage of the result ring (and of the stream-input ring)
    void* d_out = nullptr;
    void* h_out = nullptr;
    int* d_scratch = nullptr;       // tcgen05 engine: raw int32 sums, [accumulator][panel pair]
    size_t scratch_cap = 0;
    int kind = -1;                  // dg_result_kind of the panel in flight when it differs from the context's default (-1)
    uint8_t* d_out8 = nullptr;      // DG_OPT_RESULT_U8: the panel narrowed to uint8 ...
    size_t out8_cap = 0;
    uint32_t* d_ovf = nullptr;      // ... and {count, -, entries {index, value} x OVF_CAP}
    uint32_t* h_ovf = nullptr;      // pinned mirror
    uint32_t* d_tiles = nullptr;    // tcgen05 engine, square panels: live-tile list of this slot's launch
    uint32_t* h_tiles = nullptr;    // pinned staging of the same
    size_t tiles_cap = 0;
    cudaEvent_t k_start = nullptr, k_stop = nullptr, copied = nullptr, in_ready = nullptr, gemm_done = nullptr;
    // stream mode staging
    uint8_t* h_in = nullptr;
    uint8_t* d_in = nullptr;
    uint8_t* d_nib = nullptr;       // DG_INPUT_NIBBLE batches land here and are unpacked into d_in
    uint32_t* h_acgt = nullptr;
    PlaneSet batch;
    cudaEvent_t p_start = nullptr, p_stop = nullptr;
};

struct Device {
    int id = 0;
    // compute = stream of ring slot 0 (also packs / resident loads), compute2 = stream of ring slot 1:
    // consecutive panels alternate streams so the tail wave of one launch overlaps the next launch.
    cudaStream_t compute = nullptr, compute2 = nullptr, copy = nullptr, copy_in = nullptr;
    cudaStream_t cs(int slot) const { return (slot & 1) ? compute2 : compute; }
    cudaStream_t post = nullptr;   // repair + combine passes of resident panels: behind their GEMM through an event, so the next
                                   // panel's GEMM (the other compute stream) never queues behind an f64 pass
    PlaneSet set[2];
    static constexpr int NSLOT = 3;   // result ring of resident panels (stream sessions use the first two slots)
    Slot slot[NSLOT];
    size_t out_cap = 0;   // bytes of d_out / h_out
    size_t in_cap = 0;    // records of stream staging
    cudaEvent_t run_start = nullptr, run_stop = nullptr;
    unsigned long long* clk_probe = nullptr;   // DG_CLOCK_PROBE: SM clocks / ns accumulated by the GEMM launches (debug)
    // load-path scratch, allocated once (no cudaMalloc / cudaFree on the per-load path)
    uint32_t* pp_cnt = nullptr;     // [width] partial codes per site
    uint32_t* pp_cursor = nullptr;  // [width]
    double* pp_work = nullptr;
    uint64_t* pp_hits = nullptr;    // partial codes met by pack_ops_kernel (unsorted entries)
    uint32_t pp_hit_cap = 0;
    uint32_t* pp_hit_count = nullptr;
    uint32_t* h_pp_total = nullptr; // pinned {entries}
    double* h_pp_work = nullptr;    // pinned
    cudaEvent_t chunk_ev[32] = {};
    cudaStream_t unpack = nullptr;     // sessions fed with nibble rows: the unpacked copy is written here, off the critical path
    cudaEvent_t unpack_ev[32] = {};    // chunk g's unpacked bytes are in place
    unsigned long long* d_invalid = nullptr;  // [3]: ring slot 0, ring slot 1, resident loads
    unsigned long long* h_invalid = nullptr;  // [3] pinned mirror
    // pipelined all-vs-all session (dg_square_*): a deeper result ring of small panels + the chunked index
    static constexpr int NPS = 4;
    Slot pslot[NPS];
    size_t pout_cap = 0;
    uint32_t* sq_off = nullptr;      // [chunks][width + 1] offsets into sq_entries
    size_t sq_off_chunks = 0;
    uint32_t* sq_cum = nullptr;      // [width] partial codes per site over every chunk so far
    uint32_t* sq_total = nullptr;    // [1] running entry count
    uint64_t* sq_entries = nullptr;
    uint32_t sq_entries_cap = 0;
    std::vector<void*> sq_retired;   // outgrown entry buffers, freed when the session ends
    uint32_t* h_sq_total = nullptr;  // pinned [chunks]
    double* h_sq_work = nullptr;     // pinned [chunks]
    unsigned long long* h_sq_invalid = nullptr;  // pinned [chunks]
    std::vector<cudaEvent_t> sq_ev;  // per chunk: index scanned (or just packed)
    cudaEvent_t sq_ready = nullptr;  // entries of the chunks pumped so far are in place
    cudaStream_t prep = nullptr;     // packing + index scan of a session: highest priority, so it slips in between
                                     // the persistent tile launches instead of queueing behind them
    cudaStream_t fill = nullptr;     // entry scatter of pumped chunks (not behind the packing of later chunks)
    // DG_TRACE: device timeline of a session (timing events, ms relative to tr_base)
    cudaEvent_t tr_base = nullptr;
    std::vector<cudaEvent_t> tr_copy, tr_prep;             // per chunk: copy landed, pack + scan done
    std::vector<cudaEvent_t> tr_k0, tr_k1, tr_d2h;         // per launched panel
};

struct Panel {
    uint64_t row0, row1;   // major rows
    uint64_t out_base;     // global index of the first result
    uint64_t n_results;
};

struct InFlight {  // a panel / batch whose results have not been handed to the sink yet
    int dev = 0, slot = 0;
    dg_panel desc{};
    bool pack_timed = false;
    uint64_t pairs = 0;
};

}  // namespace

struct dg_ctx {
    int measure = 0, fam = 0;
    uint64_t width = 0;
    uint32_t wp = 0;
    std::vector<Device> devs;
    std::string err;
    // options
    size_t panel_bytes = (size_t)256 << 20;
    bool keep_codes = false;
    int tile_variant = 0;
    int engine = 0;       // DG_OPT_ENGINE
    int last_engine = 0;  // engine the last run used (1 LOP3, 2 tcgen05 int8, 3 tcgen05 fp4)
    // Exactness limits of the tensor engines (per site a raw sum moves by at most 4: tc_engine.cuh header):
    //   kind::mxf4 accumulates in fp32: integer sums are exact while 4 * (padded width) < 2^24;
    //   kind::i8 accumulates in int32 (4 * width < 2^31); the partial-code index packs the site in 28 bits and the TMA
    //   x coordinate (plane * bytes per plane) is an int, so the int8 engine takes widths below 2^27.
    // auto falls through fp4 -> int8 -> LOP3+POPC tiles (uint32 counts, any width below 2^31).
    bool fp4_exact() const { return width <= (1ull << 22) - 256; }
    bool i8_ok() const { return width < (1ull << 27); }
    bool tc_wanted() const { return engine >= 2 || (engine == 0 && i8_ok()); }
    bool want_fp4() const { return engine == 3 || (engine == 0 && fp4_exact()); }   // operand format of the tensor engine
    // invalid-site report
    bool have_invalid = false;
    uint64_t inv_record = 0, inv_site = 0;
    uint8_t inv_byte = 0;
    dg_timings tm{};
    double clk_cycles = 0, clk_ns = 0;   // harvest_clock
    // stream session
    bool streaming = false;
    bool s_tc = false;  // this stream session runs its batches on the tcgen05 engine
    dg_sink_fn s_sink = nullptr;
    void* s_user = nullptr;
    uint64_t s_max_batch = 0, s_rows_pushed = 0, s_batches = 0;
    std::vector<InFlight> s_queue;  // FIFO of batches not yet sunk
    double s_t0 = 0;

    // pipelined all-vs-all session (dg_square_*)
    struct SqChunk { uint64_t lo, hi; };
    bool sq_open = false;
    bool sq_tc = false, sq_fallback = false, sq_needs_pp = false;
    int sq_mode = DG_MODE_SQUARE;      // SQUARE (chunks and panels descending) or RECT (alignment 0 pushed ascending against
                                       // the resident alignment 1: panels ascending = the reference's output order)
    std::vector<Panel> sq_panels;      // this part's panels in LAUNCH order
    size_t sq_next = 0;                // next panel to launch
    std::vector<SqChunk> sq_chunks;    // descending record ranges, pushed in this order
    size_t sq_pushed = 0, sq_pumped = 0;
    std::vector<uint32_t> sq_base;     // entries before chunk g (size chunks + 1), known once chunk g-1 is pumped
    uint64_t sq_n = 0, sq_launched = 0;
    int sq_input_kind = 0;             // what the pack kernels see (nibble input is unpacked to Paradis bytes first)
    bool sq_nibble = false;            // the session's chunks arrive as DG_INPUT_NIBBLE rows
    dg_sink_fn sq_sink = nullptr;
    void* sq_user = nullptr;
    std::vector<InFlight> sq_queue;
    double sq_t0 = 0;
    int pipe_panels = 8;               // DG_OPT_PIPE_PANELS
    uint64_t pipe_chunk_bytes = 0;     // DG_OPT_PIPE_CHUNK_BYTES (0 = automatic)
    bool repack_overlap = true;        // DG_OPT_REPACK_OVERLAP
    bool sq_trace = false;
    std::vector<int> sq_trace_panel;   // panel index of the n-th launch

    // auto engine: both-partial repairs (one atomic each) the tensor engine may spend per pair before the LOP3 tiles (no
    // repair, but ~6x the time per pair-site) are the better choice; the tiles' time grows with the width, so does the budget
    double pp_budget() const { return 2.0 * std::max(1.0, (double)width / 29903.0); }
    // panel planner inputs: tile width of the tensor engine this context would run and work items (accumulators) per tile
    uint64_t plan_tn() const { return want_fp4() ? 240 : 256; }
    uint64_t plan_items() const { static const uint64_t it[4] = {1, 2, 3, 5}; return it[fam]; }
    bool result_u16 = false;  // DG_OPT_RESULT_U16: n / n_high panels hold uint16 counts (needs width <= 65535)
    bool result_u8 = false;   // DG_OPT_RESULT_U8: ... delivered as uint8 + overflow list (the device still computes uint16)
    bool u16() const { return measure <= 1 && (result_u16 || result_u8); }
    bool u8() const { return measure <= 1 && result_u8; }
    size_t elem_bytes() const { return measure <= 1 ? (u16() ? 2 : 4) : 8; }
    bool result_counts = false;   // DG_OPT_RESULT_COUNTS
    // float measures on the tensor engines (resident panels / sessions): count tuples instead of the f64 (same 8 bytes)
    bool counts16() const { return measure >= 2 && result_counts && width <= 65535; }
    int result_kind() const { return measure <= 1 ? (u16() ? DG_RESULT_U16 : DG_RESULT_U32) : DG_RESULT_F64; }
};

namespace {

void free_set(PlaneSet& s) {
    if (s.tc_ops) cudaFree(s.tc_ops);
    if (s.pp.entries) cudaFree(s.pp.entries);
    if (s.pp.site_off) cudaFree(s.pp.site_off);
    if (s.core) cudaFree(s.core);
    if (s.aux) cudaFree(s.aux);
    if (s.acgt) cudaFree(s.acgt);
    if (s.codes) cudaFree(s.codes);
    if (s.nib) cudaFree(s.nib);
    s = PlaneSet{};
}

void alloc_set(dg_ctx* c, PlaneSet& s, uint64_t n, bool with_codes, bool lop3_planes = true) {
    s.n = n;
    s.n_pad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    const size_t plane_bytes = (size_t)s.n_pad * c->wp * sizeof(uint4);
    if (lop3_planes) CUDA_CHECK(cudaMalloc(&s.core, plane_bytes));
    if (lop3_planes && c->fam != FAM_SNP) CUDA_CHECK(cudaMalloc(&s.aux, plane_bytes));
    CUDA_CHECK(cudaMalloc(&s.acgt, (size_t)s.n_pad * 4 * sizeof(uint32_t)));
    if (with_codes) CUDA_CHECK(cudaMalloc(&s.codes, (size_t)std::max<uint64_t>(1, n * c->width)));
}

// Enqueue pack_planes for `n` records whose code bytes are at d_codes.
void enqueue_pack(dg_ctx* c, unsigned long long* d_inv, PlaneSet& s, const uint8_t* d_codes, uint64_t n,
                  int input_kind, bool count_on_device, bool upper_ascii, cudaStream_t st) {
    PackParams pp{};
    pp.codes = d_codes;
    pp.n = n;
    pp.n_pad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    pp.width = c->width;
    pp.wp = c->wp;
    pp.core = s.core;
    pp.aux = s.aux;
    pp.acgt = count_on_device ? s.acgt : nullptr;
    pp.count_upper_ascii = upper_ascii ? 1 : 0;
    pp.invalid = d_inv;
    if (count_on_device) CUDA_CHECK(cudaMemsetAsync(s.acgt, 0, (size_t)pp.n_pad * 4 * sizeof(uint32_t), st));
    const uint64_t units = pp.n_pad * (c->wp / 4);
    const uint64_t blocks = std::min<uint64_t>((units + 7) / 8, (uint64_t)148 * 16);
    if (input_kind == DG_INPUT_ASCII)
        pack_planes_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(pp);
    else
        pack_planes_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(pp);
    CUDA_CHECK(cudaGetLastError());
    c->tm.pack_launches++;
}

// After a sync: did pack_planes see an invalid byte?  `probe` fetches the byte for the report.
void check_invalid(dg_ctx* c, Device& d, const uint8_t* dev_codes, uint64_t row_offset) {
    CUDA_CHECK(cudaMemcpy(d.h_invalid + 2, d.d_invalid + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const unsigned long long key = d.h_invalid[2];
    if (key == ~0ull) return;
    const unsigned long long reset = ~0ull;
    CUDA_CHECK(cudaMemcpy(d.d_invalid + 2, &reset, sizeof reset, cudaMemcpyHostToDevice));
    c->have_invalid = true;
    c->inv_record = (key >> 32) + row_offset;
    c->inv_site = key & 0xffffffffull;
    c->inv_byte = 0;
    if (dev_codes) CUDA_CHECK(cudaMemcpy(&c->inv_byte, dev_codes + (key >> 32) * c->width + c->inv_site, 1, cudaMemcpyDeviceToHost));
    fail(DG_ERR_INVALID_CODE, "invalid nucleotide byte 0x%02x in record %llu at site %llu", c->inv_byte,
         (unsigned long long)c->inv_record, (unsigned long long)c->inv_site);
}

void ensure_out_ring(dg_ctx* c, Device& d, size_t bytes) {
    if (d.out_cap >= bytes) return;
    for (auto& s : d.slot) {
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_out) cudaFreeHost(s.h_out);
        s.d_out = s.h_out = nullptr;
    }
    d.out_cap = 0;
    try {
        for (auto& s : d.slot) {
            CUDA_CHECK(cudaMalloc(&s.d_out, bytes));
            CUDA_CHECK(cudaHostAlloc(&s.h_out, bytes, cudaHostAllocDefault));
        }
    } catch (const DgError& e) {
        // a panel is never smaller than 512 rows x every column (the tensor engine's block height), whatever
        // DG_OPT_PANEL_BYTES asks for: with millions of columns that minimum may not fit
        fail(e.code, "the result ring needs %d x %zu bytes of device and of page-locked host memory (one panel = at least 512 rows x "
                     "every column, DG_OPT_PANEL_BYTES = %zu): %s", Device::NSLOT, bytes, c->panel_bytes, e.msg.c_str());
    }
    d.out_cap = bytes;
}


// ---- DG_OPT_RESULT_U8 --------------------------------------------------------------------------------------------------
constexpr uint32_t OVF_CAP = 16384;                       // overflow entries a uint8 panel may carry
constexpr size_t OVF_WORDS = 2 + 2 * (size_t)OVF_CAP;     // {count, pad, entries}
void ensure_narrow(Slot& s, size_t n_results) {
    if (!s.d_ovf) {
        CUDA_CHECK(cudaMalloc(&s.d_ovf, OVF_WORDS * 4));
        CUDA_CHECK(cudaHostAlloc(&s.h_ovf, OVF_WORDS * 4, cudaHostAllocDefault));
    }
    if (s.out8_cap >= n_results) return;
    if (s.d_out8) cudaFree(s.d_out8);
    s.d_out8 = nullptr; s.out8_cap = 0;
    CUDA_CHECK(cudaMalloc(&s.d_out8, n_results + 64));
    s.out8_cap = n_results;
}
void free_narrow(Slot& s) {
    if (s.d_out8) cudaFree(s.d_out8);
    if (s.d_ovf) cudaFree(s.d_ovf);
    if (s.h_ovf) cudaFreeHost(s.h_ovf);
    s.d_out8 = nullptr; s.d_ovf = s.h_ovf = nullptr; s.out8_cap = 0;
}
// after the panel's kernels, on their stream: uint16 d_out -> uint8 d_out8 + overflow list
void enqueue_narrow(dg_ctx* c, Slot& s, uint64_t n_results, cudaStream_t st) {
    CUDA_CHECK(cudaMemsetAsync(s.d_ovf, 0, 8, st));
    const unsigned grid = (unsigned)std::min<uint64_t>((n_results / 16 + 256) / 256, 148 * 16);
    tc::narrow_u8_kernel<<<std::max(1u, grid), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(s.d_out), s.d_out8, n_results, s.d_ovf, OVF_CAP);
    CUDA_CHECK(cudaGetLastError());
    c->tm.count_launches++;
}
// D2H of a finished panel on the copy stream (after `k_stop`)
void enqueue_panel_d2h(dg_ctx* c, Device& d, Slot& s, uint64_t n_results) {
    if (c->u8()) {
        CUDA_CHECK(cudaMemcpyAsync(s.h_out, s.d_out8, (size_t)n_results, cudaMemcpyDeviceToHost, d.copy));
        CUDA_CHECK(cudaMemcpyAsync(s.h_ovf, s.d_ovf, OVF_WORDS * 4, cudaMemcpyDeviceToHost, d.copy));
    } else {
        CUDA_CHECK(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)n_results * c->elem_bytes(), cudaMemcpyDeviceToHost, d.copy));
    }
}
// fill the result fields of a panel descriptor once its copy has landed; a uint8 panel with too many overflows is
// fetched again as uint16 (the slot's d_out still holds it)
void finish_panel_desc(dg_ctx* c, Device& d, Slot& s, dg_panel& desc) {
    desc.result_kind = s.kind >= 0 ? s.kind : c->result_kind();
    desc.data = s.h_out;
    desc.overflow = nullptr; desc.n_overflow = 0;
    if (!c->u8()) { c->tm.d2h_bytes += desc.n_results * c->elem_bytes(); return; }
    const uint32_t count = s.h_ovf[0];
    if (count <= OVF_CAP) {
        desc.result_kind = DG_RESULT_U8;
        desc.overflow = reinterpret_cast<const dg_overflow*>(s.h_ovf + 2);
        desc.n_overflow = count;
        c->tm.d2h_bytes += desc.n_results + OVF_WORDS * 4;
    } else {
        CUDA_CHECK(cudaSetDevice(d.id));
        CUDA_CHECK(cudaMemcpy(s.h_out, s.d_out, (size_t)desc.n_results * 2, cudaMemcpyDeviceToHost));
        c->tm.d2h_bytes += desc.n_results * 3 + OVF_WORDS * 4;
    }
}

// ---- tcgen05 engine: host side ------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !ptr) fail(DG_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D map over ops[n_pad][nplanes * wp8] bytes: box = 128 bytes (one swizzle atom of K) x `box_rows` records.
void make_ops_map(CUtensorMap* map, void* base, uint64_t row_bytes, uint64_t rows, uint32_t box_rows) {
    const cuuint64_t dims[2] = {row_bytes, rows};
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)tc::KB, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(DG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
}

// Which int8 planes a family stores and which plane pairs each accumulator sums (see tc_engine.cuh).
struct TcSchedule {
    int nplanes;
    uint8_t plane_id[tc::MAX_PLANES];
    int nacc;
    int npairs[5];
    uint8_t pa[5][4], pb[5][4];
    bool needs_pp;  // the n/n_high sum needs the both-partial correction
};

template <int FAM>
bool planes_match(const TcSchedule& sch) {
    if (sch.nplanes != tc::PackPlanes<FAM>::N) return false;
    for (int i = 0; i < sch.nplanes; i++)
        if (sch.plane_id[i] != tc::PackPlanes<FAM>::id(i)) return false;
    return true;
}

const TcSchedule& tc_schedule(int fam) {
    using namespace tc;
    static const TcSchedule snp = {8, {P_UA, P_UG, P_UC, P_UT, P_VA, P_VG, P_VC, P_VT}, 1, {4},
                                   {{0, 1, 2, 3}}, {{4, 5, 6, 7}}, true};
    static const TcSchedule raw = {12, {P_UA, P_UG, P_UC, P_UT, P_VA, P_VG, P_VC, P_VT, P_KA, P_KG, P_KC, P_KT}, 2, {4, 4},
                                   {{0, 1, 2, 3}, {8, 9, 10, 11}}, {{4, 5, 6, 7}, {8, 9, 10, 11}}, true};
    static const TcSchedule k80 = {6, {P_PURK, P_PYRK, P_W, P_Z, P_PURC, P_PYRC}, 3, {2, 2, 2},
                                   {{0, 1}, {2, 3}, {4, 5}}, {{0, 1}, {2, 3}, {5, 4}}, false};
    static const TcSchedule tn93 = {5, {P_K, P_PURK, P_PYRK, P_W, P_Z}, 5, {1, 1, 1, 1, 1},
                                    {{0}, {1}, {2}, {3}, {4}}, {{0}, {1}, {2}, {3}, {4}}, false};
    // pack_ops_kernel stores the planes of tc::PackPlanes<FAM> (compile-time): the schedules must list the same ones
    static const bool ok = planes_match<FAM_SNP>(snp) && planes_match<FAM_RAW>(raw) && planes_match<FAM_K80>(k80) &&
                           planes_match<FAM_TN93>(tn93);
    if (!ok) fail(DG_ERR_STATE, "internal: tensor schedules and PackPlanes disagree");
    return fam == FAM_SNP ? snp : (fam == FAM_RAW ? raw : (fam == FAM_K80 ? k80 : tn93));
}
// first stored plane of the V operand (pp_correct_scan_kernel reads the partial codes back from it)
constexpr uint32_t TC_VPLANE0 = 4;

void alloc_tc_operands(dg_ctx* c, PlaneSet& s) {
    const TcSchedule& sch = tc_schedule(c->fam);
    s.tc_fp4 = c->want_fp4();
    // one 128-byte K block = 128 sites of int8 or 256 sites of E2M1 nibbles
    const uint64_t sites_per_block = s.tc_fp4 ? 2 * tc::KB : tc::KB;
    const uint64_t wp8 = (c->width + sites_per_block - 1) / sites_per_block * tc::KB;
    s.tc_wp8 = wp8;
    CUDA_CHECK(cudaMalloc(&s.tc_ops, (size_t)s.n_pad * sch.nplanes * wp8));
    const uint64_t row_bytes = (uint64_t)sch.nplanes * wp8;
    make_ops_map(&s.map_a, s.tc_ops, row_bytes, s.n_pad, tc::TM);
    make_ops_map(&s.map_b, s.tc_ops, row_bytes, s.n_pad, s.tc_fp4 ? tc::TN_FP4 / 2 : tc::TN);  // fp4: one CTA's half of B
}

// Enqueue (no sync) the int8 operand planes (and, with count_acgt, the per-record A,T,G,C counts) of rows
// [row0, row0 + n) of the set; the chunk's padding rows up to a multiple of 128 are zero-filled.
// b_side_only: store only the planes the B (column) operand of the family's schedule reads (records that are columns of
// this rank's panels but rows of nobody's here: a rank of a multi-process run packs U planes for its own rows only).
void enqueue_tc_pack(dg_ctx* c, PlaneSet& s, const uint8_t* d_codes, uint64_t n, int input_kind, bool count_acgt,
                     cudaStream_t st, uint64_t row0 = 0, unsigned long long* d_inv = nullptr, bool upper_ascii = false,
                     Device* pp_dev = nullptr, bool b_side_only = false, bool nibble_rows = false) {
    const TcSchedule& sch = tc_schedule(c->fam);
    const uint64_t n_pad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    tc::PackI8Params pp{};
    // nibble_rows: d_codes holds DG_INPUT_NIBBLE rows; the kernel expands them in registers (no unpacked copy is read)
    pp.nibble = nibble_rows ? 1 : 0;
    pp.in_stride = nibble_rows ? (c->width + 1) / 2 : c->width;
    pp.codes = d_codes + row0 * pp.in_stride; pp.n = n; pp.n_pad = n_pad; pp.width = c->width; pp.wp8 = s.tc_wp8;
    pp.ops = s.tc_ops + (size_t)row0 * sch.nplanes * s.tc_wp8;
    pp.acgt = count_acgt ? s.acgt + row0 * 4 : nullptr;
    pp.count_upper_ascii = upper_ascii && input_kind == DG_INPUT_ASCII ? 1 : 0;
    pp.invalid = d_inv; pp.seq0 = row0;
    if (pp_dev) {   // also collect the partial ambiguity codes (per-site counts + unsorted entries) for the repair index
        pp.pp_site_cnt = pp_dev->pp_cnt; pp.pp_hits = pp_dev->pp_hits;
        pp.pp_hit_count = pp_dev->pp_hit_count; pp.pp_hit_cap = pp_dev->pp_hit_cap;
    }
    pp.ascii = input_kind == DG_INPUT_ASCII;
    pp.nplanes = sch.nplanes;
    for (int i = 0; i < sch.nplanes; i++) pp.plane_id[i] = sch.plane_id[i];
    pp.plane_mask = 0xFFFFFFFFu;
    if (b_side_only) {
        pp.plane_mask = 0;
        for (int a = 0; a < sch.nacc; a++)
            for (int i = 0; i < sch.npairs[a]; i++) pp.plane_mask |= 1u << sch.pb[a][i];
    }
    const unsigned grid = (unsigned)std::min<uint64_t>(n_pad, 148 * 8);
    auto launch = [&](auto fp4, auto fam) {
        static_assert(tc::PackPlanes<decltype(fam)::value>::N <= tc::MAX_PLANES, "plane list too long");
        if (nibble_rows) tc::pack_ops_kernel<decltype(fp4)::value, decltype(fam)::value, true><<<grid, 256, 0, st>>>(pp);
        else tc::pack_ops_kernel<decltype(fp4)::value, decltype(fam)::value, false><<<grid, 256, 0, st>>>(pp);
    };
    auto by_fam = [&](auto fp4) {
        switch (c->fam) {
        case FAM_SNP: launch(fp4, std::integral_constant<int, FAM_SNP>{}); break;
        case FAM_RAW: launch(fp4, std::integral_constant<int, FAM_RAW>{}); break;
        case FAM_K80: launch(fp4, std::integral_constant<int, FAM_K80>{}); break;
        default: launch(fp4, std::integral_constant<int, FAM_TN93>{}); break;
        }
    };
    if (s.tc_fp4) by_fam(std::true_type{});
    else by_fam(std::false_type{});
    CUDA_CHECK(cudaGetLastError());
    c->tm.pack_launches++;
}

// Buffer for the partial codes pack_ops_kernel meets: room for 1 site in 128 (real data: ~1 in 1000; alignments with
// more than ~0.8 % partial codes run on the LOP3 engine anyway).  If it overflows, the index is filled by rescanning.
void ensure_pp_hits(dg_ctx* c, Device& d, uint64_t n) {
    const uint64_t want = std::min<uint64_t>(std::max<uint64_t>(1u << 20, n * c->width / 128), 0xFFFFFFF0ull);
    if (!d.pp_hit_count) CUDA_CHECK(cudaMalloc(&d.pp_hit_count, 4));
    if (d.pp_hit_cap >= want) return;
    if (d.pp_hits) cudaFree(d.pp_hits);
    d.pp_hits = nullptr; d.pp_hit_cap = 0;
    CUDA_CHECK(cudaMalloc(&d.pp_hits, want * 8));
    d.pp_hit_cap = (uint32_t)want;
}

// Inverted index of the partial ambiguity codes of a resident alignment: the per-site counts were
// accumulated by pp_count_kernel while the chunks arrived; scan them, then scatter the entries
// (counting sort by site).  Synchronises the stream: the entry count sizes the allocation.
void finish_pp_index(dg_ctx* c, Device& d, PlaneSet& s, cudaStream_t st, bool from_hits = false) {
    if (!s.pp.site_off) CUDA_CHECK(cudaMalloc(&s.pp.site_off, (size_t)(c->width + 1) * 4));
    const int ascii = s.input_kind == DG_INPUT_ASCII;
    tc::pp_scan_kernel<<<1, 1024, 0, st>>>(d.pp_cnt, c->width, s.pp.site_off, d.pp_cursor, d.pp_work);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(d.h_pp_total, s.pp.site_off + c->width, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(d.h_pp_work, d.pp_work, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    s.pp.n_entries = *d.h_pp_total;
    s.pp.pair_work = *d.h_pp_work;
    if (s.pp.n_entries > s.pp.cap_entries || !s.pp.entries) {  // grows rarely: reloads of similar data reuse it
        if (s.pp.entries) cudaFree(s.pp.entries);
        s.pp.entries = nullptr;
        s.pp.cap_entries = std::max<uint32_t>(1024, s.pp.n_entries + s.pp.n_entries / 4);
        CUDA_CHECK(cudaMalloc(&s.pp.entries, (size_t)s.pp.cap_entries * 8));
    }
    if (s.pp.n_entries && from_hits && s.pp.n_entries <= d.pp_hit_cap) {   // scatter what pack_ops_kernel collected
        tc::pp_scatter_kernel<<<(unsigned)std::min<uint32_t>((s.pp.n_entries + 255) / 256, 148 * 8), 256, 0, st>>>(
            d.pp_hits, 0, s.pp.n_entries, d.pp_cursor, s.pp.entries);
        CUDA_CHECK(cudaGetLastError());
    } else if (s.pp.n_entries) {
        const unsigned gb = (unsigned)std::min<uint64_t>((s.n * c->width + 255) / 256, 148 * 32);
        tc::pp_fill_kernel<<<gb, 256, 0, st>>>(s.codes, s.n, c->width, ascii, d.pp_cursor, s.pp.entries);
        CUDA_CHECK(cudaGetLastError());
    }
    c->tm.pack_launches += 2;
}

// Two kernels share an SM only when they want the same shared-memory / L1 split.  The persistent GEMM needs the
// maximum carve-out; kernels with (almost) no shared memory get a small one by default, so the combine / packing /
// repair launches of the other stream would wait for the GEMM's CTAs to drain instead of running beside them (f64 and
// LSU work next to the tensor pipe).  Ask for the same carve-out for every side kernel of this device.
template <typename K>
void prefer_max_carveout(K kernel) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
}
void set_side_kernel_carveouts() {
    if (std::getenv("DG_NO_CARVEOUT")) return;   // A/B switch for the overlap measurements
    prefer_max_carveout(tc::tc_combine_kernel<4>);
    prefer_max_carveout(tc::tc_combine_kernel<3>);
    prefer_max_carveout(tc::tc_combine_kernel<2>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_SNP>);  prefer_max_carveout(tc::pack_ops_kernel<false, FAM_SNP>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_RAW>);  prefer_max_carveout(tc::pack_ops_kernel<false, FAM_RAW>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_K80>);  prefer_max_carveout(tc::pack_ops_kernel<false, FAM_K80>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_TN93>); prefer_max_carveout(tc::pack_ops_kernel<false, FAM_TN93>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_SNP, true>);  prefer_max_carveout(tc::pack_ops_kernel<false, FAM_SNP, true>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_RAW, true>);  prefer_max_carveout(tc::pack_ops_kernel<false, FAM_RAW, true>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_K80, true>);  prefer_max_carveout(tc::pack_ops_kernel<false, FAM_K80, true>);
    prefer_max_carveout(tc::pack_ops_kernel<true, FAM_TN93, true>); prefer_max_carveout(tc::pack_ops_kernel<false, FAM_TN93, true>);
    prefer_max_carveout(tc::pp_correct_kernel);
    prefer_max_carveout(tc::pp_correct_chunks_kernel);
    prefer_max_carveout(tc::pp_correct_scan_kernel<true>);
    prefer_max_carveout(tc::pp_correct_scan_kernel<false>);
    prefer_max_carveout(tc::pp_scan_chunk_kernel);
    prefer_max_carveout(tc::pp_scatter_kernel);
    prefer_max_carveout(tc::pp_fill_kernel);
    prefer_max_carveout(tc::build_tile_list_kernel);
    prefer_max_carveout(tc::publish_invalid_kernel);
}

int g_num_sms(int dev) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
}

// Engine choice for one run (DG_OPT_ENGINE: 0 auto, 1 LOP3+POPC tiles, 2 tcgen05 int8 GEMM).
// auto -> tensor cores when the operands exist and the both-partial correction is cheap relative to the
// GEMM (adversarially ambiguous alignments stay on the LOP3 tiles, which need no correction).
bool use_tc(const dg_ctx* c, const PlaneSet& A, const PlaneSet& B) {
    if (!c->tc_wanted()) return false;
    if (!A.tc_ready || !B.tc_ready || A.tc_fp4 != B.tc_fp4 || A.tc_fp4 != c->want_fp4()) {
        if (c->engine >= 2) fail(DG_ERR_STATE, "tensor-engine operands were not built for this engine (set DG_OPT_ENGINE before loading)");
        return false;
    }
    if (c->engine >= 2) return true;
    if (!tc_schedule(c->fam).needs_pp) return true;
    const double work = std::sqrt(A.pp.pair_work * B.pp.pair_work);
    return work <= c->pp_budget() * (double)A.n * (double)B.n;
}

// Where and how the GEMM epilogue stores each accumulator of a launch (tc::AccOp).
struct AccPlan {
    uint32_t nacc = 0;
    uint8_t op[5] = {};
    uint64_t off[5] = {};   // byte offsets from the launch's output pointer
    uint32_t s_pitch = 0, s_colbase = 0;   // scratch rectangle (tc::TcParams); 0 = final counts in the reference's order
    uint32_t ksplit = 1;                   // > 1: split-K launch, every op is ACC_ADD_I32 and `bytes` of scratch are zeroed first
    size_t bytes = 0;
};
// n / n_high with no repair pending: the GEMM writes the final counts
AccPlan direct_plan(const dg_ctx* c) {
    AccPlan a;
    a.nacc = 1;
    a.op[0] = c->u16() ? tc::ACC_DIV3_U16 : tc::ACC_DIV3_U32;
    return a;
}
// The scratch rectangle of a panel: rows [row0, row1) x the columns of the launch's tiles.
struct ScratchGeom { uint64_t rows; uint32_t pitch, colbase; };
ScratchGeom scratch_geom(bool fp4, int mode, uint64_t row0, uint64_t row1, uint64_t n_b) {
    const uint32_t tn = fp4 ? tc::TN_FP4 : tc::TN;
    const uint32_t cb0 = mode == DG_MODE_SQUARE ? (uint32_t)((row0 + 1) / tn) : 0;
    const uint32_t blocks = (uint32_t)((n_b + tn - 1) / tn);
    ScratchGeom g;
    g.rows = row1 - row0;
    g.colbase = cb0 * tn;
    g.pitch = (blocks > cb0 ? blocks - cb0 : 0) * tn;
    return g;
}
// bytes that hold any scratch plan of that panel
size_t scratch_upper_bound(const dg_ctx* c, bool fp4, int mode, uint64_t row0, uint64_t row1, uint64_t n_b) {
    const ScratchGeom g = scratch_geom(fp4, mode, row0, row1, n_b);
    return (size_t)tc_schedule(c->fam).nacc * ((size_t)g.rows * g.pitch * 4 + 256) + 256;
}
// every accumulator of the family into the scratch buffer of a panel.  Accumulator 0 of n / n_high / raw / jc69 is
// 3 * DIFF (up to 3 * width): while the both-partial repair is pending it is kept modulo 2^16 (or as int32 for wide
// alignments) and pp_correct adds to it; else DIFF itself is stored.  Every other sum lies in [-width, width] and is
// stored as int16 when that fits.
AccPlan scratch_plan(const dg_ctx* c, bool fp4, int mode, const Panel& p, uint64_t n_b, bool pp_pending, uint32_t ksplit = 1) {
    static const bool force32 = std::getenv("DG_SCRATCH_I32") != nullptr;
    const TcSchedule& sch = tc_schedule(c->fam);
    const ScratchGeom g = scratch_geom(fp4, mode, p.row0, p.row1, n_b);
    AccPlan a;
    a.nacc = (uint32_t)sch.nacc;
    a.s_pitch = g.pitch; a.s_colbase = g.colbase;
    const bool narrow = c->width <= 32767 && !force32;
    uint64_t off = 0;
    for (int k = 0; k < sch.nacc; k++) {
        int op;
        if (k == 0 && sch.needs_pp) op = pp_pending ? (narrow ? tc::ACC_MOD16 : tc::ACC_RAW_I32) : (narrow ? tc::ACC_DIV3_U16 : tc::ACC_DIV3_U32);
        else op = narrow ? tc::ACC_RAW_I16 : tc::ACC_RAW_I32;
        if (ksplit > 1) op = tc::ACC_ADD_I32;
        a.op[k] = (uint8_t)op;
        a.off[k] = off;
        off += ((uint64_t)g.rows * g.pitch * (tc::acc_op_16(op) ? 2 : 4) + 255) / 256 * 256;
    }
    a.ksplit = ksplit;
    a.bytes = (size_t)off;
    return a;
}

// One launch computes every accumulator of the plan: work items = accumulators x tiles.
void launch_tc_gemm(dg_ctx* c, Device& d, const PlaneSet& A, const PlaneSet& B, int mode, const Panel& p,
                    const AccPlan& plan, void* out, cudaStream_t st, Slot* ws = nullptr,
                    bool tiles_on_device = false) {
    const TcSchedule& sch = tc_schedule(c->fam);
    tc::TcParams tp{};
    tp.n_b = (uint32_t)B.n;
    tp.row0 = (uint32_t)p.row0; tp.row_end = (uint32_t)p.row1;
    tp.square = mode == DG_MODE_SQUARE ? 1 : 0;
    const bool fp4 = A.tc_fp4;
    tp.tn = fp4 ? tc::TN_FP4 : tc::TN;
    tp.col_block0 = tp.square ? (uint32_t)((p.row0 + 1) / tp.tn) : 0;
    tp.gx = (uint32_t)((B.n + tp.tn - 1) / tp.tn) - tp.col_block0;
    // tile variants: 0 = cta_group::2 pairs (default): 512 x 256 block per CTA pair, each CTA stages half of B
    //                1 = 128 x 256 per CTA,  2 = 128 x 256 per CTA in 2-CTA clusters with TMA multicast of B,
    //                3 = 256 x 256 block per CTA (two A sub-tiles, no cluster)
    int variant = fp4 ? 0 : c->tile_variant;   // the FP4 operands run on the cta_group::2 kernel only
    const uint32_t nacc_launch = plan.nacc;
    if (variant == 0 && !fp4) {
        // Small launches cannot fill 74 CTA pairs with 512 x 256 blocks: fall back to 128 x 256 blocks on 148 CTAs
        // when that finishes sooner (cost ~ rounds x rows per SM; the small tile needs ~1.6x the time per MAC).
        const int sms = g_num_sms(d.id);
        const uint64_t rows = p.row1 - p.row0;
        const uint64_t items_p = (rows + 511) / 512 * tp.gx * nacc_launch, items_1 = (rows + 127) / 128 * tp.gx * nacc_launch;
        const double cost_p = (double)((items_p + sms / 2 - 1) / (sms / 2)) * 256.0;
        const double cost_1 = (double)((items_1 + sms - 1) / sms) * 128.0 * 1.6;
        if (cost_1 < cost_p) variant = 1;
    }
    const int cl = (variant == 2 || variant == 0) ? 2 : 1;
    const int mt = (variant == 0 || variant == 3) ? 2 : 1;
    const bool pair = variant == 0;
    tp.gy = (uint32_t)((p.row1 - p.row0 + tc::TM * cl * mt - 1) / (tc::TM * cl * mt));
    tp.n_total = A.n; tp.out_base = p.out_base;
    tp.out = out;
    tp.wp8 = (uint32_t)A.tc_wp8; tp.nsb = (uint32_t)(A.tc_wp8 / tc::KB);
    tp.nacc = nacc_launch;
    for (uint32_t a = 0; a < tp.nacc; a++) {
        tp.npairs[a] = (uint32_t)sch.npairs[a];
        for (int i = 0; i < sch.npairs[a]; i++) { tp.pa[a][i] = sch.pa[a][i]; tp.pb[a][i] = sch.pb[a][i]; }
    }
    for (uint32_t a = 0; a < tp.nacc; a++) { tp.acc_off[a] = plan.off[a]; tp.acc_op[a] = plan.op[a]; }
    tp.s_pitch = plan.s_pitch; tp.s_colbase = plan.s_colbase;
    tp.ksplit = std::max<uint32_t>(1, plan.ksplit);
    static const bool l2_hint = std::getenv("DG_TC_L2_HINT") != nullptr;
    tp.l2_evict_last = l2_hint ? 1u : 0u;
    static const uint32_t raster_env = std::getenv("DG_TC_RASTER") ? (uint32_t)std::max(1, std::atoi(std::getenv("DG_TC_RASTER"))) : 0u;
    tp.raster_g = raster_env ? raster_env : tc::RASTER_G;
    if (plan.s_pitch && (plan.s_pitch != tp.gx * tp.tn || plan.s_colbase != tp.col_block0 * tp.tn))
        fail(DG_ERR_STATE, "internal: scratch geometry does not match the launch's tiles");
    tp.probe = d.clk_probe;
    tp.stages = pair ? tc::StageCfg<2, true>::N : (mt == 2 ? tc::StageCfg<2>::N : tc::StageCfg<1>::N);
    if (const char* e = std::getenv("DG_TC_STAGES")) tp.stages = (uint32_t)std::min<int>((int)tp.stages, std::max(1, std::atoi(e)));
    if (tp.gx == 0 || tp.gy == 0) return;
    uint64_t live = (uint64_t)tp.gx * tp.gy;
    if (tp.square && ws && tp.gx < (1u << 20) && tp.gy < (1u << 12)) {
        // Upper-triangle panels: tiles on / below the diagonal are dead.  Deal only the live ones (same raster
        // order: bands of RASTER_G row blocks, column-major) so every CTA (pair) gets the same amount of work.
        if (ws->tiles_cap < live) {
            if (ws->d_tiles) cudaFree(ws->d_tiles);
            if (ws->h_tiles) cudaFreeHost(ws->h_tiles);
            ws->d_tiles = ws->h_tiles = nullptr; ws->tiles_cap = 0;
            CUDA_CHECK(cudaMalloc(&ws->d_tiles, live * 4));
            CUDA_CHECK(cudaHostAlloc(&ws->h_tiles, live * 4, cudaHostAllocDefault));
            ws->tiles_cap = live;
        }
        const uint32_t rows_per_block = (uint32_t)(tc::TM * cl * mt);
        uint32_t n_live = 0;
        for (uint32_t band0 = 0; band0 < tp.gy; band0 += tp.raster_g) {
            const uint32_t gb = std::min<uint32_t>(tp.raster_g, tp.gy - band0);
            for (uint32_t bx = 0; bx < tp.gx; bx++)
                for (uint32_t by = band0; by < band0 + gb; by++) {
                    const uint64_t rowS0 = (uint64_t)tp.row0 + (uint64_t)by * rows_per_block;
                    const uint64_t rowB0 = (uint64_t)(tp.col_block0 + bx) * tp.tn;
                    if (rowS0 >= tp.row_end || rowB0 + tp.tn <= rowS0 + 1) continue;
                    ws->h_tiles[n_live++] = (by << 20) | bx;
                }
        }
        if (n_live == 0) return;
        if (tiles_on_device) {   // sessions: no small H2D copy behind the bulk uploads on the copy engine
            tc::build_tile_list_kernel<<<1, 1024, 0, st>>>(tp.gx, tp.gy, tp.raster_g, tp.row0, tp.row_end, rows_per_block,
                                                          tp.col_block0, tp.tn, ws->d_tiles);
            CUDA_CHECK(cudaGetLastError());
        } else {
            CUDA_CHECK(cudaMemcpyAsync(ws->d_tiles, ws->h_tiles, (size_t)n_live * 4, cudaMemcpyHostToDevice, st));
        }
        tp.tile_list = ws->d_tiles; tp.n_live = n_live;
        live = n_live;
    }
    const uint64_t tiles = live * tp.nacc * tp.ksplit;
    if (cl == 1) {
        auto kern = mt == 2 ? tc::tc_gemm_kernel<1, 2> : tc::tc_gemm_kernel<1, 1>;
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)g_num_sms(d.id));
        kern<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(A.map_a, B.map_a, tp);
        CUDA_CHECK(cudaGetLastError());
    } else {
        auto kern2 = fp4 ? tc::tc_gemm_kernel<2, 2, true, true>
                         : (pair ? tc::tc_gemm_kernel<2, 2, true, false> : tc::tc_gemm_kernel<2, 1, false, false>);
        CUDA_CHECK(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * (unsigned)std::min<uint64_t>(tiles, (uint64_t)g_num_sms(d.id) / 2));
        cfg.blockDim = dim3(tc::THREADS);
        cfg.dynamicSmemBytes = tc::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern2, A.map_a, fp4 ? B.map_b : B.map_a, tp));
    }
    c->tm.count_launches++;
}

// Will this (A rows, B columns) product need the both-partial repair of accumulator 0?  A stream batch has
// no index: its partial codes are found by scanning its V planes, so only B's index decides.
bool tc_pp_pending(const dg_ctx* c, const PlaneSet& A, const PlaneSet& B, bool a_is_batch) {
    return tc_schedule(c->fam).needs_pp && B.pp.n_entries != 0 && (a_is_batch || A.pp.n_entries != 0);
}

// Raw int32 sums of every accumulator (scratch[accumulator][panel pair]) -> the reference's counts -> the result
// (uint16 / uint32 count, f64 distance through the epi_* epilogues, or the debug counts).
void launch_combine(dg_ctx* c, Device& d, const PlaneSet& A, const PlaneSet& B, int mode, const Panel& p, void* d_out,
                    int* scratch, const AccPlan& plan, bool swap_roles, bool counts, cudaStream_t st, bool counts16 = false) {
    tc::CombineParams cp{};
    cp.acc = reinterpret_cast<const uint8_t*>(scratch);
    for (uint32_t a = 0; a < plan.nacc; a++) { cp.acc_off[a] = plan.off[a]; cp.acc_op[a] = plan.op[a]; }
    cp.s_pitch = plan.s_pitch; cp.s_colbase = plan.s_colbase;
    cp.literal = std::getenv("DG_EPI_LITERAL") != nullptr;
    cp.a_acgt = A.acgt; cp.b_acgt = B.acgt;
    cp.n_b = (uint32_t)B.n; cp.row0 = (uint32_t)p.row0; cp.row_end = (uint32_t)p.row1;
    cp.square = mode == DG_MODE_SQUARE ? 1 : 0;
    cp.col0 = cp.square ? (uint32_t)((p.row0 + 1) / 256 * 256) : 0;
    cp.swap_roles = swap_roles ? 1 : 0; cp.measure = c->measure; cp.fam = c->fam;
    cp.result = counts ? tc::RES_COUNTS : (c->measure <= 1 ? (c->u16() ? tc::RES_U16 : tc::RES_U32) : (counts16 ? tc::RES_COUNTS16 : tc::RES_F64));
    cp.n_total = A.n; cp.out_base = p.out_base; cp.out = d_out;
    // virtual blocks: 256-column strips x row phases (~16 per SM)
    const unsigned gx = (unsigned)((B.n - cp.col0 + 255) / 256);
    const unsigned gy = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(p.row1 - p.row0, (uint64_t)g_num_sms(d.id) * 16 / std::max(1u, gx)));
    if (gx == 0) return;
    static const int per_sm = std::getenv("DG_COMBINE_PER_SM") ? std::max(0, std::atoi(std::getenv("DG_COMBINE_PER_SM"))) : 0;
    const unsigned grid = per_sm ? (unsigned)std::min<uint64_t>((uint64_t)gx * gy, (uint64_t)g_num_sms(d.id) * per_sm) : gx * gy;
    // f64 measures: 3 CTAs per SM (85 registers) leave the compiler room to interleave the independent division / log chains
    // of one pair; the integer paths keep 4 (DG_COMBINE_MINB overrides for experiments)
    static const int minb_env = std::getenv("DG_COMBINE_MINB") ? std::atoi(std::getenv("DG_COMBINE_MINB")) : 0;
    const int minb = minb_env ? minb_env : 4;
    if (minb == 3) tc::tc_combine_kernel<3><<<grid, 256, 0, st>>>(cp, gx, gy);
    else if (minb == 2) tc::tc_combine_kernel<2><<<grid, 256, 0, st>>>(cp, gx, gy);
    else tc::tc_combine_kernel<4><<<grid, 256, 0, st>>>(cp, gx, gy);
    CUDA_CHECK(cudaGetLastError());
    c->tm.count_launches++;
}

// tcgen05 variant of enqueue_panel_kernel.  n / n_high with no correction pending: one GEMM writes the result
// directly.  Otherwise (and for the debug counts): one GEMM per accumulator into `scratch`, the both-partial
// repair of accumulator 0, then tc_combine_kernel.
uint64_t live_pair_tiles(int mode, uint64_t r0, uint64_t r1, uint64_t n_cols, uint64_t tn = 256);

// Split-K for small launches: a panel with fewer work items than CTA pairs leaves most of the chip idle while a few
// pairs walk the whole K range (1,000 x 1,000 records: 16 items of 468 K blocks on 74 pairs).  Dealing each tile's K
// range to `ks` pairs (partial sums added into a zeroed int32 scratch) fills the chip.  Cost model in K blocks: rounds x
// (blocks per item + a fixed ~30 for the launch ramp, the epilogue and its atomics); split only for a clear win.
uint32_t choose_ksplit(const dg_ctx* c, const PlaneSet& A, const PlaneSet& B, int mode, const Panel& p) {
    const char* fe = std::getenv("DG_KSPLIT");   // tests: 1 = never split, k = always k ways
    const int forced = fe ? std::atoi(fe) : 0;
    const TcSchedule& sch = tc_schedule(c->fam);
    const uint64_t KT = (uint64_t)sch.npairs[0] * (A.tc_wp8 / tc::KB);
    if (forced > 0) return (uint32_t)std::min<uint64_t>(forced, std::max<uint64_t>(1, KT));
    const uint64_t items = live_pair_tiles(mode, p.row0, p.row1, B.n, A.tc_fp4 ? tc::TN_FP4 : tc::TN) * sch.nacc;
    const uint64_t slots = 74, ovh = 30;
    if (items == 0 || items >= 2 * slots) return 1;
    auto cost = [&](uint64_t ks) { return (items * ks + slots - 1) / slots * (KT / ks + ovh); };
    uint64_t best = 1, best_cost = cost(1);
    for (uint64_t ks = 2; ks <= 16 && KT / ks >= 16; ks++)
        if (cost(ks) < best_cost) { best_cost = cost(ks); best = ks; }
    return best_cost * 4 <= cost(1) * 3 ? (uint32_t)best : 1u;
}

// Returns the stream that carries the panel's last kernel: `st`, or `post` when the repair + combine passes were put
// there (behind `ws->gemm_done`).
cudaStream_t enqueue_panel_tc(dg_ctx* c, Device& d, const PlaneSet& A, const PlaneSet& B, int mode, const Panel& p,
                              void* d_out, int* scratch, bool swap_roles, bool counts, cudaStream_t st, bool a_is_batch = false,
                              Slot* ws = nullptr, cudaEvent_t after_gemm = nullptr, cudaStream_t post = nullptr) {
    const TcSchedule& sch = tc_schedule(c->fam);
    const bool pp_pending = tc_pp_pending(c, A, B, a_is_batch);
    const uint32_t ks = scratch ? choose_ksplit(c, A, B, mode, p) : 1u;
    if (c->fam == FAM_SNP && !counts && !pp_pending && ks == 1) {
        launch_tc_gemm(c, d, A, B, mode, p, direct_plan(c), d_out, st, ws);
        if (after_gemm) CUDA_CHECK(cudaEventRecord(after_gemm, st));
        return st;
    }
    if (!scratch) fail(DG_ERR_STATE, "tensor engine: no scratch buffer");
    const AccPlan plan = scratch_plan(c, A.tc_fp4, mode, p, B.n, pp_pending, ks);
    if (ks > 1) CUDA_CHECK(cudaMemsetAsync(scratch, 0, plan.bytes, st));
    static const bool skip_gemm = std::getenv("DG_SKIP_GEMM") != nullptr, skip_combine = std::getenv("DG_SKIP_COMBINE") != nullptr;  // timing experiments only
    if (!skip_gemm) launch_tc_gemm(c, d, A, B, mode, p, plan, scratch, st, ws);
    if (after_gemm) CUDA_CHECK(cudaEventRecord(after_gemm, st));
    if (post && ws && ws->gemm_done) {   // the rest of the panel runs on the post stream
        CUDA_CHECK(cudaEventRecord(ws->gemm_done, st));
        CUDA_CHECK(cudaStreamWaitEvent(post, ws->gemm_done, 0));
        st = post;
    }
    if (pp_pending) {
        tc::PpCorrParams cp{};
        cp.b_entries = B.pp.entries; cp.b_off = B.pp.site_off;
        cp.row0 = (uint32_t)p.row0; cp.row_end = (uint32_t)p.row1; cp.n_b = (uint32_t)B.n;
        cp.square = mode == DG_MODE_SQUARE ? 1 : 0;
        cp.s_pitch = plan.s_pitch; cp.s_colbase = plan.s_colbase; cp.op = plan.op[0]; cp.out = scratch;
        if (a_is_batch) {
            cp.a_ops = A.tc_ops; cp.a_nplanes = (uint32_t)sch.nplanes; cp.a_wp8 = (uint32_t)A.tc_wp8; cp.a_vplane0 = TC_VPLANE0;
            const uint64_t units = (p.row1 - p.row0) * (A.tc_wp8 / 16);
            const unsigned gb = (unsigned)std::min<uint64_t>((units + 255) / 256, 148 * 16);
            if (A.tc_fp4) tc::pp_correct_scan_kernel<true><<<gb, 256, 0, st>>>(cp);
            else tc::pp_correct_scan_kernel<false><<<gb, 256, 0, st>>>(cp);
        } else {
            cp.a_entries = A.pp.entries; cp.a_n = A.pp.n_entries;
            tc::pp_correct_kernel<<<(A.pp.n_entries + 255) / 256, 256, 0, st>>>(cp);
        }
        CUDA_CHECK(cudaGetLastError());
        c->tm.count_launches++;
    }
    const bool c16 = !counts && !a_is_batch && ws && c->counts16();   // resident panels only
    if (ws) ws->kind = c16 ? DG_RESULT_COUNTS16 : -1;
    if (!skip_combine) launch_combine(c, d, A, B, mode, p, d_out, scratch, plan, swap_roles, counts, st, c16);
    return st;
}

void ensure_scratch(Slot& s, size_t bytes) {
    if (s.scratch_cap >= bytes) return;
    if (s.d_scratch) cudaFree(s.d_scratch);
    s.d_scratch = nullptr; s.scratch_cap = 0;
    CUDA_CHECK(cudaMalloc(&s.d_scratch, bytes));
    s.scratch_cap = bytes;
}

uint64_t sq_off(uint64_t n, uint64_t i) { return i * (2 * n - i - 1) / 2; }

// Live 512 x 256 tiles (the tensor engine's CTA-pair block) of rows [r0, r1) against n_cols columns.
uint64_t live_pair_tiles(int mode, uint64_t r0, uint64_t r1, uint64_t n_cols, uint64_t tn) {
    const uint64_t col_blocks = (n_cols + tn - 1) / tn;
    if (mode != DG_MODE_SQUARE) return (r1 - r0 + 511) / 512 * col_blocks;
    uint64_t live = 0;
    for (uint64_t rs = r0; rs < r1; rs += 512) {
        const uint64_t first = (rs + 1) / tn;   // first column block with a column > rs
        live += col_blocks > first ? col_blocks - first : 0;
    }
    return live;
}

// Which part (rank / process) owns which panel.  Panels of a triangle differ in size and their number is rarely a
// multiple of the part count, so dealing them round-robin leaves the largest share 12 - 20 % above the mean at 8 parts
// (49 panels: some parts get 7, some 6).  Longest-processing-time-first instead: panels in descending size (ties: lower
// index first) go to the part with the smallest load so far (ties: lower part).  Deterministic, so every rank computes the
// same assignment from the same plan; dg_plan_parts exposes it.
std::vector<uint32_t> assign_parts(const uint64_t* n_results, size_t count, uint32_t n_parts) {
    std::vector<uint32_t> part_of(count, 0);
    if (n_parts <= 1) return part_of;
    std::vector<size_t> order(count);
    for (size_t k = 0; k < count; k++) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return n_results[a] > n_results[b]; });
    std::vector<uint64_t> load(n_parts, 0);
    for (size_t k : order) {
        uint32_t best = 0;
        for (uint32_t q = 1; q < n_parts; q++)
            if (load[q] < load[best]) best = q;
        part_of[k] = best;
        load[best] += n_results[k];
    }
    return part_of;
}
std::vector<Panel> panels_of_part(const std::vector<Panel>& all, uint32_t part, uint32_t n_parts) {
    std::vector<uint64_t> sizes(all.size());
    for (size_t k = 0; k < all.size(); k++) sizes[k] = all[k].n_results;
    const std::vector<uint32_t> part_of = assign_parts(sizes.data(), sizes.size(), n_parts);
    std::vector<Panel> mine;
    for (size_t k = 0; k < all.size(); k++)
        if (part_of[k] == part) mine.push_back(all[k]);
    return mine;
}

// tn / items: tile width of the tensor engine in use (240 fp4, 256 int8) and work items per tile (= accumulators of the
// family's schedule): one persistent launch deals live tiles x items to the 74 CTA pairs.
std::vector<Panel> make_panels(size_t panel_bytes, size_t elem_bytes, int mode, uint64_t n_rows_total,
                               uint64_t n_cols, int tm, uint64_t tn = 256, uint64_t items = 1) {
    // rows per panel: a multiple of the tile height sized so one panel's results ~ panel_bytes.  Among the
    // candidates between half and all of that budget, take the one whose tile count fills whole rounds of the
    // 74 CTA pairs best (a panel is one persistent launch: a ragged last round idles SMs).
    std::vector<Panel> v;
    const uint64_t rows_major = mode == DG_MODE_SQUARE ? (n_rows_total ? n_rows_total - 1 : 0) : n_rows_total;
    if (rows_major == 0 || n_cols == 0) return v;
    const uint64_t quantum = std::max<uint64_t>(tm, 512);  // a multiple of every tensor-engine block height (256 / 512 rows)
    const uint64_t SLOTS = 74;
    for (uint64_t r = 0; r < rows_major;) {
        // square rows get shorter as r grows: budget by the pairs left of row r
        const uint64_t row_len = mode == DG_MODE_SQUARE ? n_rows_total - 1 - r : n_cols;
        uint64_t per = panel_bytes / (elem_bytes * std::max<uint64_t>(1, row_len));
        per = std::min<uint64_t>(per, (uint64_t)tm * 32768);  // gridDim.y limit
        per = std::max<uint64_t>(quantum, per / quantum * quantum);
        uint64_t best_rows = per;
        if (r + per < rows_major) {
            // A panel of only a few rounds loses a large share of its last one (84 tiles = 1.14 rounds of 74 pairs run at 57 %),
            // so small budgets may grow up to twice their size for a fuller last round (a small penalty keeps ties small).
            const uint64_t live_per = live_pair_tiles(mode, r, r + per, n_cols, tn) * items;
            // (launches of less than half a round are split along K instead: choose_ksplit); panels of up to 40 rounds may still
            // grow by a quarter: 17.03 rounds (config 3: 1,260 work items) run as 18, 19.9 as 20.
            const uint64_t cap = (uint64_t)tm * 32768 / quantum * quantum;
            uint64_t hi = per;
            if (live_per >= SLOTS / 2 && live_per < 3 * SLOTS) hi = std::min<uint64_t>(2 * per, cap);
            else if (live_per >= 3 * SLOTS && live_per < 40 * SLOTS) hi = std::min<uint64_t>(per + std::max<uint64_t>(quantum, per / 4 / quantum * quantum), cap);
            double best_score = -1;
            for (uint64_t cand = hi; cand >= quantum && cand * 2 >= per; cand -= quantum) {
                const uint64_t live = live_pair_tiles(mode, r, std::min(rows_major, r + cand), n_cols, tn) * items;
                const double fill = live ? (double)live / (double)((live + SLOTS - 1) / SLOTS * SLOTS) : 0.0;
                const double score = fill - (cand > per ? 0.04 * (double)(cand - per) / (double)per : 0.0);
                if (score > best_score + 1e-9) { best_score = score; best_rows = cand; }
                if (cand == quantum) break;
            }
        }
        if (r + best_rows < rows_major && rows_major - (r + best_rows) < quantum) best_rows = rows_major - r;  // no sliver panel
        Panel p;
        p.row0 = r;
        p.row1 = std::min(rows_major, r + best_rows);
        if (mode == DG_MODE_SQUARE) {
            p.out_base = sq_off(n_rows_total, p.row0);
            p.n_results = sq_off(n_rows_total, p.row1) - p.out_base;
        } else {
            p.out_base = p.row0 * n_cols;
            p.n_results = (p.row1 - p.row0) * n_cols;
        }
        v.push_back(p);
        r = p.row1;
    }
    return v;
}

// Enqueue the count tiles of one panel on d.compute, writing slot.d_out.
template <bool COUNTS>
void enqueue_panel_kernel(dg_ctx* c, Device& d, const PlaneSet& A, const PlaneSet& B, int mode,
                          const Panel& p, void* d_out, bool swap_roles, cudaStream_t st) {
    const TileShape ts = tile_shape(c->fam, c->tile_variant);
    CountParams cp{};
    cp.a_core = A.core; cp.a_aux = A.aux; cp.a_acgt = A.acgt;
    cp.b_core = B.core; cp.b_aux = B.aux; cp.b_acgt = B.acgt;
    cp.n_b = (uint32_t)B.n;
    cp.wp = c->wp;
    cp.row0 = (uint32_t)p.row0;
    cp.row_end = (uint32_t)p.row1;
    cp.square = mode == DG_MODE_SQUARE ? 1 : 0;
    cp.swap_roles = swap_roles ? 1 : 0;
    cp.n_total = A.n;
    cp.out_base = COUNTS ? 0 : p.out_base;
    cp.out = d_out;
    cp.measure = c->measure;
    cp.out_u16 = (!COUNTS && c->u16()) ? 1 : 0;
    cp.col_block0 = cp.square ? (uint32_t)((p.row0 + 1) / ts.tn) : 0;
    const uint64_t col_blocks_total = (B.n + ts.tn - 1) / ts.tn;
    dim3 grid((unsigned)(col_blocks_total - cp.col_block0), (unsigned)((p.row1 - p.row0 + ts.tm - 1) / ts.tm));
    if (grid.x == 0 || grid.y == 0) return;
    if (grid.y > 65535) fail(DG_ERR_INVALID_ARG, "panel too tall for one launch (%u row blocks)", grid.y);
    launch_count<COUNTS>(c->fam, c->tile_variant, cp, grid, st);
    c->tm.count_launches++;
}

// The SM clock the tensor-engine launches actually saw: CTA 0 of every launch adds its clock64 and globaltimer spans
// to the device's probe; called when the streams are idle.
void harvest_clock(dg_ctx* c) {
    Device& d = c->devs[0];
    if (!d.clk_probe) return;
    unsigned long long h[8];
    if (cudaSetDevice(d.id) != cudaSuccess || cudaMemcpy(h, d.clk_probe, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) return;
    cudaMemset(d.clk_probe, 0, sizeof h);
    c->clk_cycles += (double)h[4]; c->clk_ns += (double)h[5];
    static const bool verbose = std::getenv("DG_CLOCK_PROBE") != nullptr;
    if (verbose && h[5]) fprintf(stderr, "[dg clock probe] GEMM launches %.3f ms, SM clock %.0f MHz\n", h[5] * 1e-6, (double)h[4] / (double)h[5] * 1e3);
}

void harvest_kernel_time(dg_ctx* c, Slot& s) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.k_start, s.k_stop) == cudaSuccess) c->tm.count_ms += ms;
}


// ---- DG_INPUT_NIBBLE -------------------------------------------------------------------------------------------------
bool valid_input_kind(int k) { return k == DG_INPUT_PARADIS || k == DG_INPUT_ASCII || k == DG_INPUT_NIBBLE; }
uint64_t input_stride(const dg_ctx* c, int kind) { return kind == DG_INPUT_NIBBLE ? (c->width + 1) / 2 : c->width; }
// what the pack kernels see: nibble rows are unpacked to Paradis bytes on the device first
int device_kind(int kind) { return kind == DG_INPUT_NIBBLE ? DG_INPUT_PARADIS : kind; }
void ensure_nib(dg_ctx* c, PlaneSet& s, uint64_t n) {
    const uint64_t bytes = n * input_stride(c, DG_INPUT_NIBBLE);
    if (s.nib_cap >= bytes) return;
    if (s.nib) cudaFree(s.nib);
    s.nib = nullptr; s.nib_cap = 0;
    CUDA_CHECK(cudaMalloc(&s.nib, std::max<uint64_t>(bytes, 16)));
    s.nib_cap = bytes;
}
void enqueue_nibble_unpack(dg_ctx* c, const uint8_t* d_nib, uint8_t* d_codes, uint64_t n, cudaStream_t st) {
    const uint64_t wb = input_stride(c, DG_INPUT_NIBBLE);
    const uint64_t units = n * ((wb + 7) / 8);
    const unsigned grid = (unsigned)std::min<uint64_t>((units + 255) / 256, 148 * 32);
    tc::nibble_unpack_kernel<<<std::max(1u, grid), 256, 0, st>>>(d_nib, d_codes, n, c->width, wb);
    CUDA_CHECK(cudaGetLastError());
    c->tm.pack_launches++;
}

// ---- resident alignments: buffer reuse, pipelined upload, lazy LOP3 planes ---------------------------
// (Re)size the buffers of a resident alignment; buffers are kept when the new alignment fits.
void reserve_resident(dg_ctx* c, PlaneSet& s, uint64_t n, bool want_tc) {
    const uint64_t n_pad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    const bool fits = s.codes && s.cap_pad >= n_pad && (!want_tc || (s.tc_ops && s.tc_fp4 == c->want_fp4()));
    if (!fits) {
        free_set(s);
        s.cap_pad = n_pad;
        CUDA_CHECK(cudaMalloc(&s.codes, (size_t)n_pad * c->width));
        CUDA_CHECK(cudaMalloc(&s.acgt, (size_t)n_pad * 4 * sizeof(uint32_t)));
        if (want_tc) {
            s.n = n; s.n_pad = n_pad;
            alloc_tc_operands(c, s);
        }
    }
    s.n = n;
    s.n_pad = n_pad;
    s.lop3_ready = false;
    s.tc_ready = false;
}

// Build the LOP3 bit-planes of a resident alignment from its device-resident codes, once, on demand.
void ensure_lop3(dg_ctx* c, Device& d, PlaneSet& s) {
    if (s.lop3_ready) return;
    CUDA_CHECK(cudaSetDevice(d.id));
    const size_t plane_bytes = (size_t)s.cap_pad * c->wp * sizeof(uint4);
    if (!s.core) CUDA_CHECK(cudaMalloc(&s.core, plane_bytes));
    if (!s.aux && c->fam != FAM_SNP) CUDA_CHECK(cudaMalloc(&s.aux, plane_bytes));
    CUDA_CHECK(cudaEventRecord(d.slot[0].p_start, d.compute));
    enqueue_pack(c, d.d_invalid + 2, s, s.codes, s.n, s.input_kind, !s.acgt_from_host, false, d.compute);
    CUDA_CHECK(cudaEventRecord(d.slot[0].p_stop, d.compute));
    CUDA_CHECK(cudaStreamSynchronize(d.compute));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, d.slot[0].p_start, d.slot[0].p_stop));
    c->tm.pack_ms += ms;
    s.lop3_ready = true;
}

// ---- square / rect runs ------------------------------------------------------------------------
void run_panel_list(dg_ctx* c, int mode, const std::vector<Panel>& mine, bool tc_run, dg_sink_fn sink, void* user,
                    bool device_only, bool repack_descending = false);
void ensure_pp_index(dg_ctx* c, Device& d, PlaneSet& s);

void run_mode(dg_ctx* c, int mode, uint32_t part, uint32_t n_parts, dg_sink_fn sink, void* user,
              uint32_t flags) {
    if (c->streaming || c->sq_open) fail(DG_ERR_STATE, "a session is open");
    if (n_parts == 0 || part >= n_parts) fail(DG_ERR_INVALID_ARG, "bad part %u of %u", part, n_parts);
    const int wb = mode == DG_MODE_SQUARE ? 0 : 1;
    for (auto& d : c->devs) {
        if (d.set[0].n == 0) fail(DG_ERR_STATE, "alignment 0 is not loaded");
        if (d.set[wb].n == 0) fail(DG_ERR_STATE, "alignment %d is not loaded", wb);
    }
    const bool device_only = (flags & DG_RUN_DEVICE_ONLY) != 0;
    if (!device_only && sink == nullptr) fail(DG_ERR_INVALID_ARG, "sink is NULL");
    const double t0 = wall_ms();
    const int ndev = (int)c->devs.size();
    for (auto& d : c->devs) {
        CUDA_CHECK(cudaSetDevice(d.id));
        CUDA_CHECK(cudaEventRecord(d.run_start, d.compute));
        CUDA_CHECK(cudaStreamWaitEvent(d.compute2, d.run_start, 0));
    }

    for (auto& d : c->devs)
        for (int w = 0; w <= wb; w++) ensure_pp_index(c, d, d.set[w]);
    const bool tc_run = use_tc(c, c->devs[0].set[0], c->devs[0].set[wb]);
    c->last_engine = tc_run ? (c->devs[0].set[0].tc_fp4 ? 3 : 2) : 1;
    if (!tc_run)
        for (auto& d : c->devs)
            for (int w = 0; w <= wb; w++) ensure_lop3(c, d, d.set[w]);
    const bool overlap_repack = (flags & DG_RUN_REPACK) && device_only && tc_run && mode == DG_MODE_SQUARE && c->repack_overlap;
    if (overlap_repack) {
        for (auto& d : c->devs) {
            CUDA_CHECK(cudaSetDevice(d.id));
            CUDA_CHECK(cudaStreamWaitEvent(d.prep, d.run_start, 0));
        }
    } else if (flags & DG_RUN_REPACK) {  // re-run the operand packing of the engine in use from the resident codes
        for (auto& d : c->devs) {
            CUDA_CHECK(cudaSetDevice(d.id));
            for (int w = 0; w <= wb; w++) {
                PlaneSet& s = d.set[w];
                CUDA_CHECK(cudaEventRecord(d.slot[0].p_start, d.compute));
                if (tc_run) enqueue_tc_pack(c, s, s.codes, s.n, s.input_kind, !s.acgt_from_host && c->fam == FAM_TN93, d.compute);
                else enqueue_pack(c, d.d_invalid + 2, s, s.codes, s.n, s.input_kind, !s.acgt_from_host, false, d.compute);
                CUDA_CHECK(cudaEventRecord(d.slot[0].p_stop, d.compute));
                CUDA_CHECK(cudaEventSynchronize(d.slot[0].p_stop));
                float ms = 0;
                CUDA_CHECK(cudaEventElapsedTime(&ms, d.slot[0].p_start, d.slot[0].p_stop));
                c->tm.pack_ms += ms;
            }
        }
    }

    const PlaneSet& A0 = c->devs[0].set[0];
    const PlaneSet& B0 = c->devs[0].set[wb];
    const TileShape ts = tile_shape(c->fam, c->tile_variant);
    std::vector<Panel> all = make_panels(c->panel_bytes, c->elem_bytes(), mode, A0.n, B0.n, ts.tm, c->plan_tn(), c->plan_items());
    std::vector<Panel> mine = panels_of_part(all, part, n_parts);
    if (overlap_repack) std::reverse(mine.begin(), mine.end());
    run_panel_list(c, mode, mine, tc_run, sink, user, device_only, overlap_repack);
    c->tm.run_ms = 0;
    for (auto& d : c->devs) {
        CUDA_CHECK(cudaSetDevice(d.id));
        CUDA_CHECK(cudaEventRecord(d.slot[1].in_ready, d.compute2));
        CUDA_CHECK(cudaStreamWaitEvent(d.compute, d.slot[1].in_ready, 0));
        CUDA_CHECK(cudaEventRecord(d.slot[2].in_ready, d.post));
        CUDA_CHECK(cudaStreamWaitEvent(d.compute, d.slot[2].in_ready, 0));
        CUDA_CHECK(cudaEventRecord(d.run_stop, d.compute));
        CUDA_CHECK(cudaEventSynchronize(d.run_stop));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, d.run_start, d.run_stop));
        c->tm.run_ms = std::max<double>(c->tm.run_ms, ms);
    }
    harvest_clock(c);
    c->tm.total_ms = wall_ms() - t0;
}

// The panels of `mine` in order: panel k on device k % ndev, two ring slots per device (the D2H of a panel and
// its sink call overlap the tiles of the next ones); the sink sees the panels serially, in list order.
// repack_descending (kernel-only square runs with DG_RUN_REPACK): `mine` is in DESCENDING row order and the operand
// planes are re-packed on the way: before a panel starts, the records from its first row up to what is already packed
// are packed on the high-priority prep stream, so the HBM-bound packing of the next panel's rows overlaps the
// tensor-bound tiles of this one (a panel needs the records at or above its first row only).
void run_panel_list(dg_ctx* c, int mode, const std::vector<Panel>& mine, bool tc_run, dg_sink_fn sink, void* user,
                    bool device_only, bool repack_descending) {
    std::vector<cudaEvent_t> tr0, tr1, tr2;   // DG_TRACE (declared before `dump`, which reads them when the function leaves)
    struct TraceDump {   // prints and frees the DG_TRACE events when the function leaves
        std::vector<cudaEvent_t>*a = nullptr, *b = nullptr, *c = nullptr; cudaEvent_t base = nullptr; const std::vector<Panel>* panels = nullptr;
        ~TraceDump() {
            if (!a || a->empty()) return;
            cudaDeviceSynchronize();
            for (size_t k = 0; k < a->size(); k++) {
                float t0 = 0, t1 = 0, t2 = 0;
                cudaEventElapsedTime(&t0, base, (*a)[k]); cudaEventElapsedTime(&t1, base, (*b)[k]); cudaEventElapsedTime(&t2, base, (*c)[k]);
                fprintf(stderr, "  [dg trace] panel %2zu rows [%6llu,%6llu): start %7.3f  gemm end %7.3f  panel end %7.3f ms\n", k,
                        (unsigned long long)(*panels)[k].row0, (unsigned long long)(*panels)[k].row1, t0, t1, t2);
                cudaEventDestroy((*a)[k]); cudaEventDestroy((*b)[k]); cudaEventDestroy((*c)[k]);
            }
        }
    } dump;
    const int wb = mode == DG_MODE_SQUARE ? 0 : 1;
    std::vector<uint64_t> packed_lo(c->devs.size(), c->devs[0].set[0].n);
    const int ndev = (int)c->devs.size();
    const PlaneSet& A0 = c->devs[0].set[0];
    const PlaneSet& B0 = c->devs[0].set[wb];
    size_t max_bytes = 0;
    for (auto& p : mine) max_bytes = std::max(max_bytes, (size_t)p.n_results * c->elem_bytes());
    for (auto& d : c->devs) {
        CUDA_CHECK(cudaSetDevice(d.id));
        ensure_out_ring(c, d, std::max<size_t>(max_bytes, 256));
        if (c->u8() && !device_only)
            for (auto& sl : d.slot) ensure_narrow(sl, max_bytes / c->elem_bytes());
        bool any_split = false;
        if (tc_run)
            for (auto& p : mine) any_split = any_split || choose_ksplit(c, d.set[0], d.set[wb], mode, p) > 1;
        if (tc_run && (c->fam != FAM_SNP || any_split || tc_pp_pending(c, d.set[0], d.set[wb], false))) {
            size_t sb = 0;
            for (auto& p : mine) sb = std::max(sb, scratch_upper_bound(c, d.set[0].tc_fp4, mode, p.row0, p.row1, B0.n));
            for (auto& sl : d.slot) ensure_scratch(sl, sb);
        }
    }

    const int K = (int)mine.size();
    constexpr int NS = Device::NSLOT;
    const int lookahead = NS * ndev;
    // DG_POST_STREAM=1 (experiment): repair + combine passes on a third stream, so that the GEMMs run back to back.  Measured
    // on configs 2 and 3: no gain (7.03 vs 6.86 ms, 6.57 vs 6.43 ms): the tensor pass and the f64 / repair passes slow each
    // other down by what the overlap wins (the step runs at the power cap either way), so the simpler order is the default.
    static const bool use_post = std::getenv("DG_POST_STREAM") != nullptr;
    // DG_TRACE: device timeline of the panels of device 0 (start, end of the GEMM, end of the panel) on stderr
    const bool trace = std::getenv("DG_TRACE") != nullptr && ndev == 1;
    if (trace)
        for (int k = 0; k < K; k++)
            for (auto* v : {&tr0, &tr1, &tr2}) { cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); v->push_back(e); }
    if (trace) { dump.a = &tr0; dump.b = &tr1; dump.c = &tr2; dump.base = c->devs[0].run_start; dump.panels = &mine; }
    for (int k = 0; k < K + lookahead; k++) {
        const int h = k - lookahead;  // panel to hand to the sink
        if (h >= 0) {
            Device& d = c->devs[h % ndev];
            Slot& s = d.slot[(h / ndev) % NS];
            CUDA_CHECK(cudaSetDevice(d.id));
            CUDA_CHECK(cudaEventSynchronize(device_only ? s.k_stop : s.copied));
            harvest_kernel_time(c, s);
            const Panel& p = mine[h];
            c->tm.pairs += p.n_results;
            if (!device_only) {
                dg_panel desc{};
                desc.mode = mode;
                desc.row_begin = p.row0;
                desc.row_end = p.row1;
                desc.n_cols = mode == DG_MODE_SQUARE ? A0.n : B0.n;
                desc.n_results = p.n_results;
                finish_panel_desc(c, d, s, desc);
                if (sink(user, &desc) != 0) {
                    for (auto& dd : c->devs) { cudaSetDevice(dd.id); cudaDeviceSynchronize(); }
                    fail(DG_ERR_SINK, "sink aborted the run at panel %d", h);
                }
            }
        }
        if (k < K) {
            Device& d = c->devs[k % ndev];
            const int si = (k / ndev) % NS;
            Slot& s = d.slot[si];
            CUDA_CHECK(cudaSetDevice(d.id));
            const Panel& p = mine[k];
            if (repack_descending) {
                uint64_t& lo = packed_lo[k % ndev];
                if (p.row0 < lo) {
                    // this panel's own rows get every plane; the records between it and what is already packed (other
                    // ranks' panels) only serve as columns here: B-side planes (chunk boundaries: multiples of 128 records)
                    PlaneSet& ps = d.set[0];
                    const bool cnt = !ps.acgt_from_host && c->fam == FAM_TN93;
                    const uint64_t own_end = std::min<uint64_t>(lo, (p.row1 + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN);
                    enqueue_tc_pack(c, ps, ps.codes, own_end - p.row0, ps.input_kind, cnt, d.prep, p.row0);
                    if (own_end < lo)
                        enqueue_tc_pack(c, ps, ps.codes, lo - own_end, ps.input_kind, cnt, d.prep, own_end, nullptr, false, nullptr, true);
                    CUDA_CHECK(cudaEventRecord(d.sq_ready, d.prep));
                    lo = p.row0;
                }
                CUDA_CHECK(cudaStreamWaitEvent(d.cs(si), d.sq_ready, 0));
            }
            CUDA_CHECK(cudaEventRecord(s.k_start, d.cs(si)));
            if (trace) CUDA_CHECK(cudaEventRecord(tr0[k], d.cs(si)));
            cudaStream_t last = d.cs(si);
            s.kind = -1;
            if (tc_run) last = enqueue_panel_tc(c, d, d.set[0], d.set[wb], mode, p, s.d_out, s.d_scratch, false, false, d.cs(si), false, &s,
                                                trace ? tr1[k] : nullptr, use_post ? d.post : nullptr);
            else enqueue_panel_kernel<false>(c, d, d.set[0], d.set[wb], mode, p, s.d_out, false, d.cs(si));
            if (c->u8() && !device_only) enqueue_narrow(c, s, p.n_results, last);
            CUDA_CHECK(cudaEventRecord(s.k_stop, last));
            if (trace) CUDA_CHECK(cudaEventRecord(tr2[k], last));
            if (!device_only) {
                CUDA_CHECK(cudaStreamWaitEvent(d.copy, s.k_stop, 0));
                enqueue_panel_d2h(c, d, s, p.n_results);
                CUDA_CHECK(cudaEventRecord(s.copied, d.copy));
            }
        }
    }
}

// Rebuild the classic per-site index of a resident alignment that was loaded by a dg_square_* session
// (the session keeps a chunked index of its own; see sq_pump).
void ensure_pp_index(dg_ctx* c, Device& d, PlaneSet& s) {
    if (!s.pp_stale) return;
    s.pp_stale = false;
    if (!s.tc_ready || !tc_schedule(c->fam).needs_pp) return;
    CUDA_CHECK(cudaSetDevice(d.id));
    CUDA_CHECK(cudaMemsetAsync(d.pp_cnt, 0, (size_t)c->width * 4, d.compute));
    const unsigned gb = (unsigned)std::min<uint64_t>((s.n * c->width + 255) / 256, 148 * 32);
    tc::pp_count_kernel<<<gb, 256, 0, d.compute>>>(s.codes, s.n, c->width, s.input_kind == DG_INPUT_ASCII, d.pp_cnt);
    CUDA_CHECK(cudaGetLastError());
    c->tm.pack_launches++;
    finish_pp_index(c, d, s, d.compute);
}

// ---- pipelined all-vs-all session (dg_square_*) -------------------------------------------------------
// The caller pushes the alignment in chunks, HIGHEST records first.  Row i of the upper triangle needs the
// records j > i only, so once the records [lo, n) have landed every panel whose rows start at or above lo can
// run: upload (PCIe H2D), packing + tiles, and the D2H of finished panels all overlap, and panels reach the sink
// in completion order (descending rows).  The both-partial repair of n / n_high / raw / jc69 uses a chunked
// index (one per-site offset table per chunk over one shared entry buffer) that grows with the chunks.
void ensure_pipe_ring(dg_ctx* c, Device& d, size_t bytes, size_t scratch_bytes) {
    if (d.pout_cap < bytes) {
        for (auto& s : d.pslot) {
            if (s.d_out) cudaFree(s.d_out);
            if (s.h_out) cudaFreeHost(s.h_out);
            s.d_out = s.h_out = nullptr;
        }
        d.pout_cap = 0;
        for (auto& s : d.pslot) {
            CUDA_CHECK(cudaMalloc(&s.d_out, bytes));
            CUDA_CHECK(cudaHostAlloc(&s.h_out, bytes, cudaHostAllocDefault));
        }
        d.pout_cap = bytes;
    }
    (void)c;
    if (scratch_bytes)
        for (auto& s : d.pslot) ensure_scratch(s, scratch_bytes);
}

void sq_abort(dg_ctx* c) {
    for (auto& d : c->devs) {
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        cudaMemset(d.d_invalid, 0xff, 3 * sizeof(unsigned long long));
        for (int i = 0; i < 3; i++) d.h_invalid[i] = ~0ull;
        for (void* q : d.sq_retired) cudaFree(q);
        d.sq_retired.clear();
        d.set[0].n = 0; d.set[0].tc_ready = false; d.set[0].lop3_ready = false;
    }
    c->sq_queue.clear();
    c->sq_open = false;
}

void sq_sink_front(dg_ctx* c) {
    InFlight f = c->sq_queue.front();
    Device& d = c->devs[0];
    Slot& s = d.pslot[f.slot];
    CUDA_CHECK(cudaEventSynchronize(s.copied));
    harvest_kernel_time(c, s);
    c->tm.pairs += f.pairs;
    finish_panel_desc(c, d, s, f.desc);
    c->sq_queue.erase(c->sq_queue.begin());
    if (c->sq_sink(c->sq_user, &f.desc) != 0) fail(DG_ERR_SINK, "sink aborted the session");
}

void sq_report_invalid(dg_ctx* c, Device& d, unsigned long long key) {
    const unsigned long long reset = ~0ull;
    CUDA_CHECK(cudaMemcpy(d.d_invalid + 2, &reset, sizeof reset, cudaMemcpyHostToDevice));
    c->have_invalid = true;
    c->inv_record = key >> 32;
    c->inv_site = key & 0xffffffffull;
    c->inv_byte = 0;
    if (c->sq_nibble) CUDA_CHECK(cudaStreamSynchronize(d.unpack));   // the unpacked copy of that chunk
    CUDA_CHECK(cudaMemcpy(&c->inv_byte, d.set[0].codes + c->inv_record * c->width + c->inv_site, 1, cudaMemcpyDeviceToHost));
    fail(DG_ERR_INVALID_CODE, "invalid nucleotide byte 0x%02x in record %llu at site %llu", c->inv_byte,
         (unsigned long long)c->inv_record, (unsigned long long)c->inv_site);
}

void sq_launch_panel(dg_ctx* c, size_t k) {
    Device& d = c->devs[0];
    PlaneSet& S = d.set[0];
    const bool rect = c->sq_mode == DG_MODE_RECT;
    PlaneSet& B = rect ? d.set[1] : S;
    while ((int)c->sq_queue.size() >= Device::NPS) sq_sink_front(c);
    const int si = (int)(c->sq_launched % Device::NPS);
    c->sq_launched++;
    Slot& s = d.pslot[si];
    const Panel& p = c->sq_panels[k];
    cudaStream_t st = d.cs(si & 1);
    CUDA_CHECK(cudaStreamWaitEvent(st, d.sq_ready, 0));
    CUDA_CHECK(cudaEventRecord(s.k_start, st));
    const size_t li = c->sq_trace_panel.size();
    if (c->sq_trace) {
        c->sq_trace_panel.push_back((int)k);
        for (auto* v : {&d.tr_k0, &d.tr_k1, &d.tr_d2h})
            while (v->size() <= li) { cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); v->push_back(e); }
        CUDA_CHECK(cudaEventRecord(d.tr_k0[li], st));
    }
    s.kind = -1;
    const uint32_t n_entries = c->sq_base[c->sq_pumped];   // entries of every chunk pumped so far
    const bool repair = c->sq_needs_pp && n_entries != 0 && (!rect || B.pp.n_entries != 0);
    if (c->fam == FAM_SNP && !repair) {
        launch_tc_gemm(c, d, S, B, c->sq_mode, p, direct_plan(c), s.d_out, st, &s, true);
    } else {
        const AccPlan plan = scratch_plan(c, S.tc_fp4, c->sq_mode, p, B.n, repair);
        launch_tc_gemm(c, d, S, B, c->sq_mode, p, plan, s.d_scratch, st, &s, true);
        static const bool use_post = std::getenv("DG_POST_STREAM") != nullptr;
        if (use_post) {   // experiment (see run_panel_list): repair + combine on the post stream: the next panel's GEMM does not queue behind them
            CUDA_CHECK(cudaEventRecord(s.gemm_done, st));
            CUDA_CHECK(cudaStreamWaitEvent(d.post, s.gemm_done, 0));
            st = d.post;
        }
        if (repair && rect) {
            // rows' entries = the chunks that overlap the panel's rows (contiguous in the shared entry buffer),
            // columns = the classic per-site index of the resident alignment 1
            size_t ga0 = c->sq_pumped, ga1 = 0;
            for (size_t g = 0; g < c->sq_pumped; g++)
                if (c->sq_chunks[g].lo < p.row1 && c->sq_chunks[g].hi > p.row0) { ga0 = std::min(ga0, g); ga1 = std::max(ga1, g + 1); }
            if (ga0 < ga1 && c->sq_base[ga1] > c->sq_base[ga0]) {
                tc::PpCorrParams cp{};
                cp.a_entries = d.sq_entries + c->sq_base[ga0]; cp.a_n = c->sq_base[ga1] - c->sq_base[ga0];
                cp.b_entries = B.pp.entries; cp.b_off = B.pp.site_off;
                cp.row0 = (uint32_t)p.row0; cp.row_end = (uint32_t)p.row1; cp.n_b = (uint32_t)B.n;
                cp.square = 0; cp.s_pitch = plan.s_pitch; cp.s_colbase = plan.s_colbase; cp.op = plan.op[0]; cp.out = s.d_scratch;
                tc::pp_correct_kernel<<<(cp.a_n + 255) / 256, 256, 0, st>>>(cp);
                CUDA_CHECK(cudaGetLastError());
                c->tm.count_launches++;
            }
        } else if (repair) {
            // the panel's rows live in the chunks [ga0, ga1): chunks are descending, find those that overlap the rows
            size_t ga0 = c->sq_pumped, ga1 = 0;
            for (size_t g = 0; g < c->sq_pumped; g++)
                if (c->sq_chunks[g].lo < p.row1 && c->sq_chunks[g].hi > p.row0) { ga0 = std::min(ga0, g); ga1 = std::max(ga1, g + 1); }
            tc::PpChunkParams cp{};
            cp.entries = d.sq_entries; cp.off = d.sq_off; cp.off_stride = (uint32_t)(c->width + 1);
            cp.n_chunks = (uint32_t)c->sq_pumped;
            cp.a_begin = ga0 < ga1 ? c->sq_base[ga0] : 0; cp.a_end = ga0 < ga1 ? c->sq_base[ga1] : 0;
            cp.row0 = (uint32_t)p.row0; cp.row_end = (uint32_t)p.row1;
            cp.s_pitch = plan.s_pitch; cp.s_colbase = plan.s_colbase; cp.op = plan.op[0]; cp.out = s.d_scratch;
            if (cp.a_end > cp.a_begin) {
                const unsigned gb = (unsigned)std::min<uint32_t>((cp.a_end - cp.a_begin + 127) / 128, 148 * 8);
                tc::pp_correct_chunks_kernel<<<gb, 128, 0, st>>>(cp);
                CUDA_CHECK(cudaGetLastError());
                c->tm.count_launches++;
            }
        }
        s.kind = c->counts16() ? DG_RESULT_COUNTS16 : -1;
        launch_combine(c, d, S, B, c->sq_mode, p, s.d_out, s.d_scratch, plan, false, false, st, c->counts16());
    }
    if (c->u8()) enqueue_narrow(c, s, p.n_results, st);
    CUDA_CHECK(cudaEventRecord(s.k_stop, st));
    if (c->sq_trace) CUDA_CHECK(cudaEventRecord(d.tr_k1[li], st));
    CUDA_CHECK(cudaStreamWaitEvent(d.copy, s.k_stop, 0));
    enqueue_panel_d2h(c, d, s, p.n_results);
    CUDA_CHECK(cudaEventRecord(s.copied, d.copy));
    if (c->sq_trace) CUDA_CHECK(cudaEventRecord(d.tr_d2h[li], d.copy));
    InFlight f;
    f.dev = 0; f.slot = si; f.pairs = p.n_results;
    f.desc.mode = c->sq_mode;
    f.desc.result_kind = c->result_kind();
    f.desc.row_begin = p.row0; f.desc.row_end = p.row1;
    f.desc.n_cols = B.n; f.desc.n_results = p.n_results;
    c->sq_queue.push_back(f);
}

// Chunk g has been pushed (its packing and index scan are enqueued): finish its part of the index and launch
// every panel whose rows now have all their columns on the device.
void sq_pump(dg_ctx* c, size_t g) {
    Device& d = c->devs[0];
    PlaneSet& S = d.set[0];
    const auto& ch = c->sq_chunks[g];
    CUDA_CHECK(cudaEventSynchronize(d.sq_ev[g]));
    if (d.h_sq_invalid[g] != ~0ull) sq_report_invalid(c, d, d.h_sq_invalid[g]);
    uint32_t total = c->sq_base[g];
    if (c->sq_tc && c->sq_needs_pp) {
        total = d.h_sq_total[g];
        if (total > d.sq_entries_cap) {   // grow the shared entry buffer (old one is freed when the session ends)
            const uint32_t cap = std::max<uint32_t>(total + total / 2, 1u << 20);
            uint64_t* fresh = nullptr;
            CUDA_CHECK(cudaMalloc(&fresh, (size_t)cap * 8));
            if (d.sq_entries) {
                CUDA_CHECK(cudaMemcpyAsync(fresh, d.sq_entries, (size_t)c->sq_base[g] * 8, cudaMemcpyDeviceToDevice, d.fill));
                d.sq_retired.push_back(d.sq_entries);
            }
            d.sq_entries = fresh; d.sq_entries_cap = cap;
        }
        uint32_t* cursor = d.sq_off + g * (size_t)(c->width + 1) + 1;
        if (total > c->sq_base[g] && total <= d.pp_hit_cap) {
            // the pack kernel appended this chunk's partial codes at hits[base_g, total): the running hit count is the
            // running entry count
            tc::pp_scatter_kernel<<<(unsigned)std::min<uint32_t>((total - c->sq_base[g] + 255) / 256, 148 * 4), 256, 0, d.fill>>>(
                d.pp_hits, c->sq_base[g], total, cursor, d.sq_entries);
            CUDA_CHECK(cudaGetLastError());
            c->tm.pack_launches++;
        } else if (total > c->sq_base[g]) {   // hit buffer overflowed: rescan the chunk's codes
            const uint64_t nr = ch.hi - ch.lo;
            const unsigned gb = (unsigned)std::min<uint64_t>((nr * c->width + 255) / 256, 148 * 32);
            if (c->sq_nibble && c->sq_tc) CUDA_CHECK(cudaStreamWaitEvent(d.fill, d.unpack_ev[g & 31], 0));   // the unpacked copy
            tc::pp_fill_kernel<<<gb, 256, 0, d.fill>>>(S.codes + ch.lo * c->width, nr, c->width, c->sq_input_kind == DG_INPUT_ASCII,
                                                      cursor, d.sq_entries, ch.lo);
            CUDA_CHECK(cudaGetLastError());
            c->tm.pack_launches++;
        }
        // auto engine: an alignment so full of partial codes that the repair would cost more than the tiles finishes
        // on the LOP3 engine (same rule as use_tc)
        if (c->sq_mode == DG_MODE_RECT) {
            const PlaneSet& Bs = d.set[1];
            const double m = (double)ch.hi;
            if (c->engine == 0 && std::sqrt(d.h_sq_work[g] * Bs.pp.pair_work) > c->pp_budget() * m * (double)Bs.n && m >= 1024) c->sq_fallback = true;
        } else {
            const double m = (double)(c->sq_n - ch.lo);
            if (c->engine == 0 && d.h_sq_work[g] > c->pp_budget() * m * m && m >= 1024) c->sq_fallback = true;
        }
    }
    c->sq_base[g + 1] = total;
    c->sq_pumped = g + 1;
    CUDA_CHECK(cudaEventRecord(d.sq_ready, d.fill));   // the host has waited for chunk g's scan: nothing else to order
    if (!c->sq_tc || c->sq_fallback) return;
    // SQUARE: a panel needs every record at or above its first row; RECT: its own rows (alignment 1 is resident)
    while (c->sq_next < c->sq_panels.size() &&
           (c->sq_mode == DG_MODE_RECT ? c->sq_panels[c->sq_next].row1 <= ch.hi : c->sq_panels[c->sq_next].row0 >= ch.lo))
        sq_launch_panel(c, c->sq_next++);
}

void sq_begin(dg_ctx* c, int mode, uint64_t n, int input_kind, const uint64_t* acgt_counts, uint32_t part, uint32_t n_parts,
              dg_sink_fn sink, void* user) {
    if (c->streaming || c->sq_open) fail(DG_ERR_STATE, "a session is already open");
    if (c->devs.size() != 1) fail(DG_ERR_STATE, "dg_square_* sessions drive one device per context (one process per GPU)");
    if (!sink) fail(DG_ERR_INVALID_ARG, "sink is NULL");
    if (n == 0 || n >= (1ull << 31)) fail(DG_ERR_INVALID_ARG, "bad record count");
    if (n_parts == 0 || part >= n_parts) fail(DG_ERR_INVALID_ARG, "bad part %u of %u", part, n_parts);
    if (!valid_input_kind(input_kind)) fail(DG_ERR_INVALID_ARG, "bad input_kind");
    Device& d = c->devs[0];
    CUDA_CHECK(cudaSetDevice(d.id));
    c->have_invalid = false;
    PlaneSet& S = d.set[0];
    const bool want_tc = c->tc_wanted();
    const bool rect = mode == DG_MODE_RECT;
    if (rect) {
        if (d.set[1].n == 0) fail(DG_ERR_STATE, "alignment 1 is not loaded (dg_load_resident(ctx, 1, ...) before dg_rect_begin)");
        if (want_tc && (!d.set[1].tc_ready || d.set[1].tc_fp4 != c->want_fp4()))
            fail(DG_ERR_STATE, "alignment 1 was loaded for another engine (set DG_OPT_ENGINE before loading)");
        ensure_pp_index(c, d, d.set[1]);
    }
    c->sq_mode = mode;
    reserve_resident(c, S, n, want_tc);
    if (input_kind == DG_INPUT_NIBBLE) ensure_nib(c, S, n);
    S.input_kind = device_kind(input_kind);
    S.acgt_from_host = acgt_counts != nullptr;
    S.pp_stale = true;
    c->sq_tc = want_tc;
    c->sq_fallback = false;
    c->sq_needs_pp = tc_schedule(c->fam).needs_pp;
    c->sq_n = n;
    c->sq_input_kind = device_kind(input_kind);
    c->sq_nibble = input_kind == DG_INPUT_NIBBLE;
    c->sq_sink = sink; c->sq_user = user;
    c->sq_queue.clear();
    c->sq_launched = 0;
    c->last_engine = want_tc ? (c->want_fp4() ? 3 : 2) : 1;
    if (acgt_counts) {
        std::vector<uint32_t> c32(n * 4);
        for (uint64_t i = 0; i < n * 4; i++) c32[i] = (uint32_t)acgt_counts[i];
        CUDA_CHECK(cudaMemsetAsync(S.acgt, 0, (size_t)S.n_pad * 16, d.compute));
        CUDA_CHECK(cudaMemcpyAsync(S.acgt, c32.data(), (size_t)n * 16, cudaMemcpyHostToDevice, d.compute));
        CUDA_CHECK(cudaStreamSynchronize(d.compute));
    }
    // panels: the whole triangle in ~pipe_panels x n_parts pieces (small panels keep the D2H stream fed from early on)
    const TileShape ts = tile_shape(c->fam, c->tile_variant);
    const uint64_t n_cols = rect ? d.set[1].n : n;
    const uint64_t total_bytes = (rect ? n * n_cols : n * (n - 1) / 2) * c->elem_bytes();
    const size_t pb = (size_t)std::min<uint64_t>(c->panel_bytes,
                          std::max<uint64_t>(8ull << 20, total_bytes / ((uint64_t)std::max(1, c->pipe_panels) * n_parts)));  // DG_OPT_PANEL_BYTES caps it
    std::vector<Panel> all = make_panels(pb, c->elem_bytes(), mode, n, n_cols, ts.tm, c->plan_tn(), c->plan_items());
    c->sq_panels = panels_of_part(all, part, n_parts);
    if (!rect) std::reverse(c->sq_panels.begin(), c->sq_panels.end());   // launch order: descending rows
    c->sq_next = 0;
    size_t max_bytes = 256;
    for (auto& p : c->sq_panels) max_bytes = std::max(max_bytes, (size_t)p.n_results * c->elem_bytes());
    if (want_tc && c->u8())
        for (auto& sl : d.pslot) ensure_narrow(sl, max_bytes / c->elem_bytes());
    if (want_tc) {
        size_t sb = 0;
        if (c->fam != FAM_SNP || c->sq_needs_pp)
            for (auto& p : c->sq_panels) sb = std::max(sb, scratch_upper_bound(c, c->want_fp4(), mode, p.row0, p.row1, n_cols));
        ensure_pipe_ring(c, d, max_bytes, sb);
    }
    // chunks: the same for every part (a multi-rank launcher broadcasts them): <= ~40 pieces of >= 24 MB, descending,
    // every boundary but n itself a multiple of ROW_ALIGN (the pack kernel zero-fills whole 128-row groups)
    const uint64_t target = c->pipe_chunk_bytes ? c->pipe_chunk_bytes : std::max<uint64_t>(24ull << 20, n * c->width / 40);
    // the target is bytes ON THE WIRE: nibble rows carry twice the records per chunk (the per-chunk unpack / pack / scan
    // launches are what falls behind the copies once GEMM CTAs share the SMs: profiles/e2e_timeline_r02.md)
    const uint64_t rows = std::max<uint64_t>(ROW_ALIGN, target / input_stride(c, input_kind) / ROW_ALIGN * ROW_ALIGN);
    c->sq_chunks.clear();
    if (rect) {
        for (uint64_t lo = 0; lo < n;) {   // ascending
            uint64_t hi = std::min<uint64_t>(n, lo + rows);
            if (n - hi < ROW_ALIGN * 2) hi = n;   // no sliver at the top
            c->sq_chunks.push_back({lo, hi});
            lo = hi;
        }
    } else {
        for (uint64_t hi = n; hi > 0;) {
            uint64_t lo = hi > rows ? (hi - rows) / ROW_ALIGN * ROW_ALIGN : 0;
            if (lo < ROW_ALIGN * 2) lo = 0;   // no sliver at the bottom
            c->sq_chunks.push_back({lo, hi});
            hi = lo;
        }
    }
    const size_t G = c->sq_chunks.size();
    c->sq_pushed = c->sq_pumped = 0;
    c->sq_base.assign(G + 1, 0);
    if (d.sq_ev.size() < G) {
        const size_t old = d.sq_ev.size();
        d.sq_ev.resize(G, nullptr);
        for (size_t g = old; g < G; g++) CUDA_CHECK(cudaEventCreateWithFlags(&d.sq_ev[g], cudaEventDisableTiming));
    }
    if (d.sq_off_chunks < G) {
        if (d.sq_off) cudaFree(d.sq_off);
        if (d.h_sq_total) cudaFreeHost(d.h_sq_total);
        if (d.h_sq_work) cudaFreeHost(d.h_sq_work);
        if (d.h_sq_invalid) cudaFreeHost(d.h_sq_invalid);
        d.sq_off = nullptr; d.h_sq_total = nullptr; d.h_sq_work = nullptr; d.h_sq_invalid = nullptr; d.sq_off_chunks = 0;
        CUDA_CHECK(cudaMalloc(&d.sq_off, G * (size_t)(c->width + 1) * 4));
        CUDA_CHECK(cudaHostAlloc(&d.h_sq_total, G * 4, cudaHostAllocMapped));   // written by kernels (see pp_scan_chunk_kernel)
        CUDA_CHECK(cudaHostAlloc(&d.h_sq_work, G * 8, cudaHostAllocMapped));
        CUDA_CHECK(cudaHostAlloc(&d.h_sq_invalid, G * 8, cudaHostAllocMapped));
        d.sq_off_chunks = G;
    }
    if (!d.sq_cum) CUDA_CHECK(cudaMalloc(&d.sq_cum, (size_t)c->width * 4));
    if (!d.sq_total) CUDA_CHECK(cudaMalloc(&d.sq_total, 4));
    CUDA_CHECK(cudaMemsetAsync(d.sq_cum, 0, (size_t)c->width * 4, d.prep));
    CUDA_CHECK(cudaMemsetAsync(d.sq_total, 0, 4, d.prep));
    if (want_tc && c->sq_needs_pp) {
        ensure_pp_hits(c, d, n);
        CUDA_CHECK(cudaMemsetAsync(d.pp_hit_count, 0, 4, d.prep));
    }
    c->sq_trace = std::getenv("DG_TRACE") != nullptr;
    c->sq_trace_panel.clear();
    if (c->sq_trace) {
        if (!d.tr_base) CUDA_CHECK(cudaEventCreate(&d.tr_base));
        for (auto* v : {&d.tr_copy, &d.tr_prep})
            while (v->size() < G) { cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); v->push_back(e); }
        CUDA_CHECK(cudaEventRecord(d.tr_base, d.copy_in));
    }
    c->sq_t0 = wall_ms();
    c->sq_open = true;
}

void sq_push(dg_ctx* c, const uint8_t* codes, int src_dev, uint64_t lo, uint64_t hi, void* ready_event) {
    if (!c->sq_open) fail(DG_ERR_STATE, "no dg_square session is open");
    if (c->sq_pushed >= c->sq_chunks.size()) fail(DG_ERR_STATE, "every chunk of the session has been pushed");
    const auto ch = c->sq_chunks[c->sq_pushed];
    if (lo != ch.lo || hi != ch.hi)
        fail(DG_ERR_INVALID_ARG, "the session expects records [%llu, %llu) next (dg_square_next), got [%llu, %llu)",
             (unsigned long long)ch.lo, (unsigned long long)ch.hi, (unsigned long long)lo, (unsigned long long)hi);
    if (!codes) fail(DG_ERR_INVALID_ARG, "codes is NULL");
    Device& d = c->devs[0];
    PlaneSet& S = d.set[0];
    CUDA_CHECK(cudaSetDevice(d.id));
    const size_t g = c->sq_pushed;
    const uint64_t nr = hi - lo;
    const uint64_t stride = input_stride(c, c->sq_nibble ? DG_INPUT_NIBBLE : DG_INPUT_PARADIS);
    uint8_t* dst = c->sq_nibble ? S.nib + lo * stride : S.codes + lo * stride;
    if (ready_event) CUDA_CHECK(cudaStreamWaitEvent(d.copy_in, (cudaEvent_t)ready_event, 0));
    if (src_dev < 0) {
        CUDA_CHECK(cudaMemcpyAsync(dst, codes, (size_t)nr * stride, cudaMemcpyHostToDevice, d.copy_in));
        c->tm.h2d_bytes += nr * stride;
    } else {
        CUDA_CHECK(cudaMemcpyPeerAsync(dst, d.id, codes, src_dev, (size_t)nr * stride, d.copy_in));
    }
    cudaEvent_t ev = d.chunk_ev[g & 31];
    CUDA_CHECK(cudaEventRecord(ev, d.copy_in));
    if (c->sq_trace) CUDA_CHECK(cudaEventRecord(d.tr_copy[g], d.copy_in));
    // packing + index scan of this chunk: enqueued now, runs as soon as the copy has landed
    CUDA_CHECK(cudaStreamWaitEvent(d.prep, ev, 0));
    // Nibble rows: the tensor-core pack reads them directly, so the unpacked copy (kept for what runs on the resident
    // alignment later: the LOP3 engine, a rescan of the partial codes, re-packing) is written off the critical path.
    const bool fused_nibble = c->sq_nibble && c->sq_tc;
    if (fused_nibble) {
        CUDA_CHECK(cudaStreamWaitEvent(d.unpack, ev, 0));
        enqueue_nibble_unpack(c, dst, S.codes + lo * c->width, nr, d.unpack);
        CUDA_CHECK(cudaEventRecord(d.unpack_ev[g & 31], d.unpack));
    } else if (c->sq_nibble) {
        enqueue_nibble_unpack(c, dst, S.codes + lo * c->width, nr, d.prep);
    }
    if (c->sq_tc) {
        if (c->sq_needs_pp) CUDA_CHECK(cudaMemsetAsync(d.pp_cnt, 0, (size_t)c->width * 4, d.prep));
        enqueue_tc_pack(c, S, fused_nibble ? S.nib : S.codes, nr, c->sq_input_kind, !S.acgt_from_host && c->fam == FAM_TN93, d.prep, lo,
                        d.d_invalid + 2, false, c->sq_needs_pp ? &d : nullptr, false, fused_nibble);
        if (c->sq_needs_pp) {
            tc::pp_scan_chunk_kernel<<<1, 1024, 0, d.prep>>>(d.pp_cnt, d.sq_cum, c->width, d.sq_off + g * (size_t)(c->width + 1),
                                                             d.sq_total, d.h_sq_total + g, d.h_sq_work + g, d.d_invalid + 2,
                                                             d.h_sq_invalid + g);
            CUDA_CHECK(cudaGetLastError());
            c->tm.pack_launches++;
        }
    }
    if (!(c->sq_tc && c->sq_needs_pp)) {
        tc::publish_invalid_kernel<<<1, 1, 0, d.prep>>>(d.d_invalid + 2, d.h_sq_invalid + g);
        CUDA_CHECK(cudaGetLastError());
    }
    CUDA_CHECK(cudaEventRecord(d.sq_ev[g], d.prep));
    if (c->sq_trace) CUDA_CHECK(cudaEventRecord(d.tr_prep[g], d.prep));
    c->sq_pushed = g + 1;
    // Finish chunk g - LOOKAHEAD (a short host wait: its copy landed a few chunks ago) and launch the panels it
    // completes; the copies queued meanwhile keep PCIe busy.
    if (c->sq_pushed > (size_t)DG_SQUARE_LOOKAHEAD) sq_pump(c, c->sq_pushed - 1 - DG_SQUARE_LOOKAHEAD);
}

void sq_end(dg_ctx* c) {
    if (!c->sq_open) fail(DG_ERR_STATE, "no dg_square session is open");
    if (c->sq_pushed != c->sq_chunks.size())
        fail(DG_ERR_STATE, "dg_square_end: %zu of %zu chunks were pushed", c->sq_pushed, c->sq_chunks.size());
    Device& d = c->devs[0];
    PlaneSet& S = d.set[0];
    CUDA_CHECK(cudaSetDevice(d.id));
    while (c->sq_pumped < c->sq_pushed) sq_pump(c, c->sq_pumped);
    while (!c->sq_queue.empty()) sq_sink_front(c);
    if (c->sq_nibble) CUDA_CHECK(cudaStreamSynchronize(d.unpack));   // the resident code bytes are complete from here on
    if (c->sq_trace) {
        CUDA_CHECK(cudaDeviceSynchronize());
        auto at = [&](cudaEvent_t e) { float ms = 0; cudaEventElapsedTime(&ms, d.tr_base, e); return ms; };
        fprintf(stderr, "[dg_square] %zu chunks, %zu panels launched, wall %.3f ms\n", c->sq_chunks.size(),
                c->sq_trace_panel.size(), wall_ms() - c->sq_t0);
        for (size_t g = 0; g < c->sq_chunks.size(); g++)
            fprintf(stderr, "  chunk %2zu [%6llu,%6llu) copied %7.3f  prepped %7.3f\n", g, (unsigned long long)c->sq_chunks[g].lo,
                    (unsigned long long)c->sq_chunks[g].hi, at(d.tr_copy[g]), at(d.tr_prep[g]));
        for (size_t i = 0; i < c->sq_trace_panel.size(); i++) {
            const Panel& p = c->sq_panels[c->sq_trace_panel[i]];
            fprintf(stderr, "  panel %2d rows [%6llu,%6llu) %9llu pairs  tiles %7.3f - %7.3f  d2h done %7.3f\n", c->sq_trace_panel[i],
                    (unsigned long long)p.row0, (unsigned long long)p.row1, (unsigned long long)p.n_results, at(d.tr_k0[i]),
                    at(d.tr_k1[i]), at(d.tr_d2h[i]));
        }
    }
    S.tc_ready = c->sq_tc;
    if (c->sq_next < c->sq_panels.size()) {
        // LOP3 engine (forced, or chosen because the alignment is full of partial ambiguity codes): the remaining
        // panels run the classic way on the now-resident alignment
        CUDA_CHECK(cudaStreamSynchronize(d.prep));
        CUDA_CHECK(cudaStreamSynchronize(d.fill));
        std::vector<Panel> rest(c->sq_panels.begin() + c->sq_next, c->sq_panels.end());
        c->sq_next = c->sq_panels.size();
        PlaneSet& Bs = c->sq_mode == DG_MODE_RECT ? d.set[1] : S;
        if (c->sq_tc) ensure_pp_index(c, d, S);
        const bool tc_run = use_tc(c, S, Bs);
        if (!tc_run) {
            ensure_lop3(c, d, S);
            check_invalid(c, d, S.codes, 0);
            if (c->sq_mode == DG_MODE_RECT) ensure_lop3(c, d, Bs);
        }
        c->last_engine = tc_run ? (S.tc_fp4 ? 3 : 2) : 1;
        run_panel_list(c, c->sq_mode, rest, tc_run, c->sq_sink, c->sq_user, false);
    }
    CUDA_CHECK(cudaStreamSynchronize(d.prep));
    CUDA_CHECK(cudaStreamSynchronize(d.fill));
    for (void* q : d.sq_retired) cudaFree(q);
    d.sq_retired.clear();
    harvest_clock(c);
    c->tm.total_ms = wall_ms() - c->sq_t0;
    c->sq_open = false;
}

// ---- stream session ----------------------------------------------------------------------------
void stream_sink_front(dg_ctx* c) {
    InFlight f = c->s_queue.front();
    Device& d = c->devs[f.dev];
    Slot& s = d.slot[f.slot];
    CUDA_CHECK(cudaSetDevice(d.id));
    CUDA_CHECK(cudaEventSynchronize(s.copied));
    // an invalid byte in this batch?
    if (d.h_invalid[f.slot] != ~0ull) {
        const unsigned long long key = d.h_invalid[f.slot];
        c->have_invalid = true;
        c->inv_record = (key >> 32) + f.desc.row_begin;
        c->inv_site = key & 0xffffffffull;
        c->inv_byte = s.batch.input_kind == DG_INPUT_NIBBLE ? 0 : s.h_in[(key >> 32) * c->width + c->inv_site];
        fail(DG_ERR_INVALID_CODE, "invalid nucleotide byte 0x%02x in streamed record %llu at site %llu",
             c->inv_byte, (unsigned long long)c->inv_record, (unsigned long long)c->inv_site);
    }
    harvest_kernel_time(c, s);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.p_start, s.p_stop) == cudaSuccess) c->tm.pack_ms += ms;
    c->tm.pairs += f.pairs;
    c->tm.d2h_bytes += f.desc.n_results * c->elem_bytes();
    f.desc.data = s.h_out;
    c->s_queue.erase(c->s_queue.begin());
    if (c->s_sink(c->s_user, &f.desc) != 0) fail(DG_ERR_SINK, "sink aborted the stream");
}

void stream_begin(dg_ctx* c, dg_sink_fn sink, void* user, uint64_t max_batch) {
    if (c->streaming || c->sq_open) fail(DG_ERR_STATE, "a session is already open");
    if (!sink) fail(DG_ERR_INVALID_ARG, "sink is NULL");
    if (max_batch == 0) fail(DG_ERR_INVALID_ARG, "max_batch is 0");
    for (auto& d : c->devs)
        if (d.set[0].n == 0) fail(DG_ERR_STATE, "alignment 0 is not loaded");
    for (auto& d : c->devs) ensure_pp_index(c, d, d.set[0]);
    const uint64_t n_res = c->devs[0].set[0].n;
    // keep one batch's results within the panel budget
    const uint64_t cap_rows = std::max<uint64_t>(1, c->panel_bytes / (c->elem_bytes() * n_res));
    const TileShape ts = tile_shape(c->fam, c->tile_variant);
    uint64_t mb = std::min(max_batch, std::max<uint64_t>(cap_rows / ts.tm * ts.tm, ts.tm));
    c->s_max_batch = mb;
    // Streamed batches run on the tensor cores like resident panels (the both-partial repair of n / n_high /
    // raw / jc69 scans the batch's own V planes against the resident index) unless the resident alignment is
    // so full of partial ambiguity codes that the repair would dominate; then the LOP3 tiles run.
    const PlaneSet& R0 = c->devs[0].set[0];
    const bool cheap_pp = !tc_schedule(c->fam).needs_pp || R0.pp.pair_work <= c->pp_budget() * (double)R0.n * (double)R0.n;
    const bool s_tc = c->tc_wanted() && R0.tc_ready && R0.tc_fp4 == c->want_fp4() && (c->engine >= 2 || cheap_pp);
    if (!s_tc)
        for (auto& d : c->devs) ensure_lop3(c, d, d.set[0]);
    for (auto& d : c->devs) {
        CUDA_CHECK(cudaSetDevice(d.id));
        ensure_out_ring(c, d, (size_t)mb * n_res * c->elem_bytes());
        if (s_tc)
            for (int q = 0; q < 2; q++) ensure_scratch(d.slot[q], scratch_upper_bound(c, c->want_fp4(), DG_MODE_RECT, 0, mb, n_res));
        if (d.in_cap < mb || (s_tc && (!d.slot[0].batch.tc_ops || d.slot[0].batch.tc_fp4 != c->want_fp4())) ||
            (!s_tc && !d.slot[0].batch.core)) {
            for (int q = 0; q < 2; q++) {
                Slot& s = d.slot[q];
                if (s.h_in) cudaFreeHost(s.h_in);
                if (s.d_in) cudaFree(s.d_in);
                if (s.d_nib) cudaFree(s.d_nib);
                if (s.h_acgt) cudaFreeHost(s.h_acgt);
                free_set(s.batch);
                s.h_in = s.d_in = s.d_nib = nullptr; s.h_acgt = nullptr;
                CUDA_CHECK(cudaHostAlloc(&s.h_in, (size_t)mb * c->width, cudaHostAllocDefault));
                CUDA_CHECK(cudaMalloc(&s.d_in, (size_t)mb * c->width));
                CUDA_CHECK(cudaMalloc(&s.d_nib, (size_t)mb * input_stride(c, DG_INPUT_NIBBLE)));
                CUDA_CHECK(cudaHostAlloc(&s.h_acgt, (size_t)mb * 4 * sizeof(uint32_t), cudaHostAllocDefault));
                alloc_set(c, s.batch, mb, false, !s_tc);
                if (s_tc) alloc_tc_operands(c, s.batch);
            }
            d.in_cap = mb;
        }
    }
    c->streaming = true;
    c->s_tc = s_tc;
    c->last_engine = s_tc ? (R0.tc_fp4 ? 3 : 2) : 1;
    c->s_sink = sink;
    c->s_user = user;
    c->s_rows_pushed = 0;
    c->s_batches = 0;
    c->s_queue.clear();
    c->s_t0 = wall_ms();
}

void stream_push_one(dg_ctx* c, const uint8_t* codes, uint64_t nb, int input_kind, const uint64_t* acgt) {
    const int ndev = (int)c->devs.size();
    // the ring slot of this batch must have been sunk: keep at most 2*ndev batches in flight
    while ((int)c->s_queue.size() >= 2 * ndev) stream_sink_front(c);
    const uint64_t bi = c->s_batches++;
    Device& d = c->devs[bi % ndev];
    const int si = (int)((bi / ndev) & 1);
    Slot& s = d.slot[si];
    CUDA_CHECK(cudaSetDevice(d.id));
    const double th = wall_ms();
    const bool nibble = input_kind == DG_INPUT_NIBBLE;
    const uint64_t stride = input_stride(c, input_kind);
    input_kind = device_kind(input_kind);
    if (codes != s.h_in) std::memcpy(s.h_in, codes, (size_t)nb * stride);   // dg_stream_buffer: filled in place
    const bool host_counts = acgt != nullptr && c->fam == FAM_TN93;
    if (host_counts)
        for (uint64_t i = 0; i < nb * 4; i++) s.h_acgt[i] = (uint32_t)acgt[i];
    CUDA_CHECK(cudaMemcpyAsync(nibble ? s.d_nib : s.d_in, s.h_in, (size_t)nb * stride, cudaMemcpyHostToDevice, d.copy_in));
    if (host_counts)
        CUDA_CHECK(cudaMemcpyAsync(s.batch.acgt, s.h_acgt, (size_t)nb * 4 * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, d.copy_in));
    CUDA_CHECK(cudaEventRecord(s.in_ready, d.copy_in));
    c->tm.h2d_ms += wall_ms() - th;
    c->tm.h2d_bytes += nb * stride;
    cudaStream_t cst = d.cs(si);
    CUDA_CHECK(cudaStreamWaitEvent(cst, s.in_ready, 0));
    s.batch.n = nb;
    CUDA_CHECK(cudaEventRecord(s.p_start, cst));
    // tensor-core engine: the pack kernel reads the nibble rows itself (a batch's codes are not needed again: its partial
    // codes are found by scanning its V planes); the LOP3 engine packs from unpacked bytes
    if (nibble && !c->s_tc) enqueue_nibble_unpack(c, s.d_nib, s.d_in, nb, cst);
    s.batch.input_kind = nibble ? DG_INPUT_NIBBLE : input_kind;
    // fastaio.rs:250-254: the streamed tn93 records count raw upper-case chars only (:139-142)
    if (c->s_tc) enqueue_tc_pack(c, s.batch, nibble ? s.d_nib : s.d_in, nb, input_kind, !host_counts && c->fam == FAM_TN93, cst, 0,
                                 d.d_invalid + si, true, nullptr, false, nibble);
    else enqueue_pack(c, d.d_invalid + si, s.batch, s.d_in, nb, input_kind, !host_counts, input_kind == DG_INPUT_ASCII, cst);
    CUDA_CHECK(cudaEventRecord(s.p_stop, cst));
    Panel p;
    p.row0 = 0; p.row1 = nb; p.out_base = 0;
    p.n_results = nb * d.set[0].n;
    CUDA_CHECK(cudaEventRecord(s.k_start, cst));
    if (c->s_tc) enqueue_panel_tc(c, d, s.batch, d.set[0], DG_MODE_RECT, p, s.d_out, s.d_scratch, true, false, cst, true);
    else enqueue_panel_kernel<false>(c, d, s.batch, d.set[0], DG_MODE_RECT, p, s.d_out, true, cst);
    CUDA_CHECK(cudaEventRecord(s.k_stop, cst));
    CUDA_CHECK(cudaStreamWaitEvent(d.copy, s.k_stop, 0));
    CUDA_CHECK(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)p.n_results * c->elem_bytes(),
                               cudaMemcpyDeviceToHost, d.copy));
    CUDA_CHECK(cudaMemcpyAsync(d.h_invalid + si, d.d_invalid + si, sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, d.copy));
    CUDA_CHECK(cudaEventRecord(s.copied, d.copy));
    InFlight f;
    f.dev = (int)(bi % ndev);
    f.slot = si;
    f.pairs = p.n_results;
    f.desc.mode = DG_MODE_STREAM;
    f.desc.result_kind = c->result_kind();
    f.desc.row_begin = c->s_rows_pushed;
    f.desc.row_end = c->s_rows_pushed + nb;
    f.desc.n_cols = d.set[0].n;
    f.desc.n_results = p.n_results;
    c->s_queue.push_back(f);
    c->s_rows_pushed += nb;
}

void stream_abort(dg_ctx* c) {
    for (auto& d : c->devs) {
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        cudaMemset(d.d_invalid, 0xff, 3 * sizeof(unsigned long long));
        for (int i = 0; i < 3; i++) d.h_invalid[i] = ~0ull;
    }
    c->s_queue.clear();
    c->streaming = false;
}

template <typename F>
int guarded(dg_ctx* c, F&& f) {
    if (!c) return DG_ERR_INVALID_ARG;
    try {
        f();
        return DG_OK;
    } catch (const DgError& e) {
        c->err = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        c->err = e.what();
        return DG_ERR_NOMEM;
    }
}

void destroy_device(Device& d) {
    cudaSetDevice(d.id);
    cudaDeviceSynchronize();
    for (auto& s : d.set) free_set(s);
    for (auto& s : d.slot) {
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_scratch) cudaFree(s.d_scratch);
        if (s.d_tiles) cudaFree(s.d_tiles);
        if (s.h_tiles) cudaFreeHost(s.h_tiles);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_nib) cudaFree(s.d_nib);
        if (s.h_acgt) cudaFreeHost(s.h_acgt);
        free_narrow(s);
        free_set(s.batch);
        for (cudaEvent_t e : {s.k_start, s.k_stop, s.copied, s.in_ready, s.p_start, s.p_stop, s.gemm_done})
            if (e) cudaEventDestroy(e);
    }
    for (auto& s : d.pslot) {
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_scratch) cudaFree(s.d_scratch);
        if (s.d_tiles) cudaFree(s.d_tiles);
        if (s.h_tiles) cudaFreeHost(s.h_tiles);
        if (s.h_out) cudaFreeHost(s.h_out);
        free_narrow(s);
        for (cudaEvent_t e : {s.k_start, s.k_stop, s.copied, s.gemm_done})
            if (e) cudaEventDestroy(e);
    }
    if (d.sq_off) cudaFree(d.sq_off);
    if (d.sq_cum) cudaFree(d.sq_cum);
    if (d.sq_total) cudaFree(d.sq_total);
    if (d.sq_entries) cudaFree(d.sq_entries);
    for (void* q : d.sq_retired) cudaFree(q);
    if (d.h_sq_total) cudaFreeHost(d.h_sq_total);
    if (d.h_sq_work) cudaFreeHost(d.h_sq_work);
    if (d.h_sq_invalid) cudaFreeHost(d.h_sq_invalid);
    for (auto& e : d.sq_ev) if (e) cudaEventDestroy(e);
    for (auto* v : {&d.tr_copy, &d.tr_prep, &d.tr_k0, &d.tr_k1, &d.tr_d2h})
        for (auto& e : *v) if (e) cudaEventDestroy(e);
    if (d.tr_base) cudaEventDestroy(d.tr_base);
    if (d.prep) cudaStreamDestroy(d.prep);
    if (d.fill) cudaStreamDestroy(d.fill);
    if (d.sq_ready) cudaEventDestroy(d.sq_ready);
    if (d.pp_cnt) cudaFree(d.pp_cnt);
    if (d.pp_hits) cudaFree(d.pp_hits);
    if (d.pp_hit_count) cudaFree(d.pp_hit_count);
    if (d.pp_cursor) cudaFree(d.pp_cursor);
    if (d.pp_work) cudaFree(d.pp_work);
    if (d.h_pp_total) cudaFreeHost(d.h_pp_total);
    if (d.h_pp_work) cudaFreeHost(d.h_pp_work);
    for (auto& e : d.chunk_ev) if (e) cudaEventDestroy(e);
    for (auto& e : d.unpack_ev) if (e) cudaEventDestroy(e);
    if (d.unpack) cudaStreamDestroy(d.unpack);
    if (d.run_start) cudaEventDestroy(d.run_start);
    if (d.run_stop) cudaEventDestroy(d.run_stop);
    if (d.d_invalid) cudaFree(d.d_invalid);
    if (d.clk_probe) cudaFree(d.clk_probe);
    if (d.h_invalid) cudaFreeHost(d.h_invalid);
    if (d.compute) cudaStreamDestroy(d.compute);
    if (d.compute2) cudaStreamDestroy(d.compute2);
    if (d.copy) cudaStreamDestroy(d.copy);
    if (d.copy_in) cudaStreamDestroy(d.copy_in);
    if (d.post) cudaStreamDestroy(d.post);
}

void host_lut(uint8_t lut[256], uint32_t valid[8]) {
    // encoding.rs:7-38
    std::memset(lut, 0, 256);
    const char* letters = "AGCTRMWSKYVHDBN";
    const uint8_t codes[15] = {136, 72, 40, 24, 192, 160, 144, 96, 80, 48, 224, 176, 208, 112, 240};
    for (int i = 0; i < 15; i++) {
        lut[(uint8_t)letters[i]] = codes[i];
        lut[(uint8_t)(letters[i] + 32)] = codes[i];
    }
    lut[(uint8_t)'-'] = 244;
    lut[(uint8_t)'?'] = 242;
    std::memset(valid, 0, 32);
    for (int i = 0; i < 256; i++)
        if (lut[i]) valid[lut[i] >> 5] |= 1u << (lut[i] & 31);
}

}  // namespace

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

int dg_abi_version(void) { return DG_ABI_VERSION; }

int dg_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return DG_ERR_CUDA;
    }
    return n;
}

const char* dg_last_error(const dg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dg_create(const int* gpu_ids, int n_gpus, int measure, uint64_t width, dg_ctx** out) {
    if (!out) { g_create_error = "out is NULL"; return DG_ERR_INVALID_ARG; }
    *out = nullptr;
    if (measure < 0 || measure > 5) { g_create_error = "unknown measure"; return DG_ERR_INVALID_ARG; }
    if (width == 0 || width >= (1ull << 31)) { g_create_error = "width must be in [1, 2^31)"; return DG_ERR_INVALID_ARG; }
    if (n_gpus < 1) { g_create_error = "n_gpus must be >= 1"; return DG_ERR_INVALID_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device available (this library has no CPU fallback)";
        return DG_ERR_CUDA;
    }
    dg_ctx* c = new dg_ctx();
    c->measure = measure;
    c->fam = family_of(measure);
    c->width = width;
    c->wp = (uint32_t)(((width + 31) / 32 + KC - 1) / KC * KC);
    int rc = DG_OK;
    try {
        uint8_t lut[256];
        uint32_t valid[8];
        host_lut(lut, valid);
        if (n_gpus > 1) {   // primary contexts take ~1 s each to create: bring them up side by side
            std::vector<std::thread> th;
            for (int i = 0; i < n_gpus; i++) {
                const int id = gpu_ids ? gpu_ids[i] : i;
                if (id >= 0 && id < ndev) th.emplace_back([id] { if (cudaSetDevice(id) == cudaSuccess) cudaFree(nullptr); });
            }
            for (auto& t : th) t.join();
            cudaGetLastError();
        }
        for (int i = 0; i < n_gpus; i++) {
            Device d;
            d.id = gpu_ids ? gpu_ids[i] : i;
            if (d.id < 0 || d.id >= ndev) fail(DG_ERR_INVALID_ARG, "GPU id %d out of range (have %d)", d.id, ndev);
            CUDA_CHECK(cudaSetDevice(d.id));
            cudaDeviceProp prop{};
            CUDA_CHECK(cudaGetDeviceProperties(&prop, d.id));
            if (prop.major != 10)
                fail(DG_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", d.id, prop.major, prop.minor);
            CUDA_CHECK(cudaStreamCreateWithFlags(&d.compute, cudaStreamNonBlocking));
            CUDA_CHECK(cudaStreamCreateWithFlags(&d.compute2, cudaStreamNonBlocking));
            CUDA_CHECK(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
            CUDA_CHECK(cudaStreamCreateWithFlags(&d.copy_in, cudaStreamNonBlocking));
            CUDA_CHECK(cudaStreamCreateWithFlags(&d.post, cudaStreamNonBlocking));
            {
                int lo_pri = 0, hi_pri = 0;
                CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
                CUDA_CHECK(cudaStreamCreateWithPriority(&d.prep, cudaStreamNonBlocking, hi_pri));
                CUDA_CHECK(cudaStreamCreateWithPriority(&d.fill, cudaStreamNonBlocking, hi_pri));
                CUDA_CHECK(cudaStreamCreateWithPriority(&d.unpack, cudaStreamNonBlocking, lo_pri));
                CUDA_CHECK(cudaEventCreateWithFlags(&d.sq_ready, cudaEventDisableTiming));
            }
            for (auto& s : d.slot)
                for (cudaEvent_t* e : {&s.k_start, &s.k_stop, &s.copied, &s.in_ready, &s.p_start, &s.p_stop, &s.gemm_done})
                    CUDA_CHECK(cudaEventCreate(e));
            for (auto& s : d.pslot)
                for (cudaEvent_t* e : {&s.k_start, &s.k_stop, &s.copied, &s.gemm_done})
                    CUDA_CHECK(cudaEventCreate(e));
            CUDA_CHECK(cudaEventCreate(&d.run_start));
            CUDA_CHECK(cudaEventCreate(&d.run_stop));
            CUDA_CHECK(cudaMalloc(&d.pp_cnt, (size_t)width * 4));
            CUDA_CHECK(cudaMalloc(&d.pp_cursor, (size_t)width * 4));
            CUDA_CHECK(cudaMalloc(&d.pp_work, 8));
            CUDA_CHECK(cudaHostAlloc(&d.h_pp_total, 4, cudaHostAllocDefault));
            CUDA_CHECK(cudaHostAlloc(&d.h_pp_work, 8, cudaHostAllocDefault));
            for (auto& e : d.chunk_ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            for (auto& e : d.unpack_ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            set_side_kernel_carveouts();
            CUDA_CHECK(cudaMemcpyToSymbol(c_ascii_lut, lut, 256));
            CUDA_CHECK(cudaMemcpyToSymbol(c_valid_code, valid, 32));
            CUDA_CHECK(cudaMalloc(&d.d_invalid, 3 * sizeof(unsigned long long)));
            CUDA_CHECK(cudaHostAlloc(&d.h_invalid, 3 * sizeof(unsigned long long), cudaHostAllocDefault));
            for (int k = 0; k < 3; k++) d.h_invalid[k] = ~0ull;
            CUDA_CHECK(cudaMemset(d.d_invalid, 0xff, 3 * sizeof(unsigned long long)));
            CUDA_CHECK(cudaMalloc(&d.clk_probe, 8 * sizeof(unsigned long long)));
            CUDA_CHECK(cudaMemset(d.clk_probe, 0, 8 * sizeof(unsigned long long)));
            c->devs.push_back(d);
        }
    } catch (const DgError& e) {
        g_create_error = e.msg;
        rc = e.code;
    }
    if (rc != DG_OK) {
        for (auto& d : c->devs) destroy_device(d);
        delete c;
        return rc;
    }
    if (const char* e = std::getenv("DG_ENGINE")) {  // developer override: 1 = LOP3+POPC tiles, 2 = tcgen05 int8 GEMM, 3 = fp4
        const int v = std::atoi(e);
        if (v >= 0 && v <= 3 && !(v == 3 && !c->fp4_exact()) && !(v == 2 && !c->i8_ok())) c->engine = v;
    }
    *out = c;
    return DG_OK;
}

void dg_destroy(dg_ctx* ctx) {
    if (!ctx) return;
    for (auto& d : ctx->devs) destroy_device(d);
    delete ctx;
}

int dg_set_option(dg_ctx* ctx, int key, int64_t value) {
    return guarded(ctx, [&] {
        switch (key) {
        case DG_OPT_PANEL_BYTES:
            if (value < 4096) fail(DG_ERR_INVALID_ARG, "panel bytes too small");
            ctx->panel_bytes = (size_t)value;
            break;
        case DG_OPT_KEEP_CODES: ctx->keep_codes = value != 0; break;
        case DG_OPT_TILE_VARIANT: ctx->tile_variant = (int)value; break;
        case DG_OPT_RESULT_U16:
            if (value != 0 && ctx->width > 65535) fail(DG_ERR_INVALID_ARG, "uint16 results need width <= 65535");
            ctx->result_u16 = value != 0;
            break;
        case DG_OPT_RESULT_U8:
            if (value != 0 && ctx->width > 65535) fail(DG_ERR_INVALID_ARG, "uint8 results need width <= 65535");
            ctx->result_u8 = value != 0;
            break;
        case DG_OPT_RESULT_COUNTS: ctx->result_counts = value != 0; break;
        case DG_OPT_PIPE_PANELS:
            if (value < 1 || value > 4096) fail(DG_ERR_INVALID_ARG, "pipe panels must be in [1, 4096]");
            ctx->pipe_panels = (int)value;
            break;
        case DG_OPT_REPACK_OVERLAP: ctx->repack_overlap = value != 0; break;
        case DG_OPT_PIPE_CHUNK_BYTES:
            if (value < 0) fail(DG_ERR_INVALID_ARG, "chunk bytes must be >= 0");
            ctx->pipe_chunk_bytes = (uint64_t)value;
            break;
        case DG_OPT_ENGINE:
            if (value < 0 || value > 3) fail(DG_ERR_INVALID_ARG, "unknown engine %lld", (long long)value);
            if (value == 3 && !ctx->fp4_exact())
                fail(DG_ERR_INVALID_ARG, "engine 3 (tcgen05 kind::mxf4, fp32 accumulation) is exact only for widths up to %llu sites; "
                     "use 0 (auto), 2 (int8, int32 accumulation) or 1 (LOP3+POPC)", (unsigned long long)((1ull << 22) - 256));
            if (value == 2 && !ctx->i8_ok())
                fail(DG_ERR_INVALID_ARG, "engine 2 (tcgen05 kind::i8) takes widths below 2^27 sites; use 0 (auto) or 1 (LOP3+POPC)");
            ctx->engine = (int)value;
            break;
        default: fail(DG_ERR_INVALID_ARG, "unknown option %d", key);
        }
    });
}

// codes: host memory (src_dev < 0) or device memory on CUDA device src_dev (then every device of the context gets
// its replica through cudaMemcpyPeerAsync: NVLink when the devices are peers, no PCIe upload at all)
static int load_resident_impl(dg_ctx* ctx, int which, const uint8_t* codes, int src_dev, uint64_t n, int input_kind,
                              const uint64_t* acgt_counts) {
    return guarded(ctx, [&] {
        if (which < 0 || which > 1) fail(DG_ERR_INVALID_ARG, "which must be 0 or 1");
        if (!codes || n == 0) fail(DG_ERR_INVALID_ARG, "empty alignment");
        if (n >= (1ull << 31)) fail(DG_ERR_INVALID_ARG, "too many records");
        if (!valid_input_kind(input_kind)) fail(DG_ERR_INVALID_ARG, "bad input_kind");
        if (ctx->streaming || ctx->sq_open) fail(DG_ERR_STATE, "a session is open");
        ctx->have_invalid = false;
        const bool nibble = input_kind == DG_INPUT_NIBBLE;
        const uint64_t stride = input_stride(ctx, input_kind);   // bytes per record of the caller's buffer
        const int dkind = device_kind(input_kind);
        std::vector<uint32_t> c32;
        if (acgt_counts) {
            c32.resize(n * 4);
            for (uint64_t i = 0; i < n * 4; i++) c32[i] = (uint32_t)acgt_counts[i];
        }
        const bool want_tc = ctx->tc_wanted();
        const bool needs_pp = tc_schedule(ctx->fam).needs_pp;
        const bool trace = std::getenv("DG_TRACE") != nullptr;
        const double t_begin = wall_ms();
        for (auto& d : ctx->devs) {
            CUDA_CHECK(cudaSetDevice(d.id));
            PlaneSet& s = d.set[which];
            reserve_resident(ctx, s, n, want_tc);
            if (nibble) ensure_nib(ctx, s, n);
            s.input_kind = dkind;
            s.acgt_from_host = acgt_counts != nullptr;
            // Upload in chunks on the copy stream; each chunk is packed on the compute stream as soon as
            // it has landed, so packing hides behind the PCIe transfer of the next chunk.
            const uint64_t target = std::max<uint64_t>(48ull << 20, n * ctx->width / 30);  // <= 32 chunks
            const uint64_t chunk = std::max<uint64_t>(ROW_ALIGN, (target / ctx->width + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN);
            uint32_t* pp_cnt = d.pp_cnt;
            int n_ev = 0;
            s.pp_stale = false;
            try {
                if (want_tc && needs_pp) {
                    ensure_pp_hits(ctx, d, n);
                    CUDA_CHECK(cudaMemsetAsync(pp_cnt, 0, (size_t)ctx->width * 4, d.compute));
                    CUDA_CHECK(cudaMemsetAsync(d.pp_hit_count, 0, 4, d.compute));
                }
                const double th = wall_ms();
                CUDA_CHECK(cudaEventRecord(d.slot[0].p_start, d.compute));
                for (uint64_t r0 = 0; r0 < n; r0 += chunk) {
                    const uint64_t nr = std::min(chunk, n - r0);
                    uint8_t* land = nibble ? s.nib + r0 * stride : s.codes + r0 * stride;   // where the chunk lands
                    if (src_dev < 0)
                        CUDA_CHECK(cudaMemcpyAsync(land, codes + r0 * stride, (size_t)nr * stride, cudaMemcpyHostToDevice, d.copy_in));
                    else
                        CUDA_CHECK(cudaMemcpyPeerAsync(land, d.id, codes + r0 * stride, src_dev, (size_t)nr * stride, d.copy_in));
                    cudaEvent_t ev = d.chunk_ev[n_ev++ & 31];
                    CUDA_CHECK(cudaEventRecord(ev, d.copy_in));
                    CUDA_CHECK(cudaStreamWaitEvent(d.compute, ev, 0));
                    if (nibble) enqueue_nibble_unpack(ctx, land, s.codes + r0 * ctx->width, nr, d.compute);
                    if (want_tc) {
                        enqueue_tc_pack(ctx, s, s.codes, nr, dkind, !acgt_counts && ctx->fam == FAM_TN93, d.compute, r0,
                                        d.d_invalid + 2, false, needs_pp ? &d : nullptr);   // also collects the partial codes
                    }
                }
                if (acgt_counts) {
                    CUDA_CHECK(cudaMemsetAsync(s.acgt, 0, (size_t)s.n_pad * 16, d.compute));
                    CUDA_CHECK(cudaMemcpyAsync(s.acgt, c32.data(), (size_t)n * 16, cudaMemcpyHostToDevice, d.compute));
                }
                if (want_tc && needs_pp) finish_pp_index(ctx, d, s, d.compute, true);   // scan + scatter (one sync for the entry count)
                CUDA_CHECK(cudaEventRecord(d.slot[0].p_stop, d.compute));
                CUDA_CHECK(cudaStreamSynchronize(d.copy_in));
                CUDA_CHECK(cudaStreamSynchronize(d.compute));
                ctx->tm.h2d_ms += wall_ms() - th;
                if (src_dev < 0) ctx->tm.h2d_bytes += n * stride;
                float ms = 0;
                CUDA_CHECK(cudaEventElapsedTime(&ms, d.slot[0].p_start, d.slot[0].p_stop));
                ctx->tm.pack_ms += ms;  // device span of the pipelined upload + packing
                s.tc_ready = want_tc;
                if (!want_tc) ensure_lop3(ctx, d, s);  // engine 1: build the bit-planes now (also validates the bytes)
                check_invalid(ctx, d, s.codes, 0);
            } catch (...) {
                s.n = 0; s.tc_ready = false; s.lop3_ready = false;
                throw;
            }
        }
        if (trace) fprintf(stderr, "[dg_load_resident] %llu records: %.3f ms\n", (unsigned long long)n, wall_ms() - t_begin);
    });
}

int dg_load_resident(dg_ctx* ctx, int which, const uint8_t* codes, uint64_t n, int input_kind,
                     const uint64_t* acgt_counts) {
    return load_resident_impl(ctx, which, codes, -1, n, input_kind, acgt_counts);
}

int dg_load_resident_device(dg_ctx* ctx, int which, const uint8_t* d_codes, int src_device, uint64_t n, int input_kind,
                            const uint64_t* acgt_counts) {
    if (ctx && src_device < 0) { ctx->err = "src_device must be a CUDA device ordinal"; return DG_ERR_INVALID_ARG; }
    return load_resident_impl(ctx, which, d_codes, src_device, n, input_kind, acgt_counts);
}

int dg_invalid_site(const dg_ctx* ctx, uint64_t* record, uint64_t* site, uint8_t* byte) {
    if (!ctx || !ctx->have_invalid) return DG_ERR_STATE;
    if (record) *record = ctx->inv_record;
    if (site) *site = ctx->inv_site;
    if (byte) *byte = ctx->inv_byte;
    return DG_OK;
}

int dg_run_square(dg_ctx* ctx, dg_sink_fn sink, void* user, uint32_t flags) {
    return guarded(ctx, [&] { run_mode(ctx, DG_MODE_SQUARE, 0, 1, sink, user, flags); });
}

int dg_run_rect(dg_ctx* ctx, dg_sink_fn sink, void* user, uint32_t flags) {
    return guarded(ctx, [&] { run_mode(ctx, DG_MODE_RECT, 0, 1, sink, user, flags); });
}

int dg_run_part(dg_ctx* ctx, int mode, uint32_t part, uint32_t n_parts, dg_sink_fn sink, void* user,
                uint32_t flags) {
    return guarded(ctx, [&] {
        if (mode != DG_MODE_SQUARE && mode != DG_MODE_RECT) fail(DG_ERR_INVALID_ARG, "mode must be SQUARE or RECT");
        run_mode(ctx, mode, part, n_parts, sink, user, flags);
    });
}

int64_t dg_plan_panels(int measure, int mode, uint64_t n_rows, uint64_t n_cols, uint64_t panel_bytes,
                       int tile_variant, uint64_t* row_begin, uint64_t* row_end, uint64_t* n_results,
                       uint64_t cap) {
    if (measure < 0 || measure > 5 || (mode != DG_MODE_SQUARE && mode != DG_MODE_RECT) || panel_bytes < 4096)
        return DG_ERR_INVALID_ARG;
    const TileShape ts = tile_shape(family_of(measure), tile_variant);
    static const uint64_t fam_items[4] = {1, 2, 3, 5};   // accumulators of the tensor schedules (default engine: fp4, 240-column tiles)
    const std::vector<Panel> v = make_panels(panel_bytes, measure <= 1 ? 4 : 8, mode, n_rows,
                                             mode == DG_MODE_SQUARE ? n_rows : n_cols, ts.tm, tc::TN_FP4, fam_items[family_of(measure)]);
    for (size_t k = 0; k < v.size() && k < cap; k++) {
        if (row_begin) row_begin[k] = v[k].row0;
        if (row_end) row_end[k] = v[k].row1;
        if (n_results) n_results[k] = v[k].n_results;
    }
    return (int64_t)v.size();
}

int dg_plan_parts(const uint64_t* n_results, uint64_t count, uint32_t n_parts, uint32_t* part_of) {
    if (!n_results || !part_of || n_parts == 0) return DG_ERR_INVALID_ARG;
    const std::vector<uint32_t> v = assign_parts(n_results, (size_t)count, n_parts);
    for (size_t k = 0; k < v.size(); k++) part_of[k] = v[k];
    return DG_OK;
}

int64_t dg_plan_ctx(dg_ctx* ctx, int mode, uint64_t* row_begin, uint64_t* row_end, uint64_t* n_results, uint64_t cap) {
    if (!ctx) return DG_ERR_INVALID_ARG;
    int64_t count = 0;
    const int rc = guarded(ctx, [&] {
        if (mode != DG_MODE_SQUARE && mode != DG_MODE_RECT) fail(DG_ERR_INVALID_ARG, "mode must be SQUARE or RECT");
        const PlaneSet& A = ctx->devs[0].set[0];
        const PlaneSet& B = ctx->devs[0].set[mode == DG_MODE_SQUARE ? 0 : 1];
        if (A.n == 0 || B.n == 0) fail(DG_ERR_STATE, "alignment not loaded");
        const TileShape ts = tile_shape(ctx->fam, ctx->tile_variant);
        const std::vector<Panel> v = make_panels(ctx->panel_bytes, ctx->elem_bytes(), mode, A.n, B.n, ts.tm, ctx->plan_tn(), ctx->plan_items());
        for (size_t k = 0; k < v.size() && k < cap; k++) {
            if (row_begin) row_begin[k] = v[k].row0;
            if (row_end) row_end[k] = v[k].row1;
            if (n_results) n_results[k] = v[k].n_results;
        }
        count = (int64_t)v.size();
    });
    return rc != DG_OK ? rc : count;
}

int dg_square_begin(dg_ctx* ctx, uint64_t n, int input_kind, const uint64_t* acgt_counts, uint32_t part, uint32_t n_parts,
                    dg_sink_fn sink, void* user) {
    const int rc = guarded(ctx, [&] { sq_begin(ctx, DG_MODE_SQUARE, n, input_kind, acgt_counts, part, n_parts, sink, user); });
    if (rc != DG_OK && ctx && !ctx->streaming && !ctx->sq_open && rc != DG_ERR_STATE) sq_abort(ctx);
    return rc;
}

int dg_rect_begin(dg_ctx* ctx, uint64_t n, int input_kind, const uint64_t* acgt_counts, uint32_t part, uint32_t n_parts,
                  dg_sink_fn sink, void* user) {
    const int rc = guarded(ctx, [&] { sq_begin(ctx, DG_MODE_RECT, n, input_kind, acgt_counts, part, n_parts, sink, user); });
    if (rc != DG_OK && ctx && !ctx->streaming && !ctx->sq_open && rc != DG_ERR_STATE) sq_abort(ctx);
    return rc;
}

int dg_square_next(dg_ctx* ctx, uint64_t* lo, uint64_t* hi) {
    return guarded(ctx, [&] {
        if (!ctx->sq_open) fail(DG_ERR_STATE, "no dg_square session is open");
        if (!lo || !hi) fail(DG_ERR_INVALID_ARG, "lo / hi is NULL");
        if (ctx->sq_pushed < ctx->sq_chunks.size()) { *lo = ctx->sq_chunks[ctx->sq_pushed].lo; *hi = ctx->sq_chunks[ctx->sq_pushed].hi; }
        else { *lo = 0; *hi = 0; }
    });
}

int64_t dg_square_plan(dg_ctx* ctx, uint64_t* lo, uint64_t* hi, uint64_t cap) {
    if (!ctx) return DG_ERR_INVALID_ARG;
    int64_t count = 0;
    const int rc = guarded(ctx, [&] {
        if (!ctx->sq_open) fail(DG_ERR_STATE, "no dg_square session is open");
        for (size_t g = 0; g < ctx->sq_chunks.size() && g < cap; g++) {
            if (lo) lo[g] = ctx->sq_chunks[g].lo;
            if (hi) hi[g] = ctx->sq_chunks[g].hi;
        }
        count = (int64_t)ctx->sq_chunks.size();
    });
    return rc != DG_OK ? rc : count;
}

int dg_square_push(dg_ctx* ctx, const uint8_t* codes, int src_device, uint64_t lo, uint64_t hi, void* ready_event) {
    const int rc = guarded(ctx, [&] { sq_push(ctx, codes, src_device, lo, hi, ready_event); });
    if (rc != DG_OK && ctx && ctx->sq_open) sq_abort(ctx);
    return rc;
}

int dg_square_end(dg_ctx* ctx) {
    const int rc = guarded(ctx, [&] { sq_end(ctx); });
    if (rc != DG_OK && ctx && ctx->sq_open) sq_abort(ctx);
    return rc;
}

static int run_session_host(dg_ctx* ctx, int mode, const uint8_t* codes, uint64_t n, int input_kind, const uint64_t* acgt_counts,
                            uint32_t part, uint32_t n_parts, dg_sink_fn sink, void* user) {
    if (ctx && !codes) { ctx->err = "codes is NULL"; return DG_ERR_INVALID_ARG; }
    int rc = mode == DG_MODE_RECT ? dg_rect_begin(ctx, n, input_kind, acgt_counts, part, n_parts, sink, user)
                                  : dg_square_begin(ctx, n, input_kind, acgt_counts, part, n_parts, sink, user);
    if (rc != DG_OK) return rc;
    for (;;) {
        uint64_t lo = 0, hi = 0;
        if ((rc = dg_square_next(ctx, &lo, &hi)) != DG_OK) return rc;
        if (hi == lo) break;
        if ((rc = dg_square_push(ctx, codes + lo * input_stride(ctx, input_kind), -1, lo, hi, nullptr)) != DG_OK) return rc;
    }
    return dg_square_end(ctx);
}

int dg_run_square_host(dg_ctx* ctx, const uint8_t* codes, uint64_t n, int input_kind, const uint64_t* acgt_counts,
                       uint32_t part, uint32_t n_parts, dg_sink_fn sink, void* user) {
    return run_session_host(ctx, DG_MODE_SQUARE, codes, n, input_kind, acgt_counts, part, n_parts, sink, user);
}

int dg_run_rect_host(dg_ctx* ctx, const uint8_t* codes, uint64_t n, int input_kind, const uint64_t* acgt_counts,
                     uint32_t part, uint32_t n_parts, dg_sink_fn sink, void* user) {
    return run_session_host(ctx, DG_MODE_RECT, codes, n, input_kind, acgt_counts, part, n_parts, sink, user);
}

int dg_stream_begin(dg_ctx* ctx, dg_sink_fn sink, void* user, uint64_t max_batch) {
    return guarded(ctx, [&] { stream_begin(ctx, sink, user, max_batch); });
}

int dg_stream_push(dg_ctx* ctx, const uint8_t* codes, uint64_t n_batch, int input_kind,
                   const uint64_t* acgt_counts) {
    int rc = guarded(ctx, [&] {
        if (!ctx->streaming) fail(DG_ERR_STATE, "no stream session is open");
        if (!codes && n_batch) fail(DG_ERR_INVALID_ARG, "codes is NULL");
        if (!valid_input_kind(input_kind)) fail(DG_ERR_INVALID_ARG, "bad input_kind");
        for (uint64_t off = 0; off < n_batch; off += ctx->s_max_batch) {
            const uint64_t nb = std::min(ctx->s_max_batch, n_batch - off);
            stream_push_one(ctx, codes + off * input_stride(ctx, input_kind), nb, input_kind,
                            acgt_counts ? acgt_counts + off * 4 : nullptr);
        }
    });
    if (rc != DG_OK && ctx && ctx->streaming) stream_abort(ctx);
    return rc;
}

int dg_stream_buffer(dg_ctx* ctx, uint8_t** buf, uint64_t* capacity_records) {
    int rc = guarded(ctx, [&] {
        if (!ctx->streaming) fail(DG_ERR_STATE, "no stream session is open");
        if (!buf) fail(DG_ERR_INVALID_ARG, "buf is NULL");
        const int ndev = (int)ctx->devs.size();
        // the staging slot of the NEXT batch; it is free once the batch that used it before has been sunk
        while ((int)ctx->s_queue.size() >= 2 * ndev) stream_sink_front(ctx);
        const uint64_t bi = ctx->s_batches;
        *buf = ctx->devs[bi % ndev].slot[(bi / ndev) & 1].h_in;
        if (capacity_records) *capacity_records = ctx->s_max_batch;
    });
    if (rc != DG_OK && ctx && ctx->streaming) stream_abort(ctx);
    return rc;
}

int dg_stream_end(dg_ctx* ctx) {
    int rc = guarded(ctx, [&] {
        if (!ctx->streaming) fail(DG_ERR_STATE, "no stream session is open");
        while (!ctx->s_queue.empty()) stream_sink_front(ctx);
        harvest_clock(ctx);
        ctx->streaming = false;
        ctx->tm.total_ms = wall_ms() - ctx->s_t0;
    });
    if (rc != DG_OK && ctx && ctx->streaming) stream_abort(ctx);
    return rc;
}

int dg_debug_counts(dg_ctx* ctx, int which_a, int which_b, uint32_t* out) {
    return guarded(ctx, [&] {
        if (which_a < 0 || which_a > 1 || which_b < 0 || which_b > 1 || !out) fail(DG_ERR_INVALID_ARG, "bad arguments");
        Device& d = ctx->devs[0];
        CUDA_CHECK(cudaSetDevice(d.id));
        if (ctx->streaming || ctx->sq_open) fail(DG_ERR_STATE, "a session is open");
        ensure_pp_index(ctx, d, d.set[which_a]);
        ensure_pp_index(ctx, d, d.set[which_b]);
        const PlaneSet& A = d.set[which_a];
        const PlaneSet& B = d.set[which_b];
        if (A.n == 0 || B.n == 0) fail(DG_ERR_STATE, "alignment not loaded");
        const size_t bytes = (size_t)A.n * B.n * 16;
        uint4* dbuf = nullptr;
        CUDA_CHECK(cudaMalloc(&dbuf, bytes));
        try {
            Panel p;
            p.row0 = 0; p.row1 = A.n; p.out_base = 0; p.n_results = A.n * B.n;
            // tall alignments: split into launches of <= 65535 row blocks
            const TileShape ts = tile_shape(ctx->fam, ctx->tile_variant);
            const uint64_t step = (uint64_t)ts.tm * 32768;
            if (use_tc(ctx, A, B)) {
                // tensor engine: raw sums of every accumulator -> the same canonical counts
                int* scratch = nullptr;
                const uint64_t rows_step = std::min<uint64_t>(A.n, 32768);
                CUDA_CHECK(cudaMalloc(&scratch, scratch_upper_bound(ctx, A.tc_fp4, DG_MODE_RECT, 0, rows_step, B.n)));
                try {
                    for (uint64_t r = 0; r < A.n; r += rows_step) {
                        Panel q = p;
                        q.row0 = r; q.row1 = std::min<uint64_t>(A.n, r + rows_step);
                        q.out_base = 0; q.n_results = (q.row1 - q.row0) * B.n;
                        enqueue_panel_tc(ctx, d, A, B, DG_MODE_RECT, q, dbuf + r * B.n, scratch, false, true, d.compute);
                        CUDA_CHECK(cudaStreamSynchronize(d.compute));
                    }
                } catch (...) { cudaFree(scratch); throw; }
                cudaFree(scratch);
            } else {
            ensure_lop3(ctx, d, d.set[which_a]);
            ensure_lop3(ctx, d, d.set[which_b]);
            for (uint64_t r = 0; r < A.n; r += step) {
                Panel q = p;
                q.row0 = r; q.row1 = std::min<uint64_t>(A.n, r + step);
                enqueue_panel_kernel<true>(ctx, d, A, B, DG_MODE_RECT, q, dbuf + r * B.n, false, d.compute);
            }
            }
            CUDA_CHECK(cudaStreamSynchronize(d.compute));
            CUDA_CHECK(cudaMemcpy(out, dbuf, bytes, cudaMemcpyDeviceToHost));
        } catch (...) {
            cudaFree(dbuf);
            throw;
        }
        cudaFree(dbuf);
    });
}

int dg_debug_planes(dg_ctx* ctx, int which, uint32_t* core, uint32_t* aux, uint64_t* acgt, uint64_t* words_out) {
    return guarded(ctx, [&] {
        if (which < 0 || which > 1) fail(DG_ERR_INVALID_ARG, "which must be 0 or 1");
        Device& d = ctx->devs[0];
        CUDA_CHECK(cudaSetDevice(d.id));
        if (d.set[which].n == 0) fail(DG_ERR_STATE, "alignment not loaded");
        ensure_lop3(ctx, d, d.set[which]);
        const PlaneSet& s = d.set[which];
        if (words_out) *words_out = ctx->wp;
        const size_t bytes = (size_t)s.n * ctx->wp * 16;
        if (core) CUDA_CHECK(cudaMemcpy(core, s.core, bytes, cudaMemcpyDeviceToHost));
        if (aux) {
            if (!s.aux) fail(DG_ERR_STATE, "this measure family keeps no aux planes");
            CUDA_CHECK(cudaMemcpy(aux, s.aux, bytes, cudaMemcpyDeviceToHost));
        }
        if (acgt) {
            std::vector<uint32_t> tmp(s.n * 4);
            CUDA_CHECK(cudaMemcpy(tmp.data(), s.acgt, tmp.size() * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < tmp.size(); i++) acgt[i] = tmp[i];
        }
    });
}

int dg_get_timings(const dg_ctx* ctx, dg_timings* out) {
    if (!ctx || !out) return DG_ERR_INVALID_ARG;
    *out = ctx->tm;
    out->engine = (uint64_t)ctx->last_engine;
    out->sm_mhz = ctx->clk_ns > 0 ? ctx->clk_cycles / ctx->clk_ns * 1e3 : 0.0;
    return DG_OK;
}

int dg_reset_timings(dg_ctx* ctx) {
    if (!ctx) return DG_ERR_INVALID_ARG;
    ctx->tm = dg_timings{};
    ctx->clk_cycles = ctx->clk_ns = 0;
    return DG_OK;
}

void* dg_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void dg_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
