// tc_engine.cuh -- tcgen05 `kind::i8` one-hot GEMM variant of the count kernel (north_star item 2).
//
// The per-pair counts are a dense contraction over sites, so they can run on the 5th-gen tensor
// cores with exact int32 accumulation in TMEM.  Operand planes are int8, K-major:
//
//   ops[record][plane][site]      (plane ids and values: PlaneId below)
//
// A "schedule" lists, per integer count, the (A plane, B plane) pairs whose products are summed.  Every count
// is written as a bilinear form of MINIMAL rank over the four bases, so the MAC count per pair-site is
//
//   n, n_high : 3 DIFF = sum_b U_b(q) V_b(t)                                                     4 MAC
//               U_b = 1 - possibility bit, V_b = 3 * possibility bit - (|S| - 1); both 0 for N-like codes.
//               * q known (base a)    : sum_{b != a} V_b(t) = 3 [a not in St]   (exact for every t)
//               * t known (base a)    : 3 U_a(q) = 3 [a not in Sq]              (exact for every q)
//               * N-like on a side    : 0  (N, '-', '?' are never DIFF: measures.rs:17)
//               * both partial codes  : off by a small integer, repaired exactly by pp_correct_kernel from a
//                                       per-site inverted index of the (rare) partial codes
//               (J - I on four bases has rank 4: no exact formulation needs fewer planes.)
//   raw, jc69 : the same + SAME = sum_b K_b K_b                                                  8 MAC
//   k80       : CS = PURK.PURK + PYRK.PYRK = SAME + ts;  X = W.W + Z.Z = SAME - ts  (W = K_A - K_G,
//               Z = K_C - K_T);  tv = PURC.PYRC' + PYRC.PURC' (measures.rs:90-103)                6 MAC
//   tn93      : L = K.K;  PP = PURK.PURK = SP + P1;  YY = SY + P2;  W.W = SP - P1;  Z.Z = SY - P2
//               -> d = L - SP - SY (measures.rs:156-175)                                          5 MAC
// DIFF <=> (q & t) < 16 (measures.rs:14-23).  int32 accumulation is exact; the halvings are exact.
//
// n / n_high without pending corrections: the GEMM epilogue stores acc / 3 directly (uint32 or uint16).
// Otherwise each accumulator's raw int32 sums go to a scratch matrix, pp_correct repairs accumulator 0 and
// tc_combine_kernel derives the reference's counts and runs the same f64 epilogues as the LOP3 path.
//
// Kernel: persistent, warp-specialised CTA of 192 threads per SM.
//   warp 0     : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, 4-stage mbarrier ring)
//   warp 1     : TMEM allocator + MMA issuer (one elected lane, tcgen05.mma.cta_group::1.kind::i8,
//                M=128 N=256 K=32, accumulators 2 x 256 TMEM columns, tcgen05.commit -> mbarriers)
//   warps 2..5 : epilogue (tcgen05.ld 32x32b.x32 -> DIFF -> global store in reference order)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace dg {
namespace tc {

constexpr int TM = 128;          // tile rows    (UMMA M)
constexpr int TN = 256;          // tile columns (UMMA N)
constexpr int TN_FP4 = 240;      // tile columns of the FP4 variant
constexpr int KB = 128;          // K bytes per stage = one 128B swizzle atom = 128 sites of one plane
constexpr int STAGES = 4;        // barrier slots; MT = 1 uses 4 stages of 48 KB, MT = 2 uses 3 stages of 64 KB
constexpr int A_BYTES = TM * KB;
constexpr int B_BYTES = TN * KB;
constexpr int EPI_PITCH = 33;     // words per row of an epilogue warp's 32 x 32 transpose buffer (+1: conflict-free)
constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;
constexpr int SMEM_BYTES = 192 * 1024 + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
template <int MT, bool PAIR = false> struct StageCfg {
    static constexpr int BYTES = MT * A_BYTES + (PAIR ? B_BYTES / 2 : B_BYTES);  // a pair CTA stages only its half of B
    static constexpr int N = (192 * 1024) / BYTES < STAGES ? (192 * 1024) / BYTES : STAGES;
};
constexpr int THREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 22;  // mbarrier polls before the kernel traps instead of hanging

// What the epilogue stores for one accumulator (per accumulator: TcParams::acc_op / CombineParams::acc_op):
//   ACC_RAW_I32 / ACC_RAW_I16 : the raw sum into a scratch array (int16 when every sum fits: width <= 32767 and no
//                               repair pending on that accumulator -- halves the scratch traffic)
//   ACC_DIV3_U32 / _U16       : sum / 3 = DIFF (accumulator 0 of n / n_high / raw / jc69 with no both-partial repair
//                               pending): the final count of n / n_high, or a 16-bit scratch value for the combine pass
//   ACC_MOD16                 : the raw sum modulo 2^16 (accumulator 0 = 3 * DIFF with the repair pending, width <= 32767):
//                               pp_correct adds its deltas modulo 2^16 and the combine pass recovers DIFF, because of the
//                               two candidates x and x + 65536 (3 * DIFF < 2^17) exactly one is a multiple of 3
//   ACC_ADD_I32               : split-K launches (TcParams::ksplit > 1): every work item ADDS its partial sum into a zeroed
//                               int32 scratch array (small problems: the K range of a tile is dealt to several CTA pairs)
enum AccOp { ACC_RAW_I32 = 0, ACC_RAW_I16 = 1, ACC_DIV3_U32 = 2, ACC_DIV3_U16 = 3, ACC_MOD16 = 4, ACC_ADD_I32 = 5 };
__host__ __device__ constexpr bool acc_op_16(int op) { return op == ACC_RAW_I16 || op == ACC_DIV3_U16 || op == ACC_MOD16; }
__host__ __device__ constexpr bool acc_op_div3(int op) { return op == ACC_DIV3_U32 || op == ACC_DIV3_U16; }

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "h"(mask) : "memory");
}
// 2-SM (cta_group::2) TMA load: the bytes land in MY shared memory, the transaction count goes to the
// LEADER CTA's barrier at the same offset (peer bit 24 of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y) : "memory");
}
// the same with an L2 cache policy (createpolicy): operand tiles are re-read by the other tiles of the raster band
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// arrive on the barrier at this offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// one instruction for the CTA pair: D (256 x N, 128 rows in each CTA's TMEM) (+)= A (128 rows from each CTA) * B^T (N/2 rows from each CTA)
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// same, but the arrive lands on the barrier at this offset in every CTA of `mask` (cluster multicast)
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile whose rows are 128 B: 8-row groups are 1024 B apart (SBO),
// LBO = 1 (unused for swizzled K-major), descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// kind::i8 instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 256, M = 128.
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC_I8_PAIR = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);  // M = 256 over two CTAs

// kind::mxf4 block-scaled instruction descriptor (cute/arch/mma_sm100_desc.hpp, InstrDescriptorBlockScaled): a_format
// [7,10) = b_format [10,13) = 1 (E2M1), both K-major, n_dim [17,23) = N >> 3, scale_format [23] = 1 (UE8M0), m_dim
// [24,29) = M >> 4, scale-factor ids 0, k_size [31] = 0 (K = 64).  M = 256 over the CTA pair, N = 240.
constexpr uint32_t IDESC_F4_PAIR = (1u << 7) | (1u << 10) | ((uint32_t)(TN_FP4 >> 3) << 17) | (1u << 23) | ((uint32_t)((2 * TM) >> 4) << 24);
constexpr uint32_t SF_COL = TN_FP4;   // TMEM columns [240, 256) of the first accumulator's block hold the unit scale factors
__device__ __forceinline__ void tc_mma_f4_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate,
                                               uint32_t sfa_tmem, uint32_t sfb_tmem) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem) : "memory");
}

// ---- operand packing ------------------------------------------------------------------------------
// Plane ids.  A plane holds one int8 per site; every plane is 0 for N-like codes (N, '-', '?') and for the
// zero padding beyond `width` / beyond the last record, so padding lands in no count.
//   U_b   = 1 - possibility bit of base b                (A-side operand of DIFF)
//   V_b   = 3 * possibility bit - (|S| - 1)              (B-side operand of DIFF; values 3,2,1,0,-1,-2)
//   K_b   = known one-hot;  PURK / PYRK = known purine / pyrimidine;  K = known
//   W     = K_A - K_G,  Z = K_C - K_T                    (rank-1 factors of SAME - ts inside a class)
//   PURC / PYRC = code in {A,G,R} / {C,T,Y}              (measures.rs:90, 94)
enum PlaneId { P_UA = 0, P_UG, P_UC, P_UT, P_VA, P_VG, P_VC, P_VT, P_KA, P_KG, P_KC, P_KT,
               P_PURK, P_PYRK, P_K, P_W, P_Z, P_PURC, P_PYRC };
constexpr int MAX_PLANES = 12;
// The stored planes of every measure family, in storage order (the host-side schedules in dg_api.cu index into these
// lists).  Compile-time, so pack_ops_kernel unrolls the plane loop with every plane expression resolved.
template <int FAM> struct PackPlanes;
template <> struct PackPlanes<FAM_SNP> {
    static constexpr int N = 8;
    __host__ __device__ static constexpr uint32_t id(int i) { return i < 4 ? P_UA + i : P_VA + (i - 4); }
};
template <> struct PackPlanes<FAM_RAW> {
    static constexpr int N = 12;
    __host__ __device__ static constexpr uint32_t id(int i) { return i < 4 ? P_UA + i : (i < 8 ? P_VA + (i - 4) : P_KA + (i - 8)); }
};
template <> struct PackPlanes<FAM_K80> {
    static constexpr int N = 6;
    __host__ __device__ static constexpr uint32_t id(int i) {
        return i == 0 ? P_PURK : (i == 1 ? P_PYRK : (i == 2 ? P_W : (i == 3 ? P_Z : (i == 4 ? P_PURC : P_PYRC))));
    }
};
template <> struct PackPlanes<FAM_TN93> {
    static constexpr int N = 5;
    __host__ __device__ static constexpr uint32_t id(int i) { return i == 0 ? P_K : (i == 1 ? P_PURK : (i == 2 ? P_PYRK : (i == 3 ? P_W : P_Z))); }
};

struct PackI8Params {
    const uint8_t* codes;  // n x width
    uint64_t n, n_pad, width;
    uint64_t wp8;          // bytes per plane: width rounded up to 128 sites (int8) / 256 sites at 2 per byte (FP4)
    int8_t* ops;           // n_pad x nplanes x wp8
    uint32_t* acgt;        // n_pad x 4 (A,T,G,C) written per record, or NULL
    int ascii;
    int nibble;            // `codes` holds DG_INPUT_NIBBLE rows (two sites per byte, low nibble first) of in_stride bytes
    uint64_t in_stride;    // bytes between the input rows: width, or (width + 1) / 2 for nibble rows
    int count_upper_ascii; // ASCII input: count raw 'A','T','G','C' only (fastaio.rs:139-142) instead of count_bases
    int nplanes;
    uint8_t plane_id[MAX_PLANES];
    uint32_t plane_mask;   // bit pl set = store plane pl (records that only ever serve as COLUMNS need no A-side planes)
    unsigned long long* invalid;  // min over ((seq0 + record) << 32 | site) of invalid bytes; or NULL
    uint64_t seq0;                // global index of record 0 of this chunk
    // partial ambiguity codes (R Y M W S K V H D B) met while packing, for the both-partial repair index; or NULL:
    uint32_t* pp_site_cnt;        // [width] += 1 per partial code at that site
    uint64_t* pp_hits;            // unsorted entries (site << 36 | record << 4 | possibility nibble) ...
    uint32_t* pp_hit_count;       // ... appended at atomicAdd(pp_hit_count, 1) while that is < pp_hit_cap
    uint32_t pp_hit_cap;
};

constexpr uint32_t M1 = 0x01010101u;

// Four Paradis codes (one per byte) -> per-byte 0/1 words of the code bits and E = |S| - 1.
struct CodeBits { uint32_t A, G, C, T, K, E; };
__device__ __forceinline__ CodeBits code_bits(uint32_t w) {
    CodeBits b;
    b.A = (w >> 7) & M1; b.G = (w >> 6) & M1; b.C = (w >> 5) & M1; b.T = (w >> 4) & M1; b.K = (w >> 3) & M1;
    b.E = b.A + b.G + b.C + b.T - M1;   // every valid code has >= 1 possibility bit: no borrow between bytes
    return b;
}
// The four int8 values of plane `id` for the four codes of `b`.
__device__ __forceinline__ uint32_t plane_word(const CodeBits& b, uint32_t id) {
    switch (id) {
    case P_UA: return b.A ^ M1;
    case P_UG: return b.G ^ M1;
    case P_UC: return b.C ^ M1;
    case P_UT: return b.T ^ M1;
    case P_VA: return __vsub4(b.A * 3u, b.E);
    case P_VG: return __vsub4(b.G * 3u, b.E);
    case P_VC: return __vsub4(b.C * 3u, b.E);
    case P_VT: return __vsub4(b.T * 3u, b.E);
    case P_KA: return b.A & b.K;
    case P_KG: return b.G & b.K;
    case P_KC: return b.C & b.K;
    case P_KT: return b.T & b.K;
    case P_PURK: return (b.A | b.G) & b.K;
    case P_PYRK: return (b.C | b.T) & b.K;
    case P_K: return b.K;
    case P_W: return __vsub4(b.A & b.K, b.G & b.K);
    case P_Z: return __vsub4(b.C & b.K, b.T & b.K);
    case P_PURC: return (b.A | b.G) & ~(b.C | b.T);   // {A,G,R}: measures.rs:90
    default: return (b.C | b.T) & ~(b.A | b.G);       // P_PYRC {C,T,Y}: measures.rs:94
    }
}

// ---- FP4 planes: the same plane values as E2M1 nibbles, computed 8 sites at a time in nibble lanes -------------
// E2M1 codes of the values that occur: 0 -> 0x0, 1 -> 0x2, 2 -> 0x4, 3 -> 0x5, -1 -> 0xA, -2 -> 0xC.
constexpr uint32_t N1 = 0x11111111u;
// byte lanes (value < 16 in each of 4 bytes) -> 4 nibbles in the low 16 bits
__device__ __forceinline__ uint32_t squeeze4(uint32_t x) {
    x = (x | (x >> 4)) & 0x00FF00FFu;
    return (x | (x >> 8)) & 0xFFFFu;
}
// Eight Paradis codes (two words) -> 0/1 nibble-lane words of the code bits, and the two bits of E = |S| - 1.
struct NibBits { uint32_t A, G, C, T, K, e0, e1; };
__device__ __forceinline__ NibBits nib_bits(uint32_t wlo, uint32_t whi) {
    const uint32_t H = squeeze4((wlo >> 4) & 0x0F0F0F0Fu) | (squeeze4((whi >> 4) & 0x0F0F0F0Fu) << 16);  // nibble = A G C T
    const uint32_t L = squeeze4(wlo & 0x0F0F0F0Fu) | (squeeze4(whi & 0x0F0F0F0Fu) << 16);                // nibble = K . . .
    NibBits b;
    b.A = (H >> 3) & N1; b.G = (H >> 2) & N1; b.C = (H >> 1) & N1; b.T = H & N1; b.K = (L >> 3) & N1;
    const uint32_t E = b.A + b.G + b.C + b.T - N1;   // >= 1 possibility bit per valid code: no borrow between lanes
    b.e0 = E & N1; b.e1 = (E >> 1) & N1;
    return b;
}
// Each bit of the E2M1 code of V_b = 3 b - E is a 3-input boolean function of (b, e0, e1), i.e. ONE LOP3 (the inputs
// carry bits at the nibble LSB positions only and every function maps (0, 0, 0) to 0, so no masking is needed); the four
// result bits are merged with multiply-adds, which issue on the FMA pipe instead of the ALU pipe this kernel is bound by.
//   b = 1: E = 0 -> 3 (0101), 1 -> 2 (0100), 2 -> 1 (0010), 3 -> 0;   b = 0: E = 0 -> 0, 1 -> -1 (1010), 2 -> -2 (1100)
__device__ __forceinline__ uint32_t nib_v(const NibBits& n, uint32_t b) {   // V_b = 3 b - E as E2M1
    const uint32_t b0 = lop3<0x10>(b, n.e0, n.e1);   // b & ~e0 & ~e1
    const uint32_t b1 = lop3<0x24>(b, n.e0, n.e1);   // (b & e1 & ~e0) | (~b & e0 & ~e1)
    const uint32_t b2 = lop3<0x52>(b, n.e0, n.e1);   // (b & ~e1) | (~b & e1 & ~e0)
    const uint32_t b3 = lop3<0x06>(b, n.e0, n.e1);   // ~b & (e0 ^ e1)
    return b3 * 8u + b2 * 4u + b1 * 2u + b0;          // disjoint bits: the sums carry nothing
}
__device__ __forceinline__ uint32_t nib_pm(uint32_t pos, uint32_t neg) { return ((pos | neg) << 1) | (neg << 3); }  // +1 / -1
// The eight E2M1 nibbles of plane `id`.
__device__ __forceinline__ uint32_t plane_nib8(const NibBits& n, uint32_t id) {
    switch (id) {
    case P_UA: return (n.A ^ N1) << 1;
    case P_UG: return (n.G ^ N1) << 1;
    case P_UC: return (n.C ^ N1) << 1;
    case P_UT: return (n.T ^ N1) << 1;
    case P_VA: return nib_v(n, n.A);
    case P_VG: return nib_v(n, n.G);
    case P_VC: return nib_v(n, n.C);
    case P_VT: return nib_v(n, n.T);
    case P_KA: return (n.A & n.K) << 1;
    case P_KG: return (n.G & n.K) << 1;
    case P_KC: return (n.C & n.K) << 1;
    case P_KT: return (n.T & n.K) << 1;
    case P_PURK: return ((n.A | n.G) & n.K) << 1;
    case P_PYRK: return ((n.C | n.T) & n.K) << 1;
    case P_K: return n.K << 1;
    case P_W: return nib_pm(n.A & n.K, n.G & n.K);
    case P_Z: return nib_pm(n.C & n.K, n.T & n.K);
    case P_PURC: return ((n.A | n.G) & ~(n.C | n.T)) << 1;
    default: return ((n.C | n.T) & ~(n.A | n.G)) << 1;   // P_PYRC
    }
}

// Four possibility nibbles (one per byte, 0 .. 15) -> four Paradis codes: nibble << 4 | 8 if exactly one bit is set (the base
// is known); nibble 0 -> 0, which the validation below reports as an invalid byte.
__device__ __forceinline__ uint32_t nib4_to_codes(uint32_t m) {
    const uint32_t t = m & ((m | 0x10101010u) - M1);             // m & (m - 1) per byte: no borrow leaves a byte
    const uint32_t many = (t + 0x7F7F7F7Fu) & 0x80808080u;       // 0x80 where two or more bits are set
    const uint32_t some = (m + 0x7F7F7F7Fu) & 0x80808080u;       // 0x80 where the nibble is not 0
    return (m << 4) | ((some & ~many) >> 4);
}
// Four nibble bytes (8 sites, low nibble = the even site) -> two words of Paradis codes.
__device__ __forceinline__ void nib_word_to_codes(uint32_t x, uint32_t& c0, uint32_t& c1) {
    const uint32_t lo = x & 0x0F0F0F0Fu, hi = (x >> 4) & 0x0F0F0F0Fu;
    c0 = nib4_to_codes(__byte_perm(lo, hi, 0x5140));
    c1 = nib4_to_codes(__byte_perm(lo, hi, 0x7362));
}

// 16 code bytes of `row` starting at site s0 -> four words of valid Paradis codes (invalid bytes reported and
// replaced by N; sites >= width are N).  raw[] = the untranslated bytes (for the upper-case ASCII count quirk).
template <bool NIB>
__device__ __forceinline__ void load_codes16(const PackI8Params& p, const uint8_t* row, uint64_t seq, uint64_t s0,
                                             const uint8_t* lut, uint32_t (&w)[4], uint32_t (&raw)[4]) {
    if (NIB) {             // `row` holds nibble rows: sites [s0, s0 + 16) are its bytes [s0 / 2, s0 / 2 + 8)
        const uint8_t* src = row + (s0 >> 1);
        if (s0 + 16 <= p.width && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(src);
            nib_word_to_codes(__ldcs(sp), raw[0], raw[1]);
            nib_word_to_codes(__ldcs(sp + 1), raw[2], raw[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t m = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint64_t st = s0 + 4 * k + j;
                    const uint32_t nb = st < p.width ? ((uint32_t)row[st >> 1] >> (4 * (st & 1))) & 15u : 15u;
                    m |= nb << (8 * j);
                }
                raw[k] = nib4_to_codes(m);
            }
        }
    } else if (s0 + 20 <= p.width) {   // the aligned window [a & ~3, +20) stays inside this row
        const uintptr_t a = reinterpret_cast<uintptr_t>(row + s0);
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8;
        // streaming (evict-first) loads: the code bytes are read once and must not push the GEMM's operand tiles out of L2
        const uint32_t x0 = __ldcs(ap), x1 = __ldcs(ap + 1), x2 = __ldcs(ap + 2), x3 = __ldcs(ap + 3), x4 = __ldcs(ap + 4);
        raw[0] = __funnelshift_r(x0, x1, sh); raw[1] = __funnelshift_r(x1, x2, sh);
        raw[2] = __funnelshift_r(x2, x3, sh); raw[3] = __funnelshift_r(x3, x4, sh);
    } else {                    // last groups of the row: byte loads, bounded by width
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t s = s0 + 4 * k + j;
                const uint32_t byte = s < p.width ? (uint32_t)row[s] : (p.ascii ? (uint32_t)'N' : 240u);
                v |= byte << (8 * j);
            }
            raw[k] = v;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t r = raw[k];
        uint32_t t = (uint32_t)lut[r & 0xFF] | ((uint32_t)lut[(r >> 8) & 0xFF] << 8) |
                     ((uint32_t)lut[(r >> 16) & 0xFF] << 16) | ((uint32_t)lut[r >> 24] << 24);
        if ((t - M1) & ~t & 0x80808080u) {   // a zero byte = invalid nucleotide (fastaio.rs:111-113)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (((t >> (8 * j)) & 0xFFu) == 0) {
                    if (p.invalid) atomicMin(p.invalid, ((unsigned long long)(p.seq0 + seq) << 32) | (unsigned long long)(s0 + 4 * k + j));
                    t |= 0xF0u << (8 * j);
                }
        }
        w[k] = t;
    }
}

// One CTA per record (grid-stride), one thread per 16-byte store of every plane (16 sites of int8 planes, 32 sites of
// FP4 planes): the code bytes are fetched with aligned 32-bit loads + funnel shifts (rows are `width` bytes apart,
// so they are not 16-byte aligned), translated through a 256-entry LUT in shared memory (ASCII -> Paradis,
// encoding.rs:4-41; or Paradis -> itself if legal) and expanded to every stored plane with byte-SIMD arithmetic.
// Per-record A,T,G,C counts (count_bases, fastaio.rs:53-66) are reduced in the CTA: no atomics.
template <bool FP4, int FAM, bool NIB = false>
__global__ void __launch_bounds__(256, 3) pack_ops_kernel(PackI8Params p) {
    using PL = PackPlanes<FAM>;
    __shared__ uint8_t lut[256];
    __shared__ uint32_t red[8][4];
    {
        const uint32_t t = threadIdx.x;
        lut[t] = p.ascii ? c_ascii_lut[t] : (((c_valid_code[t >> 5] >> (t & 31)) & 1u) ? (uint8_t)t : (uint8_t)0);
    }
    __syncthreads();
    constexpr int H = FP4 ? 2 : 1;            // 16-site halves per thread
    constexpr uint32_t SITES = 16 * H;
    const uint32_t groups = (uint32_t)(p.wp8 / 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t seq = blockIdx.x; seq < p.n_pad; seq += gridDim.x) {
        uint32_t cA = 0, cT = 0, cG = 0, cC = 0;
        const uint8_t* row = p.codes + seq * p.in_stride;
        for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) {
            const uint64_t s0 = (uint64_t)g * SITES;
            uint32_t w[4 * H];
#pragma unroll
            for (int h = 0; h < H; h++) {
                uint32_t wh[4] = {0xF0F0F0F0u, 0xF0F0F0F0u, 0xF0F0F0F0u, 0xF0F0F0F0u};   // N-like padding
                if (seq < p.n && s0 + 16 * h < p.width) {
                    uint32_t raw[4];
                    load_codes16<NIB>(p, row, seq, s0 + 16 * h, lut, wh, raw);
                    if (p.pp_site_cnt) {
                        // partial code <=> "known" bit clear and possibility nibble != 1111 (0.1 % of real data: rare path)
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint32_t x = wh[k];
                            const uint32_t allf = x & (x >> 1) & (x >> 2) & (x >> 3) & 0x10101010u;
                            uint32_t m = ((~x & 0x08080808u) << 1) & ~allf;
                            while (m) {
                                const int j = (__ffs(m) - 1) >> 3;
                                m &= m - 1;
                                const uint64_t site = s0 + 16 * h + 4 * k + j;
                                atomicAdd(p.pp_site_cnt + site, 1u);
                                const uint32_t pos = atomicAdd(p.pp_hit_count, 1u);
                                if (pos < p.pp_hit_cap)
                                    p.pp_hits[pos] = (site << 36) | ((p.seq0 + seq) << 4) | (uint64_t)((x >> (8 * j + 4)) & 15u);
                            }
                        }
                    }
                    if (p.acgt && p.count_upper_ascii) {   // the streamed tn93 records count raw upper-case letters only
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            cA += __popc(__vcmpeq4(raw[k], 0x41414141u)); cT += __popc(__vcmpeq4(raw[k], 0x54545454u));
                            cG += __popc(__vcmpeq4(raw[k], 0x47474747u)); cC += __popc(__vcmpeq4(raw[k], 0x43434343u));
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) w[4 * h + k] = wh[k];
            }
            int8_t* base = p.ops + (seq * PL::N) * p.wp8 + (uint64_t)g * 16;
            if (FP4) {
                NibBits nb[4];
#pragma unroll
                for (int k = 0; k < 4; k++) nb[k] = nib_bits(w[2 * k], w[2 * k + 1]);
                if (p.acgt && !p.count_upper_ascii) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        cA += __popc(nb[k].A & nb[k].K); cT += __popc(nb[k].T & nb[k].K);
                        cG += __popc(nb[k].G & nb[k].K); cC += __popc(nb[k].C & nb[k].K);
                    }
                }
#pragma unroll
                for (int pl = 0; pl < PL::N; pl++) {
                    const uint32_t id = PL::id(pl);
                    if ((p.plane_mask >> pl) & 1u)
                        __stcs(reinterpret_cast<uint4*>(base + pl * p.wp8),
                               make_uint4(plane_nib8(nb[0], id), plane_nib8(nb[1], id), plane_nib8(nb[2], id), plane_nib8(nb[3], id)));
                }
            } else {
                CodeBits b[4];
#pragma unroll
                for (int k = 0; k < 4; k++) b[k] = code_bits(w[k]);
                if (p.acgt && !p.count_upper_ascii) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        cA += __popc(b[k].A & b[k].K); cT += __popc(b[k].T & b[k].K);
                        cG += __popc(b[k].G & b[k].K); cC += __popc(b[k].C & b[k].K);
                    }
                }
#pragma unroll
                for (int pl = 0; pl < PL::N; pl++) {
                    const uint32_t id = PL::id(pl);
                    if ((p.plane_mask >> pl) & 1u)
                        __stcs(reinterpret_cast<uint4*>(base + pl * p.wp8),
                               make_uint4(plane_word(b[0], id), plane_word(b[1], id), plane_word(b[2], id), plane_word(b[3], id)));
                }
            }
        }
        if (p.acgt) {   // uniform branch: CTA reduction of the four counts
            if (p.count_upper_ascii) { cA >>= 3; cT >>= 3; cG >>= 3; cC >>= 3; }   // vcmpeq4 sets 8 bits per hit
            for (int o = 16; o; o >>= 1) {
                cA += __shfl_xor_sync(0xffffffffu, cA, o); cT += __shfl_xor_sync(0xffffffffu, cT, o);
                cG += __shfl_xor_sync(0xffffffffu, cG, o); cC += __shfl_xor_sync(0xffffffffu, cC, o);
            }
            if (lane == 0) { red[warp][0] = cA; red[warp][1] = cT; red[warp][2] = cG; red[warp][3] = cC; }
            __syncthreads();
            if (threadIdx.x < 4) {
                uint32_t v = 0;
                for (int q = 0; q < 8; q++) v += red[q][threadIdx.x];
                p.acgt[seq * 4 + threadIdx.x] = seq < p.n ? v : 0u;
            }
            __syncthreads();
        }
    }
}

// ---- DG_INPUT_NIBBLE: two sites per byte -> Paradis bytes -------------------------------------------------------------
// byte = nibble << 4 | (8 if exactly one possibility bit: the base is known); nibble 0 -> 0 (invalid, reported by the
// pack kernels).  One thread per 4 input bytes (8 sites); HBM-bound and tiny next to the PCIe time it saves.
__device__ __forceinline__ uint32_t nib_to_codes2(uint32_t b) {   // one nibble byte -> two Paradis bytes (low nibble first)
    const uint32_t m0 = b & 15u, m1 = (b >> 4) & 15u;
    const uint32_t c0 = (m0 << 4) | (__popc(m0) == 1 ? 8u : 0u), c1 = (m1 << 4) | (__popc(m1) == 1 ? 8u : 0u);
    return c0 | (c1 << 8);
}
// One thread per 8 input bytes (16 sites).  Output rows are `width` bytes apart (29,903: no alignment), so a thread writes
// the bytes up to its first 4-byte boundary singly, then three aligned words cut out of its 16 bytes with funnel shifts,
// then the rest singly: 7 stores instead of 16.
__global__ void nibble_unpack_kernel(const uint8_t* __restrict__ nib, uint8_t* __restrict__ codes, uint64_t n, uint64_t width,
                                     uint64_t wb) {
    const uint64_t per_row = (wb + 7) / 8;
    const uint64_t total = n * per_row;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = u / per_row, k0 = (u - r * per_row) * 8;
        const uint8_t* src = nib + r * wb + k0;
        uint8_t* dst = codes + r * width + 2 * k0;
        const uint32_t n_in = (uint32_t)min((uint64_t)8, wb - k0);
        const uint32_t n_out = (uint32_t)min((uint64_t)16, width - 2 * k0);
        uint32_t w[5];
        if (n_in == 8 && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
            const uint2 v = __ldcs(reinterpret_cast<const uint2*>(src));
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t x = k < 2 ? v.x : v.y;
                const uint32_t b0 = (x >> (16 * (k & 1))) & 0xFFu, b1 = (x >> (16 * (k & 1) + 8)) & 0xFFu;
                w[k] = nib_to_codes2(b0) | (nib_to_codes2(b1) << 16);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t b0 = 2 * k < (int)n_in ? src[2 * k] : 0xFFu, b1 = 2 * k + 1 < (int)n_in ? src[2 * k + 1] : 0xFFu;
                w[k] = nib_to_codes2(b0) | (nib_to_codes2(b1) << 16);
            }
        }
        w[4] = 0;
        const uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
        if (n_out == 16) {
            for (uint32_t i = 0; i < head; i++) dst[i] = (uint8_t)(w[0] >> (8 * i));
            uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
#pragma unroll
            for (int j = 0; j < 3; j++) dw[j] = __funnelshift_r(w[j], w[j + 1], 8 * head);
            if (head == 0) {
                dw[3] = w[3];
            } else {
                for (uint32_t i = head + 12; i < 16; i++) dst[i] = (uint8_t)(w[3] >> (8 * (i - 12)));
            }
        } else {   // the last sites of the row
            for (uint32_t i = 0; i < n_out; i++) dst[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
        }
    }
}

// ---- inverted index of partial ambiguity codes (R Y M W S K V H D B) -----------------------------------
// entry = site << 36 | record << 4 | possibility nibble
struct PpIndex {
    uint64_t* entries = nullptr;   // sorted by site (order inside a site is arbitrary)
    uint32_t* site_off = nullptr;  // width + 1 offsets
    uint32_t n_entries = 0;
    uint32_t cap_entries = 0;
    double pair_work = 0;          // sum over sites of |L_s|^2 (cost of the correction)
};

__device__ __forceinline__ bool is_partial(uint32_t c) { return (c & 8u) == 0u && (c & 0xF0u) != 0xF0u; }

__global__ void pp_count_kernel(const uint8_t* codes, uint64_t n, uint64_t width, int ascii, uint32_t* site_cnt) {
    const uint64_t total = n * width;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = codes[u];
        if (ascii) c = c_ascii_lut[c];
        if (is_partial(c)) atomicAdd(site_cnt + (u % width), 1u);
    }
}
// single-block exclusive scan over `width` site counts; also sum of squares
__global__ void pp_scan_kernel(const uint32_t* site_cnt, uint64_t width, uint32_t* site_off, uint32_t* cursor, double* work) {
    __shared__ uint32_t part[1024];
    __shared__ double wpart[1024];
    const uint64_t per = (width + 1023) / 1024;
    const uint64_t b = threadIdx.x * per, e = min(width, b + per);
    uint32_t s = 0; double w = 0;
    for (uint64_t i = b; i < e; i++) { s += site_cnt[i]; w += (double)site_cnt[i] * site_cnt[i]; }
    part[threadIdx.x] = s; wpart[threadIdx.x] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0; double tw = 0;
        for (int i = 0; i < 1024; i++) { const uint32_t t = part[i]; part[i] = run; run += t; tw += wpart[i]; }
        site_off[width] = run;
        *work = tw;
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint64_t i = b; i < e; i++) { site_off[i] = run; cursor[i] = run; run += site_cnt[i]; }
}
// `codes` = the bytes of records rec0 .. rec0 + n - 1; entries carry the GLOBAL record index
__global__ void pp_fill_kernel(const uint8_t* codes, uint64_t n, uint64_t width, int ascii, uint32_t* cursor, uint64_t* entries,
                               uint64_t rec0 = 0) {
    const uint64_t total = n * width;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = codes[u];
        if (ascii) c = c_ascii_lut[c];
        if (is_partial(c)) {
            const uint64_t site = u % width, rec = rec0 + u / width;
            const uint32_t pos = atomicAdd(cursor + site, 1u);
            entries[pos] = (site << 36) | (rec << 4) | (uint64_t)(c >> 4);
        }
    }
}

// hits[begin, end) (unsorted, written by pack_ops_kernel) -> entries grouped by site through the per-site cursor
__global__ void pp_scatter_kernel(const uint64_t* hits, uint32_t begin, uint32_t end, uint32_t* cursor, uint64_t* entries) {
    for (uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
        const uint64_t h = hits[i];
        entries[atomicAdd(cursor + (h >> 36), 1u)] = h;
    }
}

// ---- chunked index of the pipelined all-vs-all session (dg_square_*) ------------------------------------------
// Records arrive in chunks (highest records first); chunk g gets its own per-site offsets off[g][0 .. width] into
// ONE shared entry buffer: off[g][s] = base_g + (exclusive scan of chunk g's site counts), base_g = the entries of
// all earlier chunks (*total, the running count).  The row doubles as the fill cursor: the scan stores
// off[0] = base_g and off[1 + s] = start of site s; pp_fill_kernel (cursor = off + 1) advances off[1 + s] to the
// END of site s = the start of site s + 1, which leaves exactly the width + 1 offsets.
// *work = sum over sites of (count over every chunk so far)^2: the cost of the both-partial repair
// (decides the engine like PpIndex::pair_work).  site_cnt = this chunk's counts, cum_cnt = all chunks so far.
// h_total / h_work / h_inv point into MAPPED pinned host memory: the kernel publishes the running entry total, the
// repair cost and the invalid-byte flag itself, because a tiny cudaMemcpyAsync would queue behind the bulk panel
// copies on the D2H copy engine.
__global__ void publish_invalid_kernel(const unsigned long long* d_inv, unsigned long long* h_inv) {
    *h_inv = *d_inv;
    __threadfence_system();
}
// One block of 1,024 threads.  This kernel sits on the critical path of every upload chunk of a session (pack -> scan ->
// the host sizes the entry buffer), so it is written for latency.
__global__ void __launch_bounds__(1024) pp_scan_chunk_kernel(const uint32_t* __restrict__ site_cnt, uint32_t* __restrict__ cum_cnt,
                                                             uint64_t width, uint32_t* __restrict__ off, uint32_t* total,
                                                             uint32_t* h_total, double* h_work, const unsigned long long* d_inv,
                                                             unsigned long long* h_inv) {
    // One CTA, every global access coalesced: the sites are walked in segments of 16 x 1,024; inside a segment warp w owns
    // the 32-site groups k * 32 + w (k < 16), scans each with shuffles, and the 512 group totals are scanned in two more
    // shuffle levels.  (Thread-contiguous runs made every load its own L1 wavefront: ~100 us on the one SM.)
    constexpr int G = 16;               // 2 G live registers per thread: 1,024 threads leave 64
    __shared__ uint32_t gsum[G * 32];   // group totals -> exclusive prefix inside their block of 32 groups
    __shared__ uint32_t wtot[32];       // totals of the blocks of 32 groups -> exclusive prefix
    __shared__ double wwork[32];
    __shared__ uint32_t s_run;          // offset of the segment's first site
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) { const uint32_t base = *total; s_run = base; off[0] = base; }
    double w = 0;
    for (uint64_t seg0 = 0; seg0 < width; seg0 += (uint64_t)G * 1024) {
        uint32_t c[G], q[G];
#pragma unroll
        for (int k = 0; k < G; k++) {
            const uint64_t i = seg0 + (uint64_t)(k * 32 + warp) * 32 + lane;
            const bool ok = i < width;
            c[k] = ok ? site_cnt[i] : 0u;
            q[k] = ok ? cum_cnt[i] : 0u;
        }
#pragma unroll
        for (int k = 0; k < G; k++) {
            const uint64_t i = seg0 + (uint64_t)(k * 32 + warp) * 32 + lane;
            if (i < width) {
                const uint32_t cc = q[k] + c[k];
                cum_cnt[i] = cc;
                w += (double)cc * cc;   // integers: the order of the sum does not matter
            }
            uint32_t incl = c[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            q[k] = incl - c[k];         // exclusive prefix inside the group
            if (lane == 31) gsum[k * 32 + warp] = incl;
        }
        __syncthreads();
        if (warp < G) {   // warp w scans group totals [32 w, 32 w + 32)
            const uint32_t v = gsum[warp * 32 + lane];
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += u;
            }
            gsum[warp * 32 + lane] = incl - v;
            if (lane == 31) wtot[warp] = incl;
        }
        __syncthreads();
        uint32_t seg_total = 0;
        if (warp == 0) {
            const uint32_t v = lane < G ? wtot[lane] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += u;
            }
            if (lane < G) wtot[lane] = incl - v;
            seg_total = __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncthreads();
        const uint32_t run0 = s_run;
#pragma unroll
        for (int k = 0; k < G; k++) {
            const uint64_t i = seg0 + (uint64_t)(k * 32 + warp) * 32 + lane;
            if (i < width) off[1 + i] = run0 + wtot[k] + gsum[k * 32 + warp] + q[k];
        }
        __syncthreads();
        if (t == 0) s_run = run0 + seg_total;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) wwork[warp] = w;
    __syncthreads();
    if (warp == 0) {
        double ww = wwork[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) ww += __shfl_xor_sync(0xffffffffu, ww, o);
        if (lane == 0) {
            const uint32_t run = s_run;
            *total = run;
            *h_total = run; *h_work = ww; *h_inv = *d_inv;
            __threadfence_system();
        }
    }
}

// DIFF is computed as sum_b U_b(q) V_b(t) = 3 [Sq, St disjoint] whenever at least one code is a known base or
// N-like (tc_engine.cuh header).  When BOTH codes are partial ambiguity codes the product is off by a small
// integer: add (3 * want - got) to the raw accumulator.  ma = row (U side) nibble, mb = column (V side) nibble.
__device__ __forceinline__ int pp_corr(uint32_t ma, uint32_t mb) {
    const uint32_t nma = ~ma & 15u;
    const int cb = __popc(mb);
    const int got = __popc(nma & mb) * (4 - cb) + __popc(nma & ~mb & 15u) * (1 - cb);
    return ((ma & mb) ? 0 : 3) - got;
}

// Add a repair delta to accumulator 0 of the pair at scratch index idx: int32 sums take an atomicAdd; 16-bit sums
// (ACC_MOD16: the value modulo 2^16) a wrap-around add through a CAS on the containing word.
__device__ __forceinline__ void pp_add(void* out, int op, uint64_t idx, int c) {
    if (op == ACC_RAW_I32 || op == ACC_ADD_I32) {
        atomicAdd(reinterpret_cast<int*>(out) + idx, c);
    } else {
        uint32_t* w = reinterpret_cast<uint32_t*>(out) + (idx >> 1);
        const uint32_t sh = (uint32_t)(idx & 1) * 16;
        uint32_t old = *w, assumed;
        do {
            assumed = old;
            const uint32_t nv = (assumed & ~(0xFFFFu << sh)) | ((((assumed >> sh) + (uint32_t)c) & 0xFFFFu) << sh);
            old = atomicCAS(w, assumed, nv);
        } while (old != assumed);
    }
}

struct PpCorrParams {
    const uint64_t* a_entries; uint32_t a_n;          // row alignment's entries (resident rows)
    const int8_t* a_ops; uint32_t a_nplanes, a_wp8, a_vplane0;  // or: scan the V planes of the row records (stream batches)
    const uint64_t* b_entries; const uint32_t* b_off; // column alignment's index
    uint32_t row0, row_end, n_b;
    int square;
    uint32_t s_pitch, s_colbase;                      // scratch rectangle of the panel (TcParams)
    int op;                                           // ACC_RAW_I32 or ACC_MOD16
    void* out;                                        // raw sums of accumulator 0 (3 * DIFF) of the panel
};
__device__ __forceinline__ void pp_fix_row(const PpCorrParams& p, uint32_t row, uint32_t site, uint32_t ma) {
    const uint64_t row_base = (uint64_t)(row - p.row0) * p.s_pitch;
    for (uint32_t k = p.b_off[site]; k < p.b_off[site + 1]; k++) {
        const uint64_t eb = p.b_entries[k];
        const uint32_t col = (uint32_t)((eb >> 4) & 0xFFFFFFFFull);
        if (p.square && col <= row) continue;
        const int c = pp_corr(ma, (uint32_t)(eb & 15));
        if (c) pp_add(p.out, p.op, row_base + (col - p.s_colbase), c);
    }
}
// rows given as index entries (resident alignments)
__global__ void pp_correct_kernel(PpCorrParams p) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < p.a_n; e += gridDim.x * blockDim.x) {
        const uint64_t ea = p.a_entries[e];
        const uint32_t row = (uint32_t)((ea >> 4) & 0xFFFFFFFFull);
        if (row < p.row0 || row >= p.row_end) continue;
        pp_fix_row(p, row, (uint32_t)(ea >> 36), (uint32_t)(ea & 15));
    }
}
// Session variant: the row entries are entries[a_begin, a_end) (the chunks that hold the panel's rows) and the column
// side is the list of chunk indexes 0 .. n_chunks-1 (offsets off + g * off_stride) over the same entry buffer.
struct PpChunkParams {
    const uint64_t* entries;
    const uint32_t* off; uint32_t off_stride, n_chunks;
    uint32_t a_begin, a_end;
    uint32_t row0, row_end;
    uint32_t s_pitch, s_colbase;
    int op;
    void* out;
};
__global__ void pp_correct_chunks_kernel(PpChunkParams p) {
    for (uint32_t e = p.a_begin + blockIdx.x * blockDim.x + threadIdx.x; e < p.a_end; e += gridDim.x * blockDim.x) {
        const uint64_t ea = p.entries[e];
        const uint32_t row = (uint32_t)((ea >> 4) & 0xFFFFFFFFull);
        if (row < p.row0 || row >= p.row_end) continue;
        const uint32_t site = (uint32_t)(ea >> 36), ma = (uint32_t)(ea & 15);
        const uint64_t row_base = (uint64_t)(row - p.row0) * p.s_pitch;
        for (uint32_t g = 0; g < p.n_chunks; g++) {
            const uint32_t* off = p.off + (uint64_t)g * p.off_stride;
            for (uint32_t k = off[site]; k < off[site + 1]; k++) {
                const uint64_t eb = p.entries[k];
                const uint32_t col = (uint32_t)((eb >> 4) & 0xFFFFFFFFull);
                if (col <= row) continue;
                const int c = pp_corr(ma, (uint32_t)(eb & 15));
                if (c) pp_add(p.out, p.op, row_base + (col - p.s_colbase), c);
            }
        }
    }
}
// rows found by scanning their V planes (stream batches have no index): a site holds a partial code iff
// one of its V values is negative; its possibility nibble is then {b : V_b > 0}.  FP4 planes hold E2M1 nibbles
// (two sites per byte): negative <=> bit 3, positive <=> non-zero without bit 3.
template <bool FP4>
__global__ void pp_correct_scan_kernel(PpCorrParams p) {
    const uint32_t gpr = p.a_wp8 / 16;   // 16-byte groups per record and plane
    constexpr uint32_t SITES = FP4 ? 32 : 16;
    const uint64_t total = (uint64_t)(p.row_end - p.row0) * gpr;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = p.row0 + (uint32_t)(u / gpr);
        const uint32_t g = (uint32_t)(u % gpr);
        const int8_t* base = p.a_ops + ((uint64_t)row * p.a_nplanes + p.a_vplane0) * p.a_wp8 + (uint64_t)g * 16;
        uint4 v[4];
#pragma unroll
        for (int b = 0; b < 4; b++) v[b] = *reinterpret_cast<const uint4*>(base + (uint64_t)b * p.a_wp8);
        const uint32_t neg = FP4 ? 0x88888888u : 0x80808080u;
        const uint32_t any = (v[0].x | v[1].x | v[2].x | v[3].x | v[0].y | v[1].y | v[2].y | v[3].y |
                              v[0].z | v[1].z | v[2].z | v[3].z | v[0].w | v[1].w | v[2].w | v[3].w) & neg;
        if (!any) continue;
        const uint8_t* vb[4] = {reinterpret_cast<const uint8_t*>(&v[0]), reinterpret_cast<const uint8_t*>(&v[1]),
                                reinterpret_cast<const uint8_t*>(&v[2]), reinterpret_cast<const uint8_t*>(&v[3])};
        for (uint32_t k = 0; k < SITES; k++) {
            uint32_t ma = 0, anyneg = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                uint32_t x, isneg;
                if (FP4) { x = (vb[b][k >> 1] >> ((k & 1) * 4)) & 15u; isneg = x >> 3; }
                else { x = vb[b][k]; isneg = x >> 7; }
                anyneg |= isneg;
                if (x != 0 && !isneg) ma |= 8u >> b;
            }
            if (anyneg) pp_fix_row(p, row, g * SITES + k, ma);
        }
    }
}

// The live-tile list of an upper-triangle panel built on the device (same order as the host loop in
// launch_tc_gemm: bands of `raster_g` row blocks, column-major inside a band), for launches that must not queue a
// host-to-device copy behind bulk uploads.  One block; `n_live` (computed by the host) is only checked.
__global__ void build_tile_list_kernel(uint32_t gx, uint32_t gy, uint32_t raster_g, uint32_t row0, uint32_t row_end,
                                       uint32_t rows_per_block, uint32_t col_block0, uint32_t tn, uint32_t* tiles) {
    __shared__ uint32_t part[1024];
    const uint32_t total = gx * gy;
    const uint32_t per = (total + blockDim.x - 1) / blockDim.x;
    const uint32_t q0 = min(total, threadIdx.x * per), q1 = min(total, q0 + per);
    auto live_code = [&](uint32_t q, uint32_t& code) {
        const uint32_t band = q / (raster_g * gx), r = q - band * (raster_g * gx);
        const uint32_t gb = min(raster_g, gy - band * raster_g);
        const uint32_t bx = r / gb, by = band * raster_g + (r - bx * gb);
        const uint64_t rowS0 = (uint64_t)row0 + (uint64_t)by * rows_per_block;
        const uint64_t rowB0 = (uint64_t)(col_block0 + bx) * tn;
        code = (by << 20) | bx;
        return !(rowS0 >= row_end || rowB0 + tn <= rowS0 + 1);
    };
    uint32_t cnt = 0, code;
    for (uint32_t q = q0; q < q1; q++) cnt += live_code(q, code) ? 1u : 0u;
    part[threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < blockDim.x; i++) { const uint32_t t = part[i]; part[i] = run; run += t; }
    }
    __syncthreads();
    uint32_t pos = part[threadIdx.x];
    for (uint32_t q = q0; q < q1; q++)
        if (live_code(q, code)) tiles[pos++] = code;
}

// ---- the GEMM kernel ----------------------------------------------------------------------------------
struct TcParams {
    uint32_t n_b, row0, row_end, col_block0;   // col_block0 in units of tn
    uint32_t tn;                               // tile columns: 256 (kind::i8) or 240 (kind::mxf4: 16 TMEM columns per
                                               // accumulator are left for the scale factors)
    uint32_t gx, gy;                           // tiles: gx column blocks x gy row blocks (of TM * CL rows)
    const uint32_t* tile_list;                 // square panels: the LIVE tiles (by << 20 | bx) in raster order, so that
    uint32_t n_live;                           // the static round-robin deals only real work; NULL = all gx * gy tiles
    int square;
    uint64_t n_total, out_base;
    void* out;
    uint32_t wp8;          // bytes per plane
    uint32_t nsb;          // wp8 / KB
    uint32_t nacc;         // accumulators (integer counts) computed by this launch: work items = nacc x tiles
    uint32_t npairs[5];    // plane pairs summed into each accumulator
    uint8_t pa[5][4], pb[5][4];  // stored-plane index of the A / B operand of each pair
    uint64_t acc_off[5];   // byte offset (from out) of the array accumulator a is stored to
    uint8_t acc_op[5];     // AccOp of accumulator a
    // s_pitch != 0: `out` is a SCRATCH buffer: every accumulator array is the rectangle [panel rows][s_pitch] of the
    // launch's tiles (element (row, col) at (row - row0) * s_pitch + col - s_colbase; s_pitch = gx * tn, so every
    // 32-column chunk of every tile row starts 32-byte aligned and an epilogue lane stores its row's chunk straight from
    // registers with 128-bit stores).  s_pitch == 0: `out` holds the final n / n_high counts in the reference's order.
    uint32_t s_pitch, s_colbase;
    uint32_t raster_g;        // row blocks per raster band (RASTER_G; DG_TC_RASTER overrides for experiments)
    uint32_t l2_evict_last;   // DG_TC_L2_HINT (experiment): operand TMA loads carry an evict-last L2 policy
    uint32_t ksplit;       // >= 1: work items = accumulators x tiles x ksplit, item kc sums K blocks [kc, kc + 1) * KT / ksplit
    uint32_t stages;       // pipeline depth in use (<= STAGES; tuning knob)
    unsigned long long* probe;  // DG_CLOCK_PROBE (debug): [4] += SM clocks, [5] += ns of this launch (CTA 0); or NULL
};

// Tile order: bands of RASTER_G row blocks, column-major inside a band, so the tiles in flight cover a
// near-square patch of the pair matrix and share their A / B operand rows through L2.  With CL = 2 a
// "row block" is the 256-row super-tile of a CTA pair (rank r owns rows [128 r, 128 r + 128) of it).
constexpr uint32_t RASTER_G = 8;
template <int CL, int MT>
__device__ __forceinline__ bool tile_live(const TcParams& p, uint32_t t, uint32_t rank, uint32_t& rowA0, uint32_t& rowB0) {
    uint32_t bx, by;
    if (p.tile_list) {
        const uint32_t code = __ldg(p.tile_list + t);
        bx = code & 0xFFFFFu; by = code >> 20;
    } else {
        const uint32_t G = p.raster_g;
        const uint32_t band = t / (G * p.gx), r = t - band * (G * p.gx);
        const uint32_t gb = min(G, p.gy - band * G);  // row blocks in this band
        bx = r / gb; by = band * G + (r - bx * gb);
    }
    const uint32_t rowS0 = p.row0 + by * (TM * MT * CL);
    rowB0 = (p.col_block0 + bx) * p.tn;
    rowA0 = rowS0 + rank * (TM * MT);
    if (rowS0 >= p.row_end) return false;
    if (p.square && rowB0 + p.tn <= rowS0 + 1) return false;   // decided per super-tile: identical in both CTAs
    return true;
}

// CL = 1: one CTA per tile.  CL = 2: a cluster of two CTAs shares the B tile -- each CTA fetches one
// 128-row half of it and TMA-multicasts it into both CTAs' shared memory, which cuts the L2 -> SM
// operand traffic per MAC by a third; the MMAs stay cta_group::1.  A stage is free again only when
// BOTH CTAs' MMAs have retired (multicast tcgen05.commit onto both `empty` barriers, count 2).
// MT = 2: the CTA owns TWO 128-row A sub-tiles against one B tile (a 256 x 256 block of pairs, two
// accumulators filling all 512 TMEM columns).  The kernel is bound by the operand bytes that fit in
// flight in shared memory (measured: 2 / 3 / 4 stages = 15.9 / 10.8 / 9.5 ms), and this shape needs a
// third fewer staged bytes per MAC; the price is a single-buffered accumulator (the epilogue of a tile
// no longer overlaps the next tile's MMAs, ~2 % of a tile).
// PAIR (with CL = 2): tcgen05 cta_group::2.  The leader CTA issues ONE M = 256 instruction for both SMs;
// each CTA stages its own A sub-tiles and only ITS 128-row half of the B tile (the tensor cores read the
// other half from the peer's shared memory), so a stage is 48 KB instead of 64 KB for the same MACs
// per SM and four stages fit again.  Both CTAs' TMA loads report to the leader's `full` barrier; the
// leader's commits are multicast onto both CTAs' `empty` / `tfull` barriers; both CTAs' epilogue warps
// arrive on the leader's `tempty`.
// FP4 (with PAIR): the same pipeline with tcgen05.mma kind::mxf4.block_scale: operands are E2M1 nibbles (every plane
// value 0, 1, 2, 3, -1, -2 is representable), the UE8M0 scale factors are all 2^0 (one constant block of TMEM columns,
// written once), accumulation is fp32 and exact for these integer sums (|sum| <= 3 x 4 x 29,952 << 2^24; checked on
// the device by tools/ubench_fp4 and by the parity suite).  Twice the MACs per instruction (K = 64) at the same issue
// rate; a tile is 512 x 240 because the scale factors need TMEM columns next to the two accumulators.
// Epilogue of one work item for one epilogue warp: the warp owns TMEM lane quadrant `quad` (32 rows of each 128-row
// sub-tile); tcgen05.ld hands each lane one ROW of a 32-column chunk.
//
// epi_drain_scratch: the output is the tile-aligned scratch rectangle (TcParams::s_pitch), so each lane stores the 32
// values of its row straight from registers: 4 (16-bit) or 8 (32-bit) 128-bit stores, no transpose, no per-row index.
template <int OP, bool FP4, int TNX, int MT>
__device__ __forceinline__ void epi_drain_scratch(const TcParams& p, uint8_t* outb, uint32_t tacc, uint32_t lane, uint32_t rowQ0,
                                                  uint32_t rowB0) {
    using T = typename std::conditional<acc_op_16(OP), uint16_t, uint32_t>::type;
    constexpr uint32_t NCH = (TNX + 31) / 32;
#pragma unroll 1
    for (int m = 0; m < MT; m++) {
        const uint32_t rq0 = rowQ0 + m * TM;   // first row of this warp's quadrant
        if (rq0 >= p.row_end) continue;         // warp-uniform
        const bool rv = rq0 + lane < p.row_end;
        T* const rowp = reinterpret_cast<T*>(outb) + (uint64_t)(rq0 + lane - p.row0) * p.s_pitch + (rowB0 - p.s_colbase);
#pragma unroll 1
        for (uint32_t c = 0; c < NCH; c++) {
            const uint32_t col0 = rowB0 + c * 32;
            if (col0 >= p.n_b) break;                                   // warp-uniform
            if (p.square && col0 + 32 <= rq0 + 1) continue;             // chunk entirely on / below the diagonal
            uint32_t v[32];
            tc_ld32(tacc + m * TN + c * 32, v);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                int x = FP4 ? __float2int_rn(__uint_as_float(v[j])) : (int)v[j];   // fp32 sums are integers
                if (acc_op_div3(OP)) x = (int)((uint32_t)x / 3u);
                v[j] = (uint32_t)x;
            }
            const bool whole = TNX % 32 == 0 || c + 1 < NCH;   // the last chunk of a 240-column tile holds 16 columns
            if (OP == ACC_ADD_I32) {
                if (rv) {
                    int* d = reinterpret_cast<int*>(rowp + c * 32);
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (whole || j < TNX % 32) atomicAdd(d + j, (int)v[j]);
                }
            } else if (rv) {
                uint4* d = reinterpret_cast<uint4*>(rowp + c * 32);
                if (sizeof(T) == 2) {
                    uint32_t w[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) w[k] = (v[2 * k] & 0xFFFFu) | (v[2 * k + 1] << 16);
                    __stcs(d + 0, make_uint4(w[0], w[1], w[2], w[3]));   // streaming stores: scratch is re-read once, much later
                    __stcs(d + 1, make_uint4(w[4], w[5], w[6], w[7]));
                    if (whole) {
                        __stcs(d + 2, make_uint4(w[8], w[9], w[10], w[11]));
                        __stcs(d + 3, make_uint4(w[12], w[13], w[14], w[15]));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) __stcs(d + k, make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                    if (whole) {
#pragma unroll
                        for (int k = 4; k < 8; k++) __stcs(d + k, make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                    }
                }
            }
        }
    }
}

// epi_drain_final: the output holds the final counts in the reference's order (rows of a packed triangle are not
// aligned), so the chunk is transposed through a private conflict-free shared buffer and every store instruction writes
// 32 consecutive results of one row.  Interior chunks (32 whole rows, right of the diagonal) take a fully unrolled
// path: 32 independent LDS + STG with incremental 64-bit offsets.
template <int OP, bool FP4, int TNX, int MT>
__device__ __forceinline__ void epi_drain_final(const TcParams& p, uint8_t* outb, uint32_t* ebuf, uint32_t tacc, uint32_t lane,
                                                uint32_t rowQ0, uint32_t rowB0) {
    using T = typename std::conditional<acc_op_16(OP), uint16_t, uint32_t>::type;
    T* const out = reinterpret_cast<T*>(outb);
#pragma unroll 1
    for (int m = 0; m < MT; m++) {
        const uint32_t rq0 = rowQ0 + m * TM;   // first row of this warp's quadrant
        if (rq0 >= p.row_end) continue;         // warp-uniform
        const uint32_t nrows = min(32u, p.row_end - rq0);
        // index of (rq0, column 0) [rect] / of the virtual element (rq0, rq0 + 1) [square] in the panel
        const uint64_t base0 = p.square ? (uint64_t)rq0 * (2 * p.n_total - rq0 - 1) / 2 - p.out_base
                                        : (uint64_t)(rq0 - p.row0) * p.n_b;
#pragma unroll 1
        for (uint32_t c = 0; c < (TNX + 31) / 32; c++) {
            const uint32_t col0 = rowB0 + c * 32;
            if (col0 >= p.n_b) break;                                   // warp-uniform
            if (p.square && col0 + 32 <= rq0 + 1) continue;             // chunk entirely on / below the diagonal
            uint32_t v[32];
            tc_ld32(tacc + m * TN + c * 32, v);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                uint32_t x = FP4 ? (uint32_t)__float2int_rn(__uint_as_float(v[j])) : v[j];   // fp32 sums are integers
                if (acc_op_div3(OP)) x /= 3u;
                ebuf[lane * EPI_PITCH + j] = x;
            }
            __syncwarp();
            const uint32_t col = col0 + lane;
            const bool lv = col < p.n_b && (TNX % 32 == 0 || c * 32 + lane < TNX);
            if (nrows == 32 && (!p.square || col0 > rq0 + 31)) {
                // interior chunk: every row is whole and right of the diagonal
                uint64_t off = base0 + (p.square ? (uint64_t)(col - rq0 - 1) : (uint64_t)col);
                const uint32_t step0 = p.square ? (uint32_t)(p.n_total - rq0 - 2) : p.n_b;   // off(r + 1) - off(r) = step0 - r | n_b
                if (p.square) {
#pragma unroll
                    for (uint32_t r = 0; r < 32; r++) {
                        const uint32_t val = ebuf[r * EPI_PITCH + lane];
                        if (lv) out[off] = (T)val;
                        off += (uint64_t)(step0 - r);
                    }
                } else {
#pragma unroll
                    for (uint32_t r = 0; r < 32; r++) {
                        const uint32_t val = ebuf[r * EPI_PITCH + lane];
                        if (lv) out[off] = (T)val;
                        off += (uint64_t)step0;
                    }
                }
            } else {
                uint64_t rb = base0;
#pragma unroll 4
                for (uint32_t r = 0; r < nrows; r++) {
                    const uint32_t row = rq0 + r;
                    const uint32_t val = ebuf[r * EPI_PITCH + lane];
                    if (lv && (!p.square || col > row)) out[p.square ? rb + (col - row - 1) : rb + col] = (T)val;
                    rb += p.square ? (uint64_t)(p.n_total - row - 1) : (uint64_t)p.n_b;
                }
            }
            __syncwarp();
        }
    }
}

template <int CL, int MT, bool PAIR = false, bool FP4 = false>
__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
    static_assert(!FP4 || PAIR, "the FP4 variant exists for the cta_group::2 kernel only");
    constexpr int TNX = FP4 ? TN_FP4 : TN;               // tile columns
    constexpr int BH_BYTES = (TNX / 2) * KB;              // bytes of one CTA's half of the B tile (PAIR)
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    static_assert(!PAIR || CL == 2, "cta_group::2 needs a 2-CTA cluster");
    constexpr int STAGE_BYTES = StageCfg<MT, PAIR>::BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
    uint64_t* full = bars;               // [STAGES]
    uint64_t* empty = bars + STAGES;     // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES; // [2]
    uint64_t* tempty = tfull + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ntiles = p.tile_list ? p.n_live : p.gx * p.gy;
    const uint32_t nwork = ntiles * p.nacc * p.ksplit;   // accumulator-major: concurrent tiles read the same planes (L2 reuse)
    // work item w -> accumulator a, tile t, K blocks [k0, k1) of the accumulator's npairs[a] * nsb
    auto decode = [&](uint32_t w, uint32_t& a, uint32_t& t, uint32_t& k0, uint32_t& k1) {
        const uint32_t per = ntiles * p.ksplit;
        a = w / per;
        const uint32_t r = w - a * per;
        t = r / p.ksplit;
        const uint32_t kc = r - t * p.ksplit, KT = p.npairs[a] * p.nsb;
        k0 = (uint32_t)((uint64_t)kc * KT / p.ksplit);
        k1 = (uint32_t)((uint64_t)(kc + 1) * KT / p.ksplit);
    };
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    const uint32_t cid = blockIdx.x / CL, ncl = gridDim.x / CL;   // tiles are dealt to clusters
    constexpr uint16_t MC_MASK = (1u << CL) - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, PAIR ? 1 : CL); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull + b, 1); mbar_init(tempty + b, PAIR ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (FP4) {
        // scale factors: UE8M0 0x7F = 2^0 in every byte of TMEM columns [SF_COL, SF_COL + 16) on all 128 lanes of
        // this CTA, so every block scale of A and B is 1 whatever the scale-factor layout
        if (warp >= 2) {
            const uint32_t one = 0x7F7F7F7Fu;
            for (uint32_t c = 0; c < 16; c++)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};"
                             ::"r"(tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + SF_COL + c), "r"(one) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
    }
    if (CL > 1) cluster_sync_all();   // the peer's barriers (and scale factors) exist before anything is sent to them
    tc_fence_after();
    unsigned long long probe_clk = 0, probe_ns = 0;
    if (p.probe && blockIdx.x == 0 && threadIdx.x == 0) {
        probe_clk = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(probe_ns));
    }

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t policy = p.l2_evict_last ? l2_policy_evict_last() : 0ull;
            for (uint32_t w = cid; w < nwork; w += ncl) {
                uint32_t a, t, k0, k1;
                decode(w, a, t, k0, k1);
                uint32_t rowA0, rowB0;
                if (!tile_live<CL, MT>(p, t, rank, rowA0, rowB0)) continue;
                for (uint32_t kt = k0; kt < k1; kt++) {
                    mbar_wait(empty + stage, phase ^ 1);
                    const uint32_t pr = kt / p.nsb, sb = kt - pr * p.nsb;
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    const int xb = (int)(p.pb[a][pr] * p.wp8 + sb * KB);
                    if (PAIR) {
                        // the leader's barrier counts the bytes of both CTAs
                        if (rank == 0) mbar_arrive_expect_tx(full + stage, 2 * (MT * A_BYTES + BH_BYTES));
                        if (p.l2_evict_last) {
#pragma unroll
                            for (int m = 0; m < MT; m++)
                                tma_load_2d_pair_hint(sa + m * A_BYTES, &tmA, (int)(p.pa[a][pr] * p.wp8 + sb * KB), (int)(rowA0 + m * TM), full + stage, policy);
                            tma_load_2d_pair_hint(sa + MT * A_BYTES, &tmB, xb, (int)(rowB0 + rank * (TNX / 2)), full + stage, policy);
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
#pragma unroll
                        for (int m = 0; m < MT; m++)
                            tma_load_2d_pair(sa + m * A_BYTES, &tmA, (int)(p.pa[a][pr] * p.wp8 + sb * KB), (int)(rowA0 + m * TM), full + stage);
                        tma_load_2d_pair(sa + MT * A_BYTES, &tmB, xb, (int)(rowB0 + rank * (TNX / 2)), full + stage);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(full + stage, STAGE_BYTES);
#pragma unroll
                    for (int m = 0; m < MT; m++)
                        tma_load_2d(sa + m * A_BYTES, &tmA, (int)(p.pa[a][pr] * p.wp8 + sb * KB), (int)(rowA0 + m * TM), full + stage);
                    uint8_t* sbp = sa + MT * A_BYTES;
                    if (CL == 1) {  // both 128-row halves of the B tile
                        tma_load_2d(sbp, &tmB, xb, (int)rowB0, full + stage);
                        tma_load_2d(sbp + B_BYTES / 2, &tmB, xb, (int)(rowB0 + TN / 2), full + stage);
                    } else {        // my half, delivered to both CTAs of the cluster
                        tma_load_2d_mc(sbp + rank * (B_BYTES / 2), &tmB, xb, (int)(rowB0 + rank * (TN / 2)),
                                       full + stage, MC_MASK);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (PAIR: the leader CTA issues for both) =====
        if (lane == 0 && (!PAIR || rank == 0)) {
            uint32_t stage = 0, phase = 0, it = 0;
            for (uint32_t w = cid; w < nwork; w += ncl) {
                uint32_t a, t, k0, k1;
                decode(w, a, t, k0, k1);
                uint32_t rowA0, rowB0;
                if (!tile_live<CL, MT>(p, t, rank, rowA0, rowB0)) continue;
                // MT = 1: two accumulators of 256 columns alternate; MT = 2: both are used by every tile
                const uint32_t ab = MT == 1 ? (it & 1) : 0, aphase = MT == 1 ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(tempty + ab, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ab * TN;
                for (uint32_t kt = k0; kt < k1; kt++) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t db = make_smem_desc(sa + MT * A_BYTES);
#pragma unroll
                    for (uint32_t k4 = 0; k4 < KB / 32; k4++)
#pragma unroll
                        for (int m = 0; m < MT; m++) {
                            if (FP4) tc_mma_f4_pair(d_tmem + m * TN, make_smem_desc(sa + m * A_BYTES) + 2 * k4, db + 2 * k4, IDESC_F4_PAIR, (kt != k0) || (k4 != 0),
                                                    tmem_base + SF_COL, tmem_base + SF_COL + 8);
                            else if (PAIR) tc_mma_i8_pair(d_tmem + m * TN, make_smem_desc(sa + m * A_BYTES) + 2 * k4, db + 2 * k4, IDESC_I8_PAIR, (kt != k0) || (k4 != 0));
                            else tc_mma_i8(d_tmem + m * TN, make_smem_desc(sa + m * A_BYTES) + 2 * k4, db + 2 * k4, IDESC_I8, (kt != k0) || (k4 != 0));
                        }
                    if (PAIR) tc_commit_pair(empty + stage, MC_MASK);   // frees the stage in both CTAs
                    else if (CL == 1) tc_commit(empty + stage);         // frees the smem stage when these MMAs retire
                    else tc_commit_mc(empty + stage, MC_MASK);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (PAIR) tc_commit_pair(tfull + ab, MC_MASK);  // accumulators ready in both CTAs
                else tc_commit(tfull + ab);                     // accumulator ready for the epilogue
                it++;
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> shared-memory transpose -> coalesced stores, reference order =====
        const uint32_t quad = warp & 3;  // warps 2,3,4,5 -> TMEM lane quadrants 2,3,0,1
        uint32_t* ebuf = reinterpret_cast<uint32_t*>(smem + 192 * 1024 + 256) + quad * (32 * EPI_PITCH);
        uint32_t it = 0;
        for (uint32_t w = cid; w < nwork; w += ncl) {
            uint32_t a, t, k0, k1;
            decode(w, a, t, k0, k1);
            uint32_t rowA0, rowB0;
            if (!tile_live<CL, MT>(p, t, rank, rowA0, rowB0)) continue;
            uint8_t* const outb = reinterpret_cast<uint8_t*>(p.out) + p.acc_off[a];
            const uint32_t ab = MT == 1 ? (it & 1) : 0, aphase = MT == 1 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(tfull + ab, aphase);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((quad * 32u) << 16) + ab * TN;
            const uint32_t rowQ0 = rowA0 + quad * 32;
            if (p.s_pitch) {
                switch (p.acc_op[a]) {   // warp-uniform
                case ACC_RAW_I32: epi_drain_scratch<ACC_RAW_I32, FP4, TNX, MT>(p, outb, tacc, lane, rowQ0, rowB0); break;
                case ACC_RAW_I16: epi_drain_scratch<ACC_RAW_I16, FP4, TNX, MT>(p, outb, tacc, lane, rowQ0, rowB0); break;
                case ACC_MOD16: epi_drain_scratch<ACC_MOD16, FP4, TNX, MT>(p, outb, tacc, lane, rowQ0, rowB0); break;
                case ACC_ADD_I32: epi_drain_scratch<ACC_ADD_I32, FP4, TNX, MT>(p, outb, tacc, lane, rowQ0, rowB0); break;
                case ACC_DIV3_U32: epi_drain_scratch<ACC_DIV3_U32, FP4, TNX, MT>(p, outb, tacc, lane, rowQ0, rowB0); break;
                default: epi_drain_scratch<ACC_DIV3_U16, FP4, TNX, MT>(p, outb, tacc, lane, rowQ0, rowB0); break;
                }
            } else if (p.acc_op[a] == ACC_DIV3_U16) {
                epi_drain_final<ACC_DIV3_U16, FP4, TNX, MT>(p, outb, ebuf, tacc, lane, rowQ0, rowB0);
            } else {
                epi_drain_final<ACC_DIV3_U32, FP4, TNX, MT>(p, outb, ebuf, tacc, lane, rowQ0, rowB0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_remote(tempty + ab, 0);  // the leader's MMA thread waits for both CTAs
                else mbar_arrive(tempty + ab);
            }
            it++;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.probe && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        atomicAdd(p.probe + 4, (unsigned long long)clock64() - probe_clk);
        atomicAdd(p.probe + 5, ns - probe_ns);
    }
    if (CL > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ---- combine: raw int32 sums of the accumulators -> reference counts -> result ------------------------
//   n, n_high : acc0 = 3 DIFF
//   raw, jc69 : acc0 = 3 DIFF, acc1 = SAME
//   k80       : acc0 = CS = SAME + ts, acc1 = X = SAME - ts, acc2 = tv
//   tn93      : acc0 = L, acc1 = PP = SP + P1, acc2 = YY = SY + P2, acc3 = SP - P1, acc4 = SY - P2
enum ResultKind { RES_U32 = 0, RES_F64 = 1, RES_U16 = 2, RES_COUNTS = 3, RES_COUNTS16 = 4 };
struct CombineParams {
    const uint8_t* acc;     // scratch: accumulator a's array starts at acc + acc_off[a] and holds acc_op[a] values,
    uint64_t acc_off[5];    // element (row, col) at (row - row0) * s_pitch + col - s_colbase (TcParams)
    uint8_t acc_op[5];
    uint32_t s_pitch, s_colbase;
    const uint32_t* a_acgt; const uint32_t* b_acgt;
    uint32_t n_b, row0, row_end, col0;
    int square, swap_roles, measure, fam, result;
    int literal;            // DG_EPI_LITERAL (tests): evaluate the f64 epilogues with plain divisions (epi_*_ref)
    uint64_t n_total, out_base;
    void* out;              // uint32 / uint16 / double [pairs], or uint4[pairs] canonical counts (dg_debug_counts)
};
__device__ __forceinline__ int load_acc(const CombineParams& p, int a, uint64_t idx) {
    const uint8_t* b = p.acc + p.acc_off[a];
    switch (p.acc_op[a]) {
    // streaming loads: every scratch value is read exactly once
    case ACC_RAW_I32: case ACC_ADD_I32: return __ldcs(reinterpret_cast<const int*>(b) + idx);
    case ACC_RAW_I16: return (int)__ldcs(reinterpret_cast<const short*>(b) + idx);
    case ACC_DIV3_U32: return (int)__ldcs(reinterpret_cast<const unsigned int*>(b) + idx);
    default: return (int)__ldcs(reinterpret_cast<const unsigned short*>(b) + idx);   // ACC_DIV3_U16, ACC_MOD16
    }
}
// accumulator 0 of n / n_high / raw / jc69 -> DIFF
__device__ __forceinline__ uint32_t diff_of(int op, int a0) {
    if (acc_op_div3(op)) return (uint32_t)a0;
    if (op == ACC_MOD16) return ((uint32_t)a0 % 3u == 0u) ? (uint32_t)a0 / 3u : ((uint32_t)a0 + 65536u) / 3u;
    return (uint32_t)a0 / 3u;
}

// Grid: every CTA walks virtual blocks (256-column strip x row phase).  By default there is one CTA per virtual block;
// DG_COMBINE_PER_SM bounds the grid to that many CTAs per SM (an experiment: a bounded grid leaves room for the next
// panel's persistent GEMM CTA, but at 2 - 3 CTAs per SM the f64 chains run latency-bound and the pass gets slower than
// what the overlap wins back: measured 6.4 ms unbounded vs 6.7 - 7.9 ms bounded on config 3).
// One pair: the accumulators' sums at scratch index sidx -> the reference's counts -> the result at out index idx.
__device__ __forceinline__ void combine_pair(const CombineParams& p, uint32_t row, uint32_t col) {
    const uint64_t idx = p.square ? (uint64_t)row * (2 * p.n_total - row - 1) / 2 - p.out_base + (col - row - 1)
                                  : (uint64_t)(row - p.row0) * p.n_b + col;
    const uint64_t sidx = (uint64_t)(row - p.row0) * p.s_pitch + (col - p.s_colbase);
    const int a0 = load_acc(p, 0, sidx);
    const int a1 = p.fam != FAM_SNP ? load_acc(p, 1, sidx) : 0;
    uint4 cnt = make_uint4(0, 0, 0, 0);
    if (p.fam == FAM_SNP || p.fam == FAM_RAW) {
        cnt = make_uint4(diff_of(p.acc_op[0], a0), (uint32_t)a1, 0, 0);  // {n, same}
    } else if (p.fam == FAM_K80) {
        const int tv = load_acc(p, 2, sidx);
        const int same = (a0 + a1) >> 1, ts = (a0 - a1) >> 1;
        cnt = make_uint4((uint32_t)same, (uint32_t)(ts + tv), (uint32_t)tv, 0);  // {same, ts + tv, tv}
    } else {
        const int yy = load_acc(p, 2, sidx), ww = load_acc(p, 3, sidx), zz = load_acc(p, 4, sidx);
        const int sp = (a1 + ww) >> 1, p1 = (a1 - ww) >> 1, sy = (yy + zz) >> 1, p2 = (yy - zz) >> 1;
        cnt = make_uint4((uint32_t)a0, (uint32_t)(a0 - sp - sy), (uint32_t)p1, (uint32_t)p2);  // {L, d, P1, P2}
    }
    if (p.result == RES_COUNTS) { reinterpret_cast<uint4*>(p.out)[idx] = cnt; return; }
    if (p.result == RES_COUNTS16) {   // DG_OPT_RESULT_COUNTS: the host evaluates the f64 expressions with its libm
        __stcs(reinterpret_cast<uint2*>(p.out) + idx, make_uint2(cnt.x | (cnt.y << 16), cnt.z | (cnt.w << 16)));
        return;
    }
    if (p.result == RES_U32) { __stcs(reinterpret_cast<uint32_t*>(p.out) + idx, cnt.x); return; }
    if (p.result == RES_U16) { __stcs(reinterpret_cast<unsigned short*>(p.out) + idx, (unsigned short)cnt.x); return; }
    double r;
    if (p.fam == FAM_RAW) r = p.measure == 2 ? epi_raw(cnt.x, cnt.y) : epi_jc69(cnt.x, cnt.y);
    else if (p.fam == FAM_K80) r = epi_k80(cnt.x, cnt.y, cnt.z, p.literal != 0);
    else {
        const uint4 rc = *reinterpret_cast<const uint4*>(p.a_acgt + 4 * (uint64_t)row);
        const uint4 cc = *reinterpret_cast<const uint4*>(p.b_acgt + 4 * (uint64_t)col);
        r = epi_tn93(cnt.x, cnt.y, cnt.z, cnt.w, p.swap_roles ? cc : rc, p.swap_roles ? rc : cc, p.literal != 0);
    }
    __stcs(reinterpret_cast<double*>(p.out) + idx, r);
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) tc_combine_kernel(CombineParams p, uint32_t gx, uint32_t gy) {
    for (uint32_t vb = blockIdx.x; vb < gx * gy; vb += gridDim.x) {
        const uint32_t bx = vb % gx, by = vb / gx;
        const uint32_t col = p.col0 + bx * blockDim.x + threadIdx.x;
        if (col >= p.n_b) continue;
        for (uint32_t row = p.row0 + by; row < p.row_end; row += gy) {
            if (p.square && col <= row) continue;
            combine_pair(p, row, col);
        }
    }
}

// ---- DG_OPT_RESULT_U8: uint16 counts -> uint8 + overflow list ----------------------------------------------------------
// ovf[0] = number of counts >= 255 met (may exceed cap: then the host takes the uint16 panel), entries {index, value}
// from ovf[2].  One thread per 16 counts (two 128-bit loads, one 128-bit store); HBM-bound, 3 bytes per pair.
__global__ void narrow_u8_kernel(const uint16_t* __restrict__ in, uint8_t* __restrict__ out, uint64_t n, uint32_t* ovf, uint32_t cap) {
    const uint64_t groups = n / 16;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= groups; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i0 = g * 16;
        uint32_t w[8];
        const uint32_t cnt = g < groups ? 16u : (uint32_t)(n - i0);   // the last group is the ragged tail
        if (cnt == 0) break;
        if (cnt == 16) {
            const uint4 a = __ldcs(reinterpret_cast<const uint4*>(in + i0)), b = __ldcs(reinterpret_cast<const uint4*>(in + i0) + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t lo = 2 * k < cnt ? in[i0 + 2 * k] : 0u, hi = 2 * k + 1 < cnt ? in[i0 + 2 * k + 1] : 0u;
                w[k] = lo | (hi << 16);
            }
        }
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t v = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
            if (v >= 255u) {
                const uint32_t pos = atomicAdd(ovf, 1u);
                if (pos < cap) { ovf[2 + 2 * pos] = (uint32_t)(i0 + k); ovf[3 + 2 * pos] = v; }
            }
            o[k >> 2] |= min(v, 255u) << ((k & 3) * 8);
        }
        if (cnt == 16) {
            __stcs(reinterpret_cast<uint4*>(out + i0), make_uint4(o[0], o[1], o[2], o[3]));
        } else {
            for (uint32_t k = 0; k < cnt; k++) out[i0 + k] = (uint8_t)(o[k >> 2] >> ((k & 3) * 8));
        }
    }
}

}  // namespace tc
}  // namespace dg
