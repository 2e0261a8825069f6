// ubench_fp4.cu -- can the block-scaled FP4 tensor path (tcgen05.mma kind::mxf4, E2M1 operands, UE8M0 scale = 1)
// replace kind::i8 for the pairwise counts?  Two questions, answered on the device:
//   1. EXACTNESS: with operand values in {0, 1, 2, 3, -1, -2} (all E2M1-representable) and scale factors 2^0, is
//      the fp32 accumulation in TMEM exact for integer sums up to ~3.6e5 over K = 119,808 (4 planes x 29,952 sites)?
//   2. RATE: MMAs of M128 N256 K64 issued back to back from resident shared memory on every SM.
// Row r of A holds the constant a_r, row c of B the constant b_c, so D[r][c] must equal K * a_r * b_c exactly.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp4 tools/ubench_fp4.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int TM = 128, TN = 256, KB = 128;   // 128 B per row and stage = 256 fp4 elements = 4 MMAs of K = 64
// block-scaled instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptorBlockScaled):
// a_format [7,10) = b_format [10,13) = 1 (E2M1), K-major, n_dim [17,23) = N >> 3, scale_format [23] = 1 (UE8M0),
// m_dim [24,29) = M >> 4, a_sf_id [29,31) = b_sf_id [4,6) = 0, k_size [31] = 0 (K = 64)
constexpr uint32_t IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | (1u << 23) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t a) {
    return ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (((a >> 4) & 0x3FFFu) | (1u << 16));
}
__device__ __forceinline__ uint32_t e2m1(int v) {   // E2M1 nibble of an integer in {-2..3}
    switch (v) { case 0: return 0x0; case 1: return 0x2; case 2: return 0x4; case 3: return 0x5; case -1: return 0xA; default: return 0xC; }
}
__host__ __device__ inline int a_val(int r) { return r % 3 == 0 ? 0 : 1; }
__host__ __device__ inline int b_val(int c) { const int t[6] = {3, 2, 1, 0, -1, -2}; return t[c % 6]; }

__global__ void __launch_bounds__(128, 1) k(int iters, int stages, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    // every stage holds the same tiles: A rows then B rows, 128 B per row, 8-row groups 1024 B apart (swizzle-agnostic:
    // a row is constant)
    for (int i = threadIdx.x; i < stages * (TM + TN) * KB / 4; i += blockDim.x) {
        const int row = (i * 4 / KB) % (TM + TN);
        const int v = row < TM ? a_val(row) : b_val(row - TM);
        const uint32_t nib = e2m1(v);
        ((uint32_t*)smem)[i] = nib * 0x11111111u;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(su32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // scale factors: UE8M0 0x7F = 2^0 in every byte of columns [256, 288) on all 128 lanes (layout-agnostic)
    {
        const uint32_t lane_base = (threadIdx.x >> 5) * 32;
        const uint32_t one = 0x7F7F7F7Fu;
        for (int c = 0; c < 32; c++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + (lane_base << 16) + 256 + c), "r"(one) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; it++) {
            const uint32_t sa = su32(smem + (it % stages) * (TM + TN) * KB);
            const uint64_t da = mkdesc(sa), db = mkdesc(sa + TM * KB);
            for (uint32_t k4 = 0; k4 < 4; k4++) {
                uint32_t acc = (it | k4) != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                             ::"r"(tmem), "l"(da + 2 * k4), "l"(db + 2 * k4), "r"(IDESC), "r"(acc), "r"(tmem + 256), "r"(tmem + 272) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(su32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (out && blockIdx.x == 0) {   // D[r][c]: lane = row r, column = c
        const uint32_t r = threadIdx.x;
        for (int c0 = 0; c0 < TN; c0 += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tmem + ((r & ~31u) << 16) + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; j++) out[r * TN + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, stages = 4;
    const int smem = stages * (TM + TN) * KB + 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    float* d_out; CK(cudaMalloc(&d_out, TM * TN * 4));
    // ---- exactness ----
    long long bad_total = 0; double worst = 0;
    const int iters_list[3] = {1, 117, 468};   // K = 256, 29,952, 119,808 sites
    long long first_bad[3] = {-1, -1, -1};
    for (int t = 0; t < 3; t++) {
        const int iters = iters_list[t];
        CK(cudaMemset(d_out, 0, TM * TN * 4));
        k<<<1, 128, smem>>>(iters, stages, d_out); CK(cudaDeviceSynchronize());
        std::vector<float> h(TM * TN);
        CK(cudaMemcpy(h.data(), d_out, TM * TN * 4, cudaMemcpyDeviceToHost));
        const double K = 256.0 * iters;
        for (int r = 0; r < TM; r++)
            for (int c = 0; c < TN; c++) {
                const double want = K * a_val(r) * b_val(c);
                const double err = (double)h[r * TN + c] - want;
                if (err != 0) { bad_total++; if (first_bad[t] < 0) first_bad[t] = r * TN + c; if (err < 0 ? -err > worst : err > worst) worst = err < 0 ? -err : err; }
            }
        if (first_bad[t] >= 0) {
            const int r = (int)(first_bad[t] / TN), c = (int)(first_bad[t] % TN);
            fprintf(stderr, "iters %d: first mismatch at (%d,%d): got %.1f want %.1f\n", iters, r, c, h[first_bad[t]], K * a_val(r) * b_val(c));
        }
    }
    // ---- rate ----
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    k<<<sms, 128, smem>>>(200, stages, nullptr); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0)); k<<<sms, 128, smem>>>(iters, stages, nullptr); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double macs = (double)sms * iters * 4 * TM * TN * 64;
    printf("{\"device\": \"%s\", \"sms\": %d, \"mma\": \"tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X M128 N256 K64 SS, E2M1 x E2M1, UE8M0 scales = 1\", "
           "\"exactness\": {\"K_sites\": [256, 29952, 119808], \"mismatches\": %lld, \"max_abs_err\": %.1f, \"max_abs_sum\": %.0f}, "
           "\"ms\": %.4f, \"fp4_tops\": %.1f, \"mac_per_clk_per_sm_at_1965\": %.0f}\n",
           p.name, sms, bad_total, worst, 3.0 * 119808, best, 2 * macs / (best * 1e-3) / 1e12, macs / (best * 1e-3) / sms / 1.965e9);
    return 0;
}
