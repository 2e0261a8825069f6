// ubench_tc.cu -- tcgen05 kind::i8 issue-rate micro-benchmark (the int8 tensor roofline denominator;
// MEASURED_PEAKS.json only holds bf16).  Every CTA (one per SM) issues ITERS x 4 MMAs of M=128 N=256 K=32
// (SS mode: both operands from shared memory, 128B-swizzled K-major tiles that are never reloaded), so
// the figure is the tensor pipe + shared-memory operand-read ceiling without any TMA / L2 traffic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tc tools/ubench_tc.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int TM = 128, TN = 256, KB = 128;
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t a) {
    return ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (((a >> 4) & 0x3FFFu) | (1u << 16));
}

__global__ void __launch_bounds__(128, 1) k(int iters, int stages) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < stages * (TM + TN) * KB / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01000100u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(su32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; it++) {
            const uint32_t sa = su32(smem + (it % stages) * (TM + TN) * KB);
            const uint64_t da = mkdesc(sa), db = mkdesc(sa + TM * KB);
            for (uint32_t k4 = 0; k4 < 4; k4++) {
                uint32_t acc = (it | k4) != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + (it & 1) * TN), "l"(da + 2 * k4), "l"(db + 2 * k4), "r"(IDESC), "r"(acc) : "memory");
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
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, iters = 20000, stages = 4;
    const int smem = stages * (TM + TN) * KB + 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<sms, 128, smem>>>(200, stages); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0)); k<<<sms, 128, smem>>>(iters, stages); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double macs = (double)sms * iters * 4 * TM * TN * 32;
    printf("{\"device\": \"%s\", \"sms\": %d, \"mma\": \"tcgen05.mma.cta_group::1.kind::i8 M128 N256 K32 SS\", \"ms\": %.4f, "
           "\"int8_tops\": %.1f, \"mac_per_clk_per_sm_at_1965\": %.0f}\n",
           p.name, sms, best, 2 * macs / (best * 1e-3) / 1e12, macs / (best * 1e-3) / sms / 1.965e9);
    return 0;
}
