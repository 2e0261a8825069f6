// ubench_int.cu -- integer-pipe micro-benchmarks for the roofline denominators of the count kernel
// (MEASURED_PEAKS.json only holds HBM and bf16 peaks).  Prints one JSON object.
//   lop3      : independent LOP3 chains            (alu pipe)
//   popc      : independent POPC                   (which pipe? measured)
//   imad_iadd : IMAD x*1+y accumulations           (fma pipe)
//   mix_snp   : 4 LOP3 + 1 POPC + 1 IMAD.IADD      (the n / n_high word-pair recipe)
//   mix_csa   : 4 LOP3 + 3-2 carry-save + 1/2 POPC (Harley-Seal level 1)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_int tools/ubench_int.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 2048;
constexpr int ILP = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t a[ILP], acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; acc[i] = i; }
    const uint32_t b = seed ^ 0x5bd1e995u, c = seed * 7u + 3u, d = ~seed, e = seed + 11u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) {  // 4 dependent LOP3 per slot
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(d), "r"(e));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(a[i]) : "r"(c), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(e), "r"(b));
            } else if (MODE == 1) {  // POPC only
                uint32_t p;
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(a[i]));
                a[i] = p + it;  // keep a dependency (1 IADD per POPC)
            } else if (MODE == 2) {  // IMAD.IADD style accumulation
                asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(acc[i]) : "r"(a[i]));
                asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(acc[i]) : "r"(b));
                asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(acc[i]) : "r"(c));
                asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(acc[i]) : "r"(d));
            } else if (MODE == 3) {  // SNP recipe: 4 LOP3 + POPC + add
                uint32_t x, p;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xC0;" : "=r"(x) : "r"(a[i]), "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0xEA;" : "+r"(x) : "r"(a[i]), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0xEA;" : "+r"(x) : "r"(a[i]), "r"(d));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0x15;" : "+r"(x) : "r"(a[i]), "r"(e));
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x));
                acc[i] += p;
                a[i] += it;
            } else if (MODE == 4) {  // two words: 8 LOP3 + CSA (2 LOP3) + 1 POPC(twos) ; ones carried
                uint32_t x, y, s, cy, p;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xC0;" : "=r"(x) : "r"(a[i]), "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0xEA;" : "+r"(x) : "r"(a[i]), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0xEA;" : "+r"(x) : "r"(a[i]), "r"(d));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0x15;" : "+r"(x) : "r"(a[i]), "r"(e));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xC0;" : "=r"(y) : "r"(a[i]), "r"(d), "r"(e));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0xEA;" : "+r"(y) : "r"(a[i]), "r"(b));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0xEA;" : "+r"(y) : "r"(a[i]), "r"(e));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0x15;" : "+r"(y) : "r"(a[i]), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(acc[i]), "r"(x), "r"(y));   // sum
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(cy) : "r"(acc[i]), "r"(x), "r"(y));  // carry
                acc[i] = s;
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(cy));
                a[i] += p;
            }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= a[i] ^ acc[i];
    if (r == 0x12345678u) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
int run(const char* name, double ops_per_slot, int sms, bool last) {
    uint32_t* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8));
    const int blocks = sms * 8;  // 8 CTAs x 256 threads = 64 warps / SM
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++) k<MODE><<<blocks, 256>>>(out, 12345u + w, cyc);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k<MODE><<<blocks, 256>>>(out, 777u + rep, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double slots = (double)blocks * 256 * ITERS * ILP;
    const double lane_ops = slots * ops_per_slot;
    const double per_s = lane_ops / (best * 1e-3);
    const double clk_mhz = (double)h / (best * 1e-3) / 1e6;
    printf("  \"%s\": {\"ms\": %.4f, \"lane_ops_per_s\": %.4e, \"slots_per_s\": %.4e, \"ops_per_slot\": %.1f, "
           "\"lane_ops_per_clk_per_sm\": %.2f, \"sm_clock_mhz_est\": %.0f}%s\n",
           name, best, per_s, slots / (best * 1e-3), ops_per_slot, per_s / (clk_mhz * 1e6) / sms, clk_mhz, last ? "" : ",");
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", p.name, p.multiProcessorCount, p.clockRate);
    int sms = p.multiProcessorCount;
    if (run<0>("lop3", 4, sms, false)) return 1;
    if (run<1>("popc_plus_iadd", 2, sms, false)) return 1;
    if (run<2>("imad_iadd", 4, sms, false)) return 1;
    if (run<3>("mix_snp_4lop3_1popc_1add", 6, sms, false)) return 1;   // + 1 loop-carried add on a[i]
    if (run<4>("mix_csa_2words", 12, sms, true)) return 1;             // algorithmic ops of 2 word-pairs
    printf("}\n");
    return 0;
}
