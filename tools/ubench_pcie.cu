// PCIe micro-benchmark: pinned host <-> device copies, each direction alone and both at once
// (the e2e step moves 598 MB up and 400 MB down; is the sum or the max the floor?).
// build: nvcc -O3 -o tools/ubench_pcie tools/ubench_pcie.cu ; run: tools/ubench_pcie [MB]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 512;
    const size_t bytes = mb << 20;
    void *h_up, *h_dn, *d_up, *d_dn;
    CK(cudaHostAlloc(&h_up, bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_dn, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d_up, bytes));
    CK(cudaMalloc(&d_dn, bytes));
    cudaStream_t su, sd;
    CK(cudaStreamCreateWithFlags(&su, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
    cudaEvent_t u0, u1, d0, d1;
    CK(cudaEventCreate(&u0)); CK(cudaEventCreate(&u1)); CK(cudaEventCreate(&d0)); CK(cudaEventCreate(&d1));
    float up_alone = 0, dn_alone = 0, up_both = 0, dn_both = 0;
    for (int it = 0; it < 4; it++) {
        CK(cudaEventRecord(u0, su)); CK(cudaMemcpyAsync(d_up, h_up, bytes, cudaMemcpyHostToDevice, su)); CK(cudaEventRecord(u1, su));
        CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&up_alone, u0, u1));
        CK(cudaEventRecord(d0, sd)); CK(cudaMemcpyAsync(h_dn, d_dn, bytes, cudaMemcpyDeviceToHost, sd)); CK(cudaEventRecord(d1, sd));
        CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&dn_alone, d0, d1));
        CK(cudaEventRecord(u0, su)); CK(cudaEventRecord(d0, sd));
        CK(cudaMemcpyAsync(d_up, h_up, bytes, cudaMemcpyHostToDevice, su));
        CK(cudaMemcpyAsync(h_dn, d_dn, bytes, cudaMemcpyDeviceToHost, sd));
        CK(cudaEventRecord(u1, su)); CK(cudaEventRecord(d1, sd));
        CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&up_both, u0, u1)); CK(cudaEventElapsedTime(&dn_both, d0, d1));
    }
    // chunked both: 24 MB pieces each way, like the pipelined session
    const size_t piece = 24u << 20;
    CK(cudaEventRecord(u0, su)); CK(cudaEventRecord(d0, sd));
    for (size_t o = 0; o + piece <= bytes; o += piece) {
        CK(cudaMemcpyAsync((char*)d_up + o, (char*)h_up + o, piece, cudaMemcpyHostToDevice, su));
        CK(cudaMemcpyAsync((char*)h_dn + o, (char*)d_dn + o, piece, cudaMemcpyDeviceToHost, sd));
    }
    CK(cudaEventRecord(u1, su)); CK(cudaEventRecord(d1, sd));
    float upc = 0, dnc = 0;
    CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&upc, u0, u1)); CK(cudaEventElapsedTime(&dnc, d0, d1));
    const double gb = bytes / 1e9, gbc = (bytes / piece * piece) / 1e9;
    printf("{\"mb\": %zu, \"h2d_alone_gbs\": %.1f, \"d2h_alone_gbs\": %.1f, \"h2d_concurrent_gbs\": %.1f, \"d2h_concurrent_gbs\": %.1f, "
           "\"sum_concurrent_gbs\": %.1f, \"h2d_concurrent_24mb_pieces_gbs\": %.1f, \"d2h_concurrent_24mb_pieces_gbs\": %.1f}\n",
           mb, gb / up_alone * 1e3, gb / dn_alone * 1e3, gb / up_both * 1e3, gb / dn_both * 1e3,
           gb / up_both * 1e3 + gb / dn_both * 1e3, gbc / upc * 1e3, gbc / dnc * 1e3);
    return 0;
}
