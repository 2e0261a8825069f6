// Where does the start-up time of a fresh process go?  CUDA runtime / context creation vs. dg_create.
// build: nvcc -O2 -o tools/ubench_init tools/ubench_init.cu -ldl ; run: tools/ubench_init distance_b200/_lib/libdistance_gpu.so
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <dlfcn.h>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    double t = now();
    auto lap = [&](const char* what) { const double n = now(); printf("%-40s %.3f s\n", what, n - t); t = n; };
    int n = 0;
    cudaGetDeviceCount(&n); lap("cudaGetDeviceCount");
    cudaSetDevice(0); lap("cudaSetDevice(0)");
    cudaFree(0); lap("cudaFree(0) (context)");
    cudaStream_t s; cudaStreamCreate(&s); lap("cudaStreamCreate");
    void* d; cudaMalloc(&d, 1 << 20); lap("cudaMalloc 1 MB");
    void* h; cudaHostAlloc(&h, 1 << 20, cudaHostAllocDefault); lap("cudaHostAlloc 1 MB");
    void* h2; cudaHostAlloc(&h2, 256 << 20, cudaHostAllocDefault); lap("cudaHostAlloc 256 MB");
    if (argc > 1) {
        void* lib = dlopen(argv[1], RTLD_NOW); lap("dlopen libdistance_gpu.so");
        if (lib) {
            typedef int (*create_fn)(const int*, int, int, uint64_t, void**);
            create_fn create = (create_fn)dlsym(lib, "dg_create");
            void* ctx = nullptr;
            int rc = create(nullptr, 1, 1, 29903, &ctx); lap("dg_create (1 GPU)");
            printf("rc %d\n", rc);
        }
    }
    return 0;
}
