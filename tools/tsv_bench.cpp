// Throughput of the ordered TSV writer (host/tsv.cpp) on synthetic panels, no GPU needed.
// build: g++ -O2 -std=c++17 -pthread -o tools/tsv_bench tools/tsv_bench.cpp distance_b200/csrc/host/tsv.cpp distance_b200/csrc/host/fasta.cpp
// usage: tools/tsv_bench [n=20000] [kind: u16|f64] [threads] > /dev/null   (report on stderr)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <unistd.h>
#include <vector>
#include "../distance_b200/csrc/host/tsv.hpp"
int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 20000;
    const bool f64 = argc > 2 && !strcmp(argv[2], "f64");
    const int threads = argc > 3 ? atoi(argv[3]) : 8;
    std::vector<std::string> ids(n);
    for (uint64_t i = 0; i < n; i++) { char b[32]; snprintf(b, sizeof b, "s%06llu", (unsigned long long)i); ids[i] = b; }
    host::TsvWriter w(1, threads);
    w.set_ids(&ids, &ids);
    w.write_header();
    const uint64_t rows_per_panel = 1536;
    std::vector<uint16_t> u(rows_per_panel * n);
    std::vector<double> d(f64 ? rows_per_panel * n : 0);
    for (size_t i = 0; i < u.size(); i++) u[i] = (uint16_t)(i * 2654435761u % 90);
    for (size_t i = 0; i < d.size(); i++) d[i] = (double)(i * 2654435761u % 1000) / 29903.0;
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t lines = 0;
    for (uint64_t r0 = 0; r0 + 1 < n; r0 += rows_per_panel) {
        dg_panel p{};
        p.mode = DG_MODE_SQUARE; p.result_kind = f64 ? DG_RESULT_F64 : DG_RESULT_U16;
        p.row_begin = r0; p.row_end = std::min(n - 1, r0 + rows_per_panel); p.n_cols = n;
        uint64_t cnt = 0;
        for (uint64_t i = p.row_begin; i < p.row_end; i++) cnt += n - 1 - i;
        p.n_results = cnt; p.data = f64 ? (const void*)d.data() : (const void*)u.data();
        w.write_panel(p);
        lines += cnt;
    }
    w.flush();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "{\"n\": %llu, \"kind\": \"%s\", \"threads\": %d, \"lines\": %llu, \"seconds\": %.3f, \"lines_per_s\": %.4g}\n",
            (unsigned long long)n, f64 ? "f64" : "u16", threads, (unsigned long long)lines, s, lines / s);
    return 0;
}
