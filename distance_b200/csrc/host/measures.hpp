// measures.hpp -- the f64 expressions of measures.rs evaluated on the HOST from the integer counts of a
// DG_RESULT_COUNTS16 panel (include/distance_gpu.h), with the platform's libm.
//
// Why: Rust's f64::ln is the platform's log (glibc here), CUDA's log may differ from it by an ulp, and one ulp can
// flip the 12th printed decimal of a few lines in 10^8.  With the counts (exact integers from the device) and the same
// expressions in the same order (this translation unit is built with -ffp-contract=off, so nothing is contracted into
// an FMA), the TSV text of jc69 / k80 / tn93 is byte-identical to what the reference prints.  raw is one IEEE division
// and is identical either way.
#pragma once
#include <cmath>
#include <cstdint>

namespace host {

inline double raw_from_counts(uint32_t n, uint32_t same) {
    const uint64_t d = (uint64_t)same + n;              // measures.rs:59-66: d counts SAME and DIFF sites
    return (double)n / (double)d;                       // measures.rs:68
}
inline double jc69_from_counts(uint32_t n, uint32_t same) {
    const double p = raw_from_counts(n, same);          // measures.rs:73
    return -0.75 * std::log(1.0 - (4.0 / 3.0) * p);     // measures.rs:76
}
inline double k80_from_counts(uint32_t same, uint32_t e, uint32_t tv) {
    const uint32_t ts = e - tv;
    const uint64_t count_L = (uint64_t)same + e;        // measures.rs:85-107
    const double P = (double)ts / (double)count_L;
    const double Q = (double)tv / (double)count_L;
    return -0.5 * std::log((1.0 - 2.0 * P - Q) * std::sqrt(1.0 - 2.0 * Q));   // measures.rs:109-112
}
// q / t: A, T, G, C counts of the reference's `query` / `target` record (fastaio.rs:53-66)
inline double tn93_from_counts(uint32_t count_L, uint32_t count_d, uint32_t count_P1, uint32_t count_P2, const uint32_t* q,
                               const uint32_t* t) {
    const uint64_t qA = q[0], qT = q[1], qG = q[2], qC = q[3];
    const uint64_t tA = t[0], tT = t[1], tG = t[2], tC = t[3];
    const uint64_t L = qA + qT + qG + qC + tA + tT + tG + tC;                    // measures.rs:118-125
    const double g_A = ((double)tA + (double)qA) / (double)L;                    // :128-131
    const double g_C = ((double)tC + (double)qC) / (double)L;
    const double g_G = ((double)tG + (double)qG) / (double)L;
    const double g_T = ((double)tT + (double)qT) / (double)L;
    const double g_R = ((double)tA + (double)qA + (double)tG + (double)qG) / (double)L;   // :133-137
    const double g_Y = ((double)tC + (double)qC + (double)tT + (double)qT) / (double)L;   // :139-143
    const double k1 = 2.0 * g_A * g_G / g_R;                                     // :146-148
    const double k2 = 2.0 * g_T * g_C / g_Y;
    const double k3 = 2.0 * (g_R * g_Y - g_A * g_G * g_Y / g_R - g_T * g_C * g_R / g_Y);
    const double P1 = (double)count_P1 / (double)count_L;                        // :178-180
    const double P2 = (double)count_P2 / (double)count_L;
    const double Q = (double)(uint64_t)(count_d - (count_P1 + count_P2)) / (double)count_L;
    const double w1 = 1.0 - P1 / k1 - Q / (2.0 * g_R);                           // :183-185
    const double w2 = 1.0 - P2 / k2 - Q / (2.0 * g_Y);
    const double w3 = 1.0 - Q / (2.0 * g_R * g_Y);
    double d = -k1 * std::log(w1) - k2 * std::log(w2) - k3 * std::log(w3);       // :187
    if (d == 0.0) d = 0.0;                                                       // :188-190
    return d;
}

}  // namespace host
