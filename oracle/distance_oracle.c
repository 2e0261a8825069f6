/*
 * distance_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the pairwise-comparison hot path of
 * benjamincjackson/distance v0.3.1 (Rust), used as the checker for the CUDA
 * path.  Only tests/ (incl. the fixture generator tests/golden/make_golden.py),
 * __graft_entry__.smoke() and the timed CPU baselines (bench.py's cpu_baseline /
 * `--impl reference` legs; the same bounded-sample baseline in the measurement
 * script tools/run_configs.py) may load this library.  The product
 * (libdistance_gpu, the `distance` binary, distance_b200/) never links, loads or
 * calls anything in this directory.
 *
 * PARITY PIN: the reference is Rust and there is no Rust toolchain in this
 * image, so the reference itself cannot be run here (oracle/_ref does not
 * exist; cpu_baseline.kind is "port").  The oracle is pinned instead against
 * every golden vector the reference's own tests hold for this path
 * (tests/test_oracle_golden.py; SURVEY.md section 8c lists them with file:line).
 *
 * Every function cites the reference lines it follows.  Loops are byte-wise,
 * one pair at a time, in the same branch order and the same f64 expression
 * order as the reference.  Build with -ffp-contract=off (see oracle/Makefile)
 * so no FMA contraction changes an f64 result; ln/sqrt come from libm exactly
 * as Rust's f64::ln / f64::sqrt do on linux-gnu.
 *
 * Rust precedence reminder: `a & b < 16` is `(a & b) < 16` and `a & 8 == 8`
 * is `(a & 8) == 8` in Rust, the opposite of C.  Every such test below is
 * parenthesised explicitly.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define OR_API __attribute__((visibility("default")))

/* measure ids shared with include/distance_gpu.h (lib.rs:477-488 dispatch) */
enum { OR_N = 0, OR_N_HIGH = 1, OR_RAW = 2, OR_JC69 = 3, OR_K80 = 4, OR_TN93 = 5 };

/* ------------------------------------------------------------------------- */
/* encoding.rs:4-41  ASCII -> Paradis 8-bit code, 0 = invalid                 */
/* ------------------------------------------------------------------------- */
OR_API void or_encoding_array(uint8_t a[256]) {
    memset(a, 0, 256);
    a['A'] = 136; a['a'] = 136;
    a['G'] = 72;  a['g'] = 72;
    a['C'] = 40;  a['c'] = 40;
    a['T'] = 24;  a['t'] = 24;
    a['R'] = 192; a['r'] = 192;
    a['M'] = 160; a['m'] = 160;
    a['W'] = 144; a['w'] = 144;
    a['S'] = 96;  a['s'] = 96;
    a['K'] = 80;  a['k'] = 80;
    a['Y'] = 48;  a['y'] = 48;
    a['V'] = 224; a['v'] = 224;
    a['H'] = 176; a['h'] = 176;
    a['D'] = 208; a['d'] = 208;
    a['B'] = 112; a['b'] = 112;
    a['N'] = 240; a['n'] = 240;
    a['-'] = 244;
    a['?'] = 242;
}

/* fastaio.rs:101-118 `encode`: LUT per byte; first invalid char aborts.
 * Returns -1 on success, else the index of the first invalid character
 * (the caller builds the message of fastaio.rs:89-91 from it). */
OR_API int64_t or_encode(const uint8_t *ascii, uint64_t w, uint8_t *out) {
    uint8_t lut[256];
    or_encoding_array(lut);
    for (uint64_t i = 0; i < w; i++) {
        if (lut[ascii[i]] == 0) return (int64_t)i;
        out[i] = lut[ascii[i]];
    }
    return -1;
}

/* fastaio.rs:53-66 `count_bases`: histogram of ENCODED bytes (so case
 * insensitive).  out order = A, T, G, C (the struct's field order). */
OR_API void or_count_bases(const uint8_t *seq, uint64_t w, uint64_t out[4]) {
    uint64_t counting[256];
    memset(counting, 0, sizeof counting);
    for (uint64_t i = 0; i < w; i++) counting[seq[i]] += 1;
    out[0] = counting[136];
    out[1] = counting[24];
    out[2] = counting[72];
    out[3] = counting[40];
}

/* fastaio.rs:120-145 `encode_count_bases` (the -s stream variant): counts the
 * RAW characters 'A','T','G','C' only (upper case), fastaio.rs:139-142. */
OR_API int64_t or_encode_count_bases(const uint8_t *ascii, uint64_t w, uint8_t *out,
                                     uint64_t counts[4]) {
    uint8_t lut[256];
    uint64_t counting[256];
    or_encoding_array(lut);
    memset(counting, 0, sizeof counting);
    for (uint64_t i = 0; i < w; i++) {
        if (lut[ascii[i]] == 0) return (int64_t)i;
        out[i] = lut[ascii[i]];
        counting[ascii[i]] += 1;
    }
    counts[0] = counting['A'];
    counts[1] = counting['T'];
    counts[2] = counting['G'];
    counts[3] = counting['C'];
    return -1;
}

/* fastaio.rs:67-75 `get_differences`: sites where code < 240 and != consensus.
 * Returns the number of indices written to out (capacity w). */
OR_API uint64_t or_get_differences(const uint8_t *seq, const uint8_t *other, uint64_t w,
                                   uint64_t *out) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < w; i++) {
        if ((seq[i] < 240) && (seq[i] != other[i])) out[n++] = i;
    }
    return n;
}

/* fastaio.rs:289-336 `consensus`: per-column argmax over A,G,C,T counts;
 * non-ACGT bytes count as A (lookup default 0, :295-302); ties -> first max in
 * the order A,G,C,T because the scan uses strict `>` (:321-328). seqs is
 * n x w row-major. */
OR_API void or_consensus(const uint8_t *seqs, uint64_t n, uint64_t w, uint8_t *out) {
    uint64_t *counts = (uint64_t *)calloc(w * 4, sizeof(uint64_t));
    size_t lookup[256];
    memset(lookup, 0, sizeof lookup);
    lookup[136] = 0; lookup[72] = 1; lookup[40] = 2; lookup[24] = 3;
    for (uint64_t r = 0; r < n; r++) {
        const uint8_t *s = seqs + r * w;
        for (uint64_t i = 0; i < w; i++) counts[i * 4 + lookup[s[i]]] += 1;
    }
    static const uint8_t back_translate[4] = {136, 72, 40, 24};
    for (uint64_t i = 0; i < w; i++) {
        size_t maxidx = 0;
        uint64_t maxval = 0;
        for (size_t k = 0; k < 4; k++) {
            if (counts[i * 4 + k] > maxval) { maxval = counts[i * 4 + k]; maxidx = k; }
        }
        out[i] = back_translate[maxidx];
    }
    free(counts);
}

/* ------------------------------------------------------------------------- */
/* measures.rs                                                               */
/* ------------------------------------------------------------------------- */

/* measures.rs:14-23 `snp` (-m n_high) */
OR_API int64_t or_snp(const uint8_t *query, const uint8_t *target, uint64_t w) {
    int64_t d = 0;
    for (uint64_t i = 0; i < w; i++) {
        if ((query[i] & target[i]) < 16) d += 1;
    }
    return d;
}

/* Rust slice::binary_search on a sorted u64 slice: Ok(pos) / Err.  Returns pos
 * or -1. (Lists from get_differences are strictly increasing, so any-match =
 * the match.) */
static int64_t bsearch_u64(const uint64_t *a, uint64_t len, uint64_t key) {
    uint64_t lo = 0, hi = len;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (a[mid] == key) return (int64_t)mid;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return -1;
}

/* measures.rs:28-53 `snp_consensus` (-m n).  Includes the literal
 * `start = pos` (:42): pos is relative to the slice [start..], so the window
 * start only ever under-advances. */
OR_API int64_t or_snp_consensus(const uint8_t *query, const uint8_t *target,
                                const uint64_t *qdiff, uint64_t nq,
                                const uint64_t *tdiff, uint64_t nt) {
    int64_t d = 0;
    for (uint64_t k = 0; k < nq; k++) {
        uint64_t idx = qdiff[k];
        if ((query[idx] & target[idx]) < 16) d += 1;
    }
    uint64_t start = 0;
    for (uint64_t k = 0; k < nt; k++) {
        uint64_t idx = tdiff[k];
        int64_t pos = bsearch_u64(qdiff + start, nq - start, idx);
        if (pos >= 0) { start = (uint64_t)pos; continue; }
        if ((query[idx] & target[idx]) < 16) d += 1;
    }
    return d;
}

/* measures.rs:56-66 the `raw` site loop: d = compared sites, n = differences */
OR_API void or_raw_counts(const uint8_t *query, const uint8_t *target, uint64_t w,
                          uint64_t *d_out, uint64_t *n_out) {
    uint64_t d = 0, n = 0;
    for (uint64_t i = 0; i < w; i++) {
        if (((query[i] & 8) == 8) && (query[i] == target[i])) {
            d += 1;
        } else if ((query[i] & target[i]) < 16) {
            d += 1;
            n += 1;
        }
    }
    *d_out = d; *n_out = n;
}

/* measures.rs:56-69 `raw`: n as f64 / d as f64 (:68); d == 0 -> NaN */
OR_API double or_raw(const uint8_t *query, const uint8_t *target, uint64_t w) {
    uint64_t d, n;
    or_raw_counts(query, target, w, &d, &n);
    return (double)n / (double)d;
}

/* measures.rs:72-77 `jc69` */
OR_API double or_jc69(const uint8_t *query, const uint8_t *target, uint64_t w) {
    double p = or_raw(query, target, w);
    return -0.75 * log(1.0 - (4.0 / 3.0) * p);
}

/* measures.rs:85-107 the `k80` site loop */
OR_API void or_k80_counts(const uint8_t *query, const uint8_t *target, uint64_t w,
                          uint64_t *L_out, uint64_t *ts_out, uint64_t *tv_out) {
    uint64_t count_L = 0, ts = 0, tv = 0;
    for (uint64_t i = 0; i < w; i++) {
        uint8_t q = query[i], t = target[i];
        if (((q & 8) == 8) && (q == t)) {
            count_L += 1;
        } else if ((q & t) < 16) {
            if (((q & 55) == 0) && ((t & 55) == 0)) {
                ts += 1; count_L += 1;
            } else if (((q & 199) == 0) && ((t & 199) == 0)) {
                ts += 1; count_L += 1;
            } else if ((((q & 55) == 0) && ((t & 199) == 0)) ||
                       (((q & 199) == 0) && ((t & 55) == 0))) {
                tv += 1; count_L += 1;
            }
        }
    }
    *L_out = count_L; *ts_out = ts; *tv_out = tv;
}

/* measures.rs:80-113 `k80`; epilogue :109-112 */
OR_API double or_k80(const uint8_t *query, const uint8_t *target, uint64_t w) {
    uint64_t count_L, ts, tv;
    or_k80_counts(query, target, w, &count_L, &ts, &tv);
    double P = (double)ts / (double)count_L;
    double Q = (double)tv / (double)count_L;
    return -0.5 * log((1.0 - 2.0 * P - Q) * sqrt(1.0 - 2.0 * Q));
}

/* measures.rs:156-175 the `tn93` site loop */
OR_API void or_tn93_counts(const uint8_t *query, const uint8_t *target, uint64_t w,
                           uint64_t *L_out, uint64_t *d_out, uint64_t *P1_out,
                           uint64_t *P2_out) {
    uint64_t count_P1 = 0, count_P2 = 0, count_d = 0, count_L = 0;
    for (uint64_t i = 0; i < w; i++) {
        uint8_t q = query[i], t = target[i];
        if (((q & 8) == 8) && (q == t)) {
            count_L += 1;
        } else if (((q & t) < 16) && ((q & 8) == 8) && ((t & 8) == 8)) {
            count_d += 1;
            count_L += 1;
            if ((q | t) == 200) count_P1 += 1;
            else if ((q | t) == 56) count_P2 += 1;
        }
    }
    *L_out = count_L; *d_out = count_d; *P1_out = count_P1; *P2_out = count_P2;
}

/* measures.rs:116-193 `tn93`.  qc / tc = per-record base counts in the order
 * A, T, G, C (fastaio.rs:17-20).  Expression order follows :118-148 and
 * :178-190 literally. */
OR_API double or_tn93(const uint8_t *query, const uint8_t *target, uint64_t w,
                      const uint64_t qc[4], const uint64_t tc[4]) {
    uint64_t qA = qc[0], qT = qc[1], qG = qc[2], qC = qc[3];
    uint64_t tA = tc[0], tT = tc[1], tG = tc[2], tC = tc[3];
    uint64_t L = qA + qT + qG + qC + tA + tT + tG + tC;

    double g_A = ((double)tA + (double)qA) / (double)L;
    double g_C = ((double)tC + (double)qC) / (double)L;
    double g_G = ((double)tG + (double)qG) / (double)L;
    double g_T = ((double)tT + (double)qT) / (double)L;
    double g_R = ((double)tA + (double)qA + (double)tG + (double)qG) / (double)L;
    double g_Y = ((double)tC + (double)qC + (double)tT + (double)qT) / (double)L;

    double k1 = 2.0 * g_A * g_G / g_R;
    double k2 = 2.0 * g_T * g_C / g_Y;
    double k3 = 2.0 * (g_R * g_Y - g_A * g_G * g_Y / g_R - g_T * g_C * g_R / g_Y);

    uint64_t count_L, count_d, count_P1, count_P2;
    or_tn93_counts(query, target, w, &count_L, &count_d, &count_P1, &count_P2);

    double P1 = (double)count_P1 / (double)count_L;
    double P2 = (double)count_P2 / (double)count_L;
    double Q = (double)(count_d - (count_P1 + count_P2)) / (double)count_L;

    double w1 = 1.0 - P1 / k1 - Q / (2.0 * g_R);
    double w2 = 1.0 - P2 / k2 - Q / (2.0 * g_Y);
    double w3 = 1.0 - Q / (2.0 * g_R * g_Y);

    double d = -k1 * log(w1) - k2 * log(w2) - k3 * log(w3);
    if (d == 0.0) d = 0.0; /* :187-190 normalises -0.0 to +0.0 */
    return d;
}

/* ------------------------------------------------------------------------- */
/* An "alignment" for the drivers below: n x w Paradis bytes row-major, plus  */
/* the per-record fields the reference precomputes in set_up (lib.rs:219-241) */
/* ------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *seqs;
    uint64_t n, w;
    const uint64_t *acgt;      /* n x 4 (A,T,G,C) or NULL */
    const uint64_t *diff_off;  /* n+1 offsets into diff_idx, or NULL (-m n) */
    const uint64_t *diff_idx;
} or_aln;

/* lib.rs:477-488 + call sites lib.rs:434 / lib.rs:325: one pair -> FloatInt.
 * Returns the float as f64; for n / n_high returns the integer in *iout. */
static double pair_eval(int measure, const or_aln *a, uint64_t i, const or_aln *b, uint64_t j,
                        int64_t *iout) {
    const uint8_t *q = a->seqs + i * a->w;
    const uint8_t *t = b->seqs + j * b->w;
    uint64_t w = a->w;
    switch (measure) {
    case OR_N:
        *iout = or_snp_consensus(q, t, a->diff_idx + a->diff_off[i],
                                 a->diff_off[i + 1] - a->diff_off[i],
                                 b->diff_idx + b->diff_off[j],
                                 b->diff_off[j + 1] - b->diff_off[j]);
        return 0.0;
    case OR_N_HIGH: *iout = or_snp(q, t, w); return 0.0;
    case OR_RAW:  return or_raw(q, t, w);
    case OR_JC69: return or_jc69(q, t, w);
    case OR_K80:  return or_k80(q, t, w);
    case OR_TN93: return or_tn93(q, t, w, a->acgt + 4 * i, b->acgt + 4 * j);
    }
    return NAN;
}

/* lib.rs:626-633: ints `{}`, floats `{:.12}`.  Rust prints NaN as "NaN",
 * infinities as "inf"/"-inf", and keeps the sign of -0.0. */
OR_API int or_format_float12(double d, char *buf, size_t cap) {
    if (isnan(d)) return snprintf(buf, cap, "NaN");
    if (isinf(d)) return snprintf(buf, cap, d > 0 ? "inf" : "-inf");
    return snprintf(buf, cap, "%.12f", d);
}

/* ------------------------------------------------------------------------- */
/* Ordered drivers: results in the order of generate_pairs_square             */
/* (lib.rs:502-547: i in 0..n-1, j in i+1..n), generate_pairs_rectangle       */
/* (lib.rs:551-596: i in 0..n1, j in 0..n2) and the stream worker loop        */
/* (lib.rs:322-325: streamed record major, loaded record minor,               */
/*  f(record_1 = loaded, record_2 = streamed)).                               */
/* Exactly one of fout / iout is written depending on the measure.            */
/* ------------------------------------------------------------------------- */
typedef struct {
    int measure, mode; /* mode 0 square, 1 rect, 2 stream */
    or_aln a, b;
    double *fout; int64_t *iout;
    uint64_t row0, row1;      /* major-index range handled by this thread */
    uint64_t done;
} or_job;

static uint64_t sq_off(uint64_t n, uint64_t i) { /* index of pair (i, i+1) */
    return i * (2 * n - i - 1) / 2;
}

static void *job_run(void *p) {
    or_job *jb = (or_job *)p;
    int64_t iv = 0;
    uint64_t done = 0;
    for (uint64_t r = jb->row0; r < jb->row1; r++) {
        if (jb->mode == 0) {
            uint64_t n = jb->a.n;
            uint64_t base = sq_off(n, r);
            for (uint64_t j = r + 1; j < n; j++) {
                double f = pair_eval(jb->measure, &jb->a, r, &jb->a, j, &iv);
                if (jb->fout) jb->fout[base + (j - r - 1)] = f;
                if (jb->iout) jb->iout[base + (j - r - 1)] = iv;
                done++;
            }
        } else if (jb->mode == 1) {
            uint64_t n2 = jb->b.n;
            for (uint64_t j = 0; j < n2; j++) {
                double f = pair_eval(jb->measure, &jb->a, r, &jb->b, j, &iv);
                if (jb->fout) jb->fout[r * n2 + j] = f;
                if (jb->iout) jb->iout[r * n2 + j] = iv;
                done++;
            }
        } else { /* stream: r = streamed record (b), minor = loaded record (a) */
            uint64_t n1 = jb->a.n;
            for (uint64_t i = 0; i < n1; i++) {
                double f = pair_eval(jb->measure, &jb->a, i, &jb->b, r, &iv);
                if (jb->fout) jb->fout[r * n1 + i] = f;
                if (jb->iout) jb->iout[r * n1 + i] = iv;
                done++;
            }
        }
    }
    jb->done = done;
    return NULL;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct { or_job *jobs; uint64_t nslab; uint64_t next; pthread_mutex_t mu; } or_pool;

static void *pool_worker(void *arg) {
    or_pool *pl = (or_pool *)arg;
    for (;;) {
        pthread_mutex_lock(&pl->mu);
        uint64_t k = pl->next++;
        pthread_mutex_unlock(&pl->mu);
        if (k >= pl->nslab) break;
        job_run(&pl->jobs[k]);
    }
    return NULL;
}

/* Run one mode over all major rows with `threads` pthreads (the reference's
 * -t worker pool, lib.rs:422-458, minus its channels).  Rows are cut into
 * slabs of roughly equal pair count which the workers pull from a shared
 * counter.  Returns pairs computed; *seconds = wall time of the compute loops. */
OR_API uint64_t or_run(int measure, int mode, const or_aln *a, const or_aln *b, double *fout,
                       int64_t *iout, int threads, double *seconds) {
    uint64_t rows = (mode == 0) ? (a->n ? a->n - 1 : 0) : (mode == 1 ? a->n : b->n);
    if (threads < 1) threads = 1;
    uint64_t nslab = (uint64_t)threads * 16;
    if (nslab > rows) nslab = rows ? rows : 1;
    or_job *jobs = (or_job *)calloc(nslab, sizeof(or_job));
    uint64_t r = 0;
    double total_pairs = (mode == 0) ? 0.5 * (double)a->n * (double)(a->n - 1)
                                     : (double)a->n * (double)(b ? b->n : a->n);
    double per = total_pairs / (double)nslab, acc = 0;
    for (uint64_t s = 0; s < nslab; s++) {
        jobs[s].measure = measure; jobs[s].mode = mode;
        jobs[s].a = *a; jobs[s].b = b ? *b : *a;
        jobs[s].fout = fout; jobs[s].iout = iout;
        jobs[s].row0 = r;
        double target = per * (double)(s + 1);
        while (r < rows && acc < target) {
            acc += (mode == 0) ? (double)(a->n - 1 - r)
                               : (mode == 1 ? (double)jobs[s].b.n : (double)a->n);
            r++;
        }
        if (s == nslab - 1) r = rows;
        jobs[s].row1 = r;
    }
    or_pool pool;
    pool.jobs = jobs; pool.nslab = nslab; pool.next = 0;
    pthread_mutex_init(&pool.mu, NULL);
    double t0 = now_s();
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, pool_worker, &pool);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    double t1 = now_s();
    uint64_t done = 0;
    for (uint64_t s = 0; s < nslab; s++) done += jobs[s].done;
    free(th); free(jobs);
    pthread_mutex_destroy(&pool.mu);
    if (seconds) *seconds = t1 - t0;
    return done;
}

/* Bounded CPU-baseline sample: the first `rows` major rows of the mode
 * (square: rows 0..rows-1 against all j>i), split across `threads`.  No
 * output arrays: results are folded into a checksum so the loops cannot be
 * optimised away.  Returns pairs computed. */
typedef struct {
    int measure, mode; or_aln a, b; uint64_t row0, row1, stride; uint64_t done; double sink;
} or_bjob;

static void *bjob_run(void *p) {
    or_bjob *jb = (or_bjob *)p;
    int64_t iv = 0; double acc = 0; uint64_t done = 0;
    for (uint64_t r = jb->row0; r < jb->row1; r += jb->stride) {
        if (jb->mode == 0) {
            for (uint64_t j = r + 1; j < jb->a.n; j++) {
                double f = pair_eval(jb->measure, &jb->a, r, &jb->a, j, &iv);
                acc += (f == f ? f : 0.0) + (double)iv; done++;
            }
        } else if (jb->mode == 1) {
            for (uint64_t j = 0; j < jb->b.n; j++) {
                double f = pair_eval(jb->measure, &jb->a, r, &jb->b, j, &iv);
                acc += (f == f ? f : 0.0) + (double)iv; done++;
            }
        } else {
            for (uint64_t i = 0; i < jb->a.n; i++) {
                double f = pair_eval(jb->measure, &jb->a, i, &jb->b, r, &iv);
                acc += (f == f ? f : 0.0) + (double)iv; done++;
            }
        }
    }
    jb->done = done; jb->sink = acc;
    return NULL;
}

OR_API uint64_t or_bench(int measure, int mode, const or_aln *a, const or_aln *b, uint64_t rows,
                         int threads, double *seconds, double *checksum) {
    if (threads < 1) threads = 1;
    or_bjob *jobs = (or_bjob *)calloc((size_t)threads, sizeof(or_bjob));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t].measure = measure; jobs[t].mode = mode;
        jobs[t].a = *a; jobs[t].b = b ? *b : *a;
        jobs[t].row0 = (uint64_t)t; jobs[t].row1 = rows; jobs[t].stride = (uint64_t)threads;
    }
    double t0 = now_s();
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, bjob_run, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    double t1 = now_s();
    uint64_t done = 0; double cs = 0;
    for (int t = 0; t < threads; t++) { done += jobs[t].done; cs += jobs[t].sink; }
    free(th); free(jobs);
    if (seconds) *seconds = t1 - t0;
    if (checksum) *checksum = cs;
    return done;
}
