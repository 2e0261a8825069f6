// main.cpp -- the `distance` command line on top of libdistance_gpu (C ABI, include/distance_gpu.h).
//
// Same flags, defaults, mutual-exclusion rules, output text and exit codes as the reference CLI
// (lib.rs:68-131 get_cli_arguments, lib.rs:162-267 set_up, lib.rs:490-498 run, main.rs:4-16); the
// pairwise work between "records are parsed" and "ordered results reach the writer" goes to the GPU
// through dg_load_resident / dg_run_square / dg_run_rect / dg_stream_*.  The host is C++ because the
// build image has no Rust toolchain; INTEGRATION.md shows the Rust binding of the same ABI.
//
//   -t  = threads formatting TSV text (the reference's worker count; results never depended on it)
//   -b  = accepted and validated, otherwise unused (the reference's pair batch size)
//   Devices: every visible GPU by default, like the reference takes every core (lib.rs:252-264, num_cpus::get());
//   small inputs take fewer (one GPU per 2e13 pair-sites, ~0.2 s of tensor work: a CUDA context costs more than that).
//   DISTANCE_GPUS=<k> or DISTANCE_GPUS=0,2,3 restricts / chooses them; there is no flag for it so the CLI surface stays
//   the reference's.
#include <cerrno>
#include <chrono>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../../include/distance_gpu.h"
#include "fasta.hpp"
#include "tsv.hpp"

using namespace host;

namespace {

const char* kVersion = "0.3.1";  // Cargo.toml:3
const char* kMeasures[] = {"n", "n_high", "raw", "jc69", "k80", "tn93"};

const char* kHelp =
    "Calculate genetic distances within/between fasta-format alignments of DNA sequences\n"
    "\n"
    "Usage: All sequences across all input files must be the same length.\n"
    "\n"
    "       distance alignment.fasta\n"
    "       cat alignment.fasta | distance\n"
    "       distance alignment.fasta -o distances.tsv\n"
    "       distance -t 8 -m jc69 alignment.fasta -o jc69.tsv\n"
    "       distance alignment1.fasta alignment2.fasta > distances2.tsv\n"
    "       distance -i smallAlignment.fasta -s bigAlignment.fasta -o distances3.tsv\n"
    "       cat bigAlignment.fasta | distance smallAlignment.fasta -s - > distances3.tsv\n"
    "\n"
    "Options:\n"
    "  -i, --input [<input>...]     One or two input alignment files in fasta format. Loaded into memory. This flag can be omitted and the files passed as positional arguments\n"
    "  -s, --stream <stream>        One input alignment file in fasta format. Streamed from disk (or stdin using \"-s -\"). Requires exactly one file also be loaded\n"
    "  -m, --measure <measure>      Which distance measure to use [default: raw] [possible values: n, n_high, raw, jc69, k80, tn93]\n"
    "  -o, --output <output>        Output file in tab-separated-value format. Omit this option to print to stdout\n"
    "  -t, --threads <threads>      How many threads to spin up for pairwise comparisons. Omitting this option spins up the number of available CPUs\n"
    "  -b, --batchsize <batchsize>  Try setting this >(>) 1 to tune the workload per thread [default: 1]\n"
    "  -l, --licenses               Print licence information and exit\n"
    "  -h, --help                   Print help\n"
    "  -V, --version                Print version\n";

const char* kLicences =
    "\ndistance (B200 host) follows the reference program's licensing: the reference `distance` is\n"
    "Copyright 2022 Ben Jackson, GNU LIBRARY GENERAL PUBLIC LICENSE Version 2; its FASTA reading\n"
    "derives from Rust-Bio (MIT licence, Copyright (c) 2016 Johannes Koester, the Rust-Bio team,\n"
    "Google Inc.).  See the LICENSE file of the reference distribution for the full texts.\n";

struct Args {
    std::vector<std::string> flag_inputs, pos_inputs;
    bool have_stream = false;
    std::string stream;
    std::string measure = "raw";
    bool have_output = false;
    std::string output;
    bool have_threads = false;
    uint64_t threads = 0;
    uint64_t batchsize = 1;
    bool licenses = false;
};

[[noreturn]] void clap_error(const std::string& msg) {  // clap prints to stderr and exits 2
    fprintf(stderr, "error: %s\n\nFor more information, try '--help'.\n", msg.c_str());
    std::exit(2);
}

bool parse_usize(const std::string& s, uint64_t* out) {
    if (s.empty()) return false;
    size_t i = 0;
    if (s[0] == '+') i = 1;
    if (i >= s.size()) return false;
    uint64_t v = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        const uint64_t nv = v * 10 + (uint64_t)(s[i] - '0');
        if (nv < v) return false;
        v = nv;
    }
    *out = v;
    return true;
}

Args parse_args(int argc, char** argv) {
    Args a;
    bool only_positional = false;
    for (int k = 1; k < argc; k++) {
        std::string t = argv[k];
        auto value = [&](const std::string& name, bool inline_ok, const std::string& inline_val) -> std::string {
            if (inline_ok) return inline_val;
            if (k + 1 >= argc) clap_error("a value is required for '" + name + "' but none was supplied");
            return argv[++k];
        };
        if (only_positional || t == "-" || t.empty() || t[0] != '-') {
            if (a.pos_inputs.size() >= 2) clap_error("unexpected argument '" + t + "' found");
            a.pos_inputs.push_back(t);
            continue;
        }
        if (t == "--") { only_positional = true; continue; }
        std::string key = t, inl;
        bool has_inl = false;
        if (t.rfind("--", 0) == 0) {
            const size_t eq = t.find('=');
            if (eq != std::string::npos) { key = t.substr(0, eq); inl = t.substr(eq + 1); has_inl = true; }
        } else if (t.size() > 2) {  // -m=jc69 / -mjc69
            key = t.substr(0, 2);
            inl = t.substr(t[2] == '=' ? 3 : 2);
            has_inl = true;
        }
        if (key == "-h" || key == "--help") { fputs(kHelp, stdout); std::exit(0); }
        if (key == "-V" || key == "--version") { printf("distance %s\n", kVersion); std::exit(0); }
        if (key == "-l" || key == "--licenses") { a.licenses = true; continue; }
        if (key == "-i" || key == "--input") {  // num_args(0..=2), lib.rs:89
            if (has_inl) a.flag_inputs.push_back(inl);
            while (a.flag_inputs.size() < 2 && k + 1 < argc && (argv[k + 1][0] != '-' || !strcmp(argv[k + 1], "-")))
                a.flag_inputs.push_back(argv[++k]);
            continue;
        }
        if (key == "-s" || key == "--stream") { a.stream = value("--stream <stream>", has_inl, inl); a.have_stream = true; continue; }
        if (key == "-m" || key == "--measure") {
            a.measure = value("--measure <measure>", has_inl, inl);
            bool ok = false;
            for (const char* m : kMeasures) ok |= a.measure == m;
            if (!ok)
                clap_error("invalid value '" + a.measure + "' for '--measure <measure>'\n  [possible values: n, n_high, raw, jc69, k80, tn93]");
            continue;
        }
        if (key == "-o" || key == "--output") { a.output = value("--output <output>", has_inl, inl); a.have_output = true; continue; }
        if (key == "-t" || key == "--threads") {
            const std::string v = value("--threads <threads>", has_inl, inl);
            if (!parse_usize(v, &a.threads)) clap_error("invalid value '" + v + "' for '--threads <threads>': invalid digit found in string");
            a.have_threads = true;
            continue;
        }
        if (key == "-b" || key == "--batchsize") {
            const std::string v = value("--batchsize <batchsize>", has_inl, inl);
            if (!parse_usize(v, &a.batchsize)) clap_error("invalid value '" + v + "' for '--batchsize <batchsize>': invalid digit found in string");
            continue;
        }
        clap_error("unexpected argument '" + t + "' found");
    }
    return a;
}

int open_read(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw io_error_os(errno);  // File::open(path)? (lib.rs:193, 206)
    return fd;
}

int measure_id(const std::string& m) {
    for (int i = 0; i < 6; i++)
        if (m == kMeasures[i]) return i;
    return -1;
}

// GPUs the run can use, from the device files of the driver (no CUDA call: the point is to decide BEFORE the driver starts)
int count_gpu_device_files() {
    int n = 0;
    char path[64];
    for (int i = 0; i < 64; i++) {
        snprintf(path, sizeof path, "/dev/nvidia%d", i);
        if (access(path, F_OK) == 0) n++;
    }
    return n;
}
// A run that needs k of the box's GPUs hides the others from the driver before its first CUDA call: on an 8-GPU NVSwitch
// box the start-up of the driver costs 8 - 10 s with every device visible (measured: dg_create done at 9.7 s for a 2-GPU
// run) against ~1 s on a single-GPU box.  Respects an existing CUDA_VISIBLE_DEVICES / DISTANCE_GPUS.
void limit_visible_devices(double pair_sites) {
    if (std::getenv("CUDA_VISIBLE_DEVICES") || std::getenv("DISTANCE_GPUS") || pair_sites <= 0) return;
    const int have = count_gpu_device_files();
    const int want = (int)std::max(1.0, std::ceil(pair_sites / 2e13));
    if (have <= 1 || want >= have) return;
    std::string v;
    for (int i = 0; i < want; i++) v += (i ? "," : "") + std::to_string(i);
    setenv("CUDA_VISIBLE_DEVICES", v.c_str(), 1);
}

// pair_sites: the work of the run (0 = unknown, e.g. a stream of unknown length: every visible device)
std::vector<int> gpu_list(double pair_sites) {
    std::vector<int> ids;
    const char* e = std::getenv("DISTANCE_GPUS");
    if (!e || !*e) {
        const int visible = std::max(1, dg_device_count());
        int k = visible;
        if (pair_sites > 0) k = (int)std::min<double>(visible, std::max(1.0, std::ceil(pair_sites / 2e13)));
        for (int i = 0; i < k; i++) ids.push_back(i);
        return ids;
    }
    std::string s = e;
    if (s.find(',') == std::string::npos) {
        const int k = std::atoi(s.c_str());
        for (int i = 0; i < std::max(1, k); i++) ids.push_back(i);
        return ids;
    }
    size_t p = 0;
    while (p <= s.size()) {
        const size_t q = s.find(',', p);
        const std::string tok = s.substr(p, q == std::string::npos ? std::string::npos : q - p);
        if (!tok.empty()) ids.push_back(std::atoi(tok.c_str()));
        if (q == std::string::npos) break;
        p = q + 1;
    }
    return ids.empty() ? std::vector<int>{0} : ids;
}

struct SinkState {
    TsvWriter* w;
    std::string error;
};

int sink_cb(void* user, const dg_panel* p) {
    SinkState* st = static_cast<SinkState*>(user);
    try {
        st->w->write_panel(*p);
        return 0;
    } catch (const DistanceError& e) {
        st->error = e.what();
        return 1;
    }
}

void gpu_check(dg_ctx* ctx, int rc, SinkState* st = nullptr) {
    if (rc == DG_OK) return;
    if (rc == DG_ERR_SINK && st && !st->error.empty()) throw DistanceError(st->error);
    throw message_error(std::string("GPU engine: ") + dg_last_error(ctx));
}

int run(const Args& a) {
    // ---- set_up (lib.rs:162-267), same order of checks -------------------------------------------
    if (!a.pos_inputs.empty() && !a.flag_inputs.empty())  // lib.rs:182-184
        throw message_error("For loading input files, don't use both positional arguments and the -i/--input flag");
    std::vector<std::string> inputs = a.flag_inputs;
    inputs.insert(inputs.end(), a.pos_inputs.begin(), a.pos_inputs.end());
    std::vector<int> fds;
    if (inputs.empty()) fds.push_back(0);              // stdin, lib.rs:189-191
    for (const auto& p : inputs) fds.push_back(open_read(p));
    int stream_fd = -1;
    if (a.have_stream) {
        if (inputs.size() != 1)  // lib.rs:196-199
            throw message_error("If you stream one file, you must also provide exactly one other file to be loaded");
        stream_fd = a.stream == "-" ? 0 : open_read(a.stream);  // lib.rs:200-207
    }
    const bool trace = std::getenv("DG_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    // -t (lib.rs:252-264): here the worker count of the host side -- FASTA parsing and TSV formatting
    const uint64_t threads = a.have_threads ? std::max<uint64_t>(1, a.threads)
                                            : std::max(1u, std::thread::hardware_concurrency());
    std::vector<Alignment> loaded;
    for (size_t k = 0; k < fds.size(); k++) {
        loaded.push_back(load_fasta(fds[k], (int)std::min<uint64_t>(threads, 64)));
        if (k == 1) check_same_width(loaded[0], loaded[1]);
        if (fds[k] != 0) ::close(fds[k]);
    }
    int out_fd = 1;
    if (a.have_output) {
        out_fd = ::open(a.output.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);  // File::create, lib.rs:248-250
        if (out_fd < 0) throw io_error_os(errno);
    }

    // ---- run (lib.rs:490-498) ----------------------------------------------------------------------
    if (trace) fprintf(stderr, "[distance] %.3f s: inputs parsed\n", since());
    const uint64_t width = loaded[0].width;
    TsvWriter writer(out_fd, (int)std::min<uint64_t>(threads, 256));
    SinkState st{&writer, {}};
    double pair_sites = 0;   // stream mode: unknown length -> every device
    if (stream_fd < 0) {
        const double n0 = (double)loaded[0].n(), n1 = loaded.size() > 1 ? (double)loaded[1].n() : 0;
        pair_sites = (loaded.size() > 1 ? n0 * n1 : n0 * (n0 - 1) / 2) * (double)width;
    }
    limit_visible_devices(pair_sites);
    const std::vector<int> gpus = gpu_list(pair_sites);
    if (trace) fprintf(stderr, "[distance] %zu GPU(s)\n", gpus.size());
    dg_ctx* ctx = nullptr;
    int rc = dg_create(gpus.data(), (int)gpus.size(), measure_id(a.measure), width, &ctx);
    if (rc != DG_OK) throw message_error(std::string("GPU engine: ") + dg_last_error(nullptr));
    struct Guard { dg_ctx* c; ~Guard() { if (c) dg_destroy(c); } } guard{ctx};
    if (trace) fprintf(stderr, "[distance] %.3f s: dg_create done (CUDA context, streams)\n", since());
    if (width <= 65535) dg_set_option(ctx, DG_OPT_RESULT_U16, 1);  // n / n_high: half the D2H bytes, same text
    // raw / jc69 / k80 / tn93 over loaded files: the panels carry the integer counts and the writer's threads evaluate the
    // f64 expressions of measures.rs with this host's libm, so the text is byte-identical to the reference's (CUDA's log
    // may be an ulp off glibc's).  Streamed runs keep the device's fused f64 epilogue.  DISTANCE_DEVICE_F64=1 switches it off.
    const int mid = measure_id(a.measure);
    const bool host_f64 = mid >= 2 && stream_fd < 0 && width <= 65535 && !std::getenv("DISTANCE_DEVICE_F64");
    std::vector<uint32_t> acgt[2];
    if (host_f64) {
        dg_set_option(ctx, DG_OPT_RESULT_COUNTS, 1);
        if (mid == DG_MEASURE_TN93)   // count_bases (fastaio.rs:53-66): the codes of A / T / G / C, either case
            for (size_t k = 0; k < loaded.size(); k++) {
                acgt[k].assign(loaded[k].n() * 4, 0);
                const uint8_t* base = loaded[k].data();
                const uint64_t nrec = loaded[k].n();
                const int T = (int)std::min<uint64_t>(threads, 64);
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++)
                    th.emplace_back([&, t] {
                        for (uint64_t r = t; r < nrec; r += T) {
                            const uint8_t* s = base + r * width;
                            uint32_t cA = 0, cT = 0, cG = 0, cC = 0;
                            for (uint64_t i = 0; i < width; i++) {
                                const uint8_t ch = s[i] | 0x20;   // lower-case the letter
                                cA += ch == 'a'; cT += ch == 't'; cG += ch == 'g'; cC += ch == 'c';
                            }
                            uint32_t* o = acgt[k].data() + 4 * r;
                            o[0] = cA; o[1] = cT; o[2] = cG; o[3] = cC;
                        }
                    });
                for (auto& x : th) x.join();
            }
        writer.set_counts_measure(mid, acgt[0].empty() ? nullptr : acgt[0].data(),
                                  loaded.size() > 1 ? (acgt[1].empty() ? nullptr : acgt[1].data()) : (acgt[0].empty() ? nullptr : acgt[0].data()));
    }
    // 32 MiB result panels: the pinned ring (two panels per GPU) costs ~0.5 ms per MB to page-lock, and the writer
    // formats a panel far faster than the GPU produces it, so larger panels buy nothing here
    dg_set_option(ctx, DG_OPT_PANEL_BYTES, 32 << 20);

    for (size_t k = 0; k < loaded.size(); k++)
        gpu_check(ctx, dg_load_resident(ctx, (int)k, loaded[k].data(), loaded[k].n(), DG_INPUT_ASCII, nullptr));

    if (trace) fprintf(stderr, "[distance] %.3f s: GPU context up, alignments resident\n", since());
    writer.write_header();  // gather_write writes the header before anything arrives (lib.rs:613)
    if (stream_fd >= 0) {
        // stream() (lib.rs:269-365) + stream_fasta (fastaio.rs:215-286)
        std::vector<std::string> streamed_ids;
        writer.set_ids(&loaded[0].ids, &streamed_ids);
        const uint64_t batch = std::max<uint64_t>(64, std::min<uint64_t>(8192, (64ull << 20) / std::max<uint64_t>(1, width)));
        gpu_check(ctx, dg_stream_begin(ctx, sink_cb, &st, batch), &st);
        // whole-record blocks parsed in parallel straight into the library's pinned staging buffer (no host copy)
        StreamBlockParser parser(stream_fd, width, (int)std::min<uint64_t>(threads, 64));
        uint64_t records = 0;
        for (;;) {
            uint8_t* buf = nullptr;
            uint64_t cap = 0;
            gpu_check(ctx, dg_stream_buffer(ctx, &buf, &cap), &st);
            const uint64_t got = parser.next(buf, cap, streamed_ids);
            if (got == 0) break;
            records += got;
            gpu_check(ctx, dg_stream_push(ctx, buf, got, DG_INPUT_ASCII, nullptr), &st);
        }
        gpu_check(ctx, dg_stream_end(ctx), &st);
        if (records == 0) throw message_error("Empty FASTA file");  // fastaio.rs:281-283 (after the header was written)
    } else if (loaded.size() == 1) {
        writer.set_ids(&loaded[0].ids, &loaded[0].ids);
        gpu_check(ctx, dg_run_square(ctx, sink_cb, &st, 0), &st);
    } else {
        writer.set_ids(&loaded[0].ids, &loaded[1].ids);
        gpu_check(ctx, dg_run_rect(ctx, sink_cb, &st, 0), &st);
    }
    writer.flush();
    if (out_fd != 1) ::close(out_fd);
    if (trace) fprintf(stderr, "[distance] %.3f s: %llu lines written\n", since(), (unsigned long long)writer.lines());
    // Everything is written: leave without tearing the CUDA context and the page-locked buffers down one by one
    // (~0.3 s); the OS reclaims them with the process.
    fflush(stderr);
    std::_Exit(0);
}

}  // namespace

// Hidden developer check (not part of the reference CLI): the exact `{:.12}` formatter against the C
// library's exact %.12f on N pseudo-random doubles plus edge cases.  `distance --selftest-format N`.
int selftest_format(uint64_t n) {
    uint64_t x = 0x9E3779B97F4A7C15ull, bad = 0;
    auto check = [&](double d) {
        std::string got;
        format_float12(d, got);
        char want[512];
        if (d != d) snprintf(want, sizeof want, "NaN");
        else if (d == 1.0 / 0.0) snprintf(want, sizeof want, "inf");
        else if (d == -1.0 / 0.0) snprintf(want, sizeof want, "-inf");
        else snprintf(want, sizeof want, "%.12f", d);
        if (got != want) { if (bad++ < 10) fprintf(stderr, "MISMATCH %a: got %s want %s\n", d, got.c_str(), want); }
    };
    const double edge[] = {0.0, -0.0, 0.5e-12, 1.5e-12, 2.5e-12, 0.9999999999995, 0.1333333333333333, 1.0, 123456.789,
                           1e-300, 4.9e-324, 1e22, 2147483648.0, 0.0000000000005, 0.00000000000049999, 1.0 / 0.0, -1.0 / 0.0, 0.0 / 0.0};
    for (double d : edge) { check(d); check(-d); }
    for (uint64_t i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        double d;
        if (i % 3 == 0) {  // raw bit patterns: every exponent
            uint64_t b = x; std::memcpy(&d, &b, 8);
        } else if (i % 3 == 1) {  // distances: [0, 4)
            d = (double)(x >> 11) / 9007199254740992.0 * 4.0;
        } else {  // ties at the 12th place: k / 2^j
            d = (double)(x >> 40) / (double)(1ull << (20 + (x & 31)));
        }
        check(d);
    }
    printf("selftest-format: %llu values, %llu mismatches\n", (unsigned long long)(n + 36), (unsigned long long)bad);
    return bad ? 1 : 0;
}

// Hidden developer check: the parallel loader against the sequential reader on the same file (ids, width and every
// byte must agree; errors must be the same text), with the time each takes.  `distance --selftest-parse FILE [threads]`.
int selftest_parse(const char* path, int threads) {
    auto load = [&](bool parallel, Alignment& a, std::string& err, double& secs) {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) { err = "open failed"; return; }
        const auto t0 = std::chrono::steady_clock::now();
        try {
            if (parallel) {
                a = load_fasta(fd, threads);
            } else {
                FastaReader rd(fd);
                std::string id;
                bool first = true;
                for (;;) {
                    const size_t before = a.seqs.size();
                    if (!rd.next(id, a.seqs)) break;
                    const uint64_t len = a.seqs.size() - before;
                    if (first) { a.width = len; first = false; }
                    else if (len != a.width)
                        throw message_error("Different length sequences in alignment(s): " + std::to_string(len) + " vs " + std::to_string(a.width));
                    a.ids.push_back(id);
                }
                if (a.ids.empty()) throw message_error("Empty FASTA file");
            }
        } catch (const DistanceError& e) {
            err = e.what();
        }
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ::close(fd);
    };
    Alignment fast, slow;
    std::string ef, es;
    double tf = 0, ts = 0;
    load(true, fast, ef, tf);
    load(false, slow, es, ts);
    bool same = ef == es;
    if (same && ef.empty()) {
        same = fast.ids == slow.ids && fast.width == slow.width &&
               std::memcmp(fast.data(), slow.data(), (size_t)(fast.n() * fast.width)) == 0;
    }
    printf("selftest-parse: %s; records %llu width %llu; parallel loader %.3f s (%s), sequential reader %.3f s; error '%s'\n",
           same ? "identical" : "MISMATCH", (unsigned long long)slow.n(), (unsigned long long)slow.width, tf,
           fast.fast_ ? "parallel pass" : "fell back to the sequential reader", ts, es.c_str());
    if (std::getenv("DG_SELFTEST_DUMP"))   // tests: what the reader made of the file, one "id<TAB>sequence" line per record
        for (uint64_t r = 0; r < slow.n(); r++)
            printf("%s\t%.*s\n", slow.ids[r].c_str(), (int)slow.width, (const char*)slow.data() + r * slow.width);
    return same ? 0 : 1;
}

// Hidden developer check: the block parser of the streamed file against the sequential reader with stream_fasta's checks
// (same ids, bytes, record count, error text).  `distance --selftest-stream FILE WIDTH BATCH [threads]`.
int selftest_stream(const char* path, uint64_t width, uint64_t batch, int threads) {
    std::vector<std::string> ids_b, ids_s;
    std::vector<uint8_t> seq_b, seq_s;
    std::string eb, es;
    {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) { perror("open"); return 1; }
        try {
            StreamBlockParser parser(fd, width, threads);
            std::vector<uint8_t> buf(batch * width);
            for (;;) {
                const uint64_t got = parser.next(buf.data(), batch, ids_b);
                if (got == 0) break;
                seq_b.insert(seq_b.end(), buf.begin(), buf.begin() + (std::ptrdiff_t)(got * width));
            }
        } catch (const DistanceError& e) { eb = e.what(); }
        ::close(fd);
    }
    {
        const int fd = ::open(path, O_RDONLY);
        try {
            FastaReader rd(fd, false);
            std::string id;
            for (;;) {
                const size_t before = seq_s.size();
                if (!rd.next(id, seq_s)) break;
                const uint64_t len = seq_s.size() - before;
                if (len != width)
                    throw message_error("Different length sequences in alignment(s): " + std::to_string(len) + " vs " + std::to_string(width));
                validate_record(id, seq_s.data() + before, len);
                ids_s.push_back(id);
            }
        } catch (const DistanceError& e) { es = e.what(); }
        ::close(fd);
    }
    // on an error the records delivered before it must be a prefix of the sequential reader's
    bool same = eb == es;
    if (es.empty()) same = same && ids_b == ids_s && seq_b == seq_s;
    else same = same && ids_b.size() <= ids_s.size() && std::equal(ids_b.begin(), ids_b.end(), ids_s.begin()) &&
                std::equal(seq_b.begin(), seq_b.end(), seq_s.begin());
    printf("selftest-stream: %s; records %zu (block parser %zu); error '%s'\n", same ? "identical" : "MISMATCH", ids_s.size(),
           ids_b.size(), es.c_str());
    return same ? 0 : 1;
}

// Hidden developer check: the pooled TSV writer against a line-by-line restatement of gather_write (lib.rs:626-633)
// for every mode and result kind, with short and long ids, several thread counts, panels larger than one chunk.
// `distance --selftest-tsv`.
int selftest_tsv() {
    uint64_t x = 0x243F6A8885A308D3ull;
    auto rnd = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    int bad = 0, cases = 0;
    for (int threads : {1, 3, 8}) {
        for (int mode : {DG_MODE_SQUARE, DG_MODE_RECT, DG_MODE_STREAM}) {
            for (int kind : {DG_RESULT_U16, DG_RESULT_U32, DG_RESULT_F64}) {
                const uint64_t n1 = 700 + rnd() % 200, n2 = mode == DG_MODE_SQUARE ? n1 : 300 + rnd() % 100;
                std::vector<std::string> ids1, ids2;
                auto make_id = [&](uint64_t i, const char* pre) {
                    std::string s = pre + std::to_string(i);
                    if (rnd() % 7 == 0) s += "|EPI_ISL_" + std::to_string(rnd() % 100000000) + "|a/very/long/identifier/2020-03-01";
                    return s;
                };
                for (uint64_t i = 0; i < n1; i++) ids1.push_back(make_id(i, "s"));
                for (uint64_t i = 0; i < n2; i++) ids2.push_back(make_id(i, "q"));
                char path[] = "/tmp/dg_tsv_selftest_XXXXXX";
                const int fd = mkstemp(path);
                if (fd < 0) { perror("mkstemp"); return 1; }
                // every other case through an O_APPEND descriptor: the writer then keeps ordered write() calls (as for a
                // pipe) instead of positioned parallel writes
                if ((cases & 1) && fcntl(fd, F_SETFL, fcntl(fd, F_GETFL) | O_APPEND) < 0) { perror("fcntl"); return 1; }
                std::string want = "sequence1\tsequence2\tdistance\n";
                {
                    TsvWriter w(fd, threads);
                    w.set_ids(&ids1, mode == DG_MODE_SQUARE ? &ids1 : &ids2);
                    w.write_header();
                    // RECT: rows = ids1, cols = ids2.  STREAM: rows = streamed ids (ids2), cols = loaded ids (ids1).
                    const uint64_t rows_total = mode == DG_MODE_SQUARE ? n1 - 1 : (mode == DG_MODE_RECT ? n1 : n2);
                    const uint64_t cols = mode == DG_MODE_SQUARE ? n1 : (mode == DG_MODE_RECT ? n2 : n1);
                    for (uint64_t r0 = 0; r0 < rows_total;) {
                        const uint64_t r1 = std::min<uint64_t>(rows_total, r0 + 150 + rnd() % 300);
                        uint64_t cnt = 0;
                        for (uint64_t r = r0; r < r1; r++) cnt += mode == DG_MODE_SQUARE ? n1 - 1 - r : cols;
                        std::vector<uint16_t> d16(cnt);
                        std::vector<uint32_t> d32(cnt);
                        std::vector<double> d64(cnt);
                        for (uint64_t k = 0; k < cnt; k++) {
                            const uint64_t v = rnd();
                            d16[k] = (uint16_t)(v % 5 == 0 ? v >> 20 : v % 120);
                            d32[k] = (uint32_t)(v % 11 == 0 ? v >> 33 : v % 70000);
                            switch (v % 9) {
                            case 0: d64[k] = 0.0; break;
                            case 1: d64[k] = -0.0; break;
                            case 2: d64[k] = std::nan(""); break;
                            case 3: d64[k] = (v & 1024) ? INFINITY : -INFINITY; break;
                            case 4: d64[k] = (double)(v >> 11) * 1e3; break;
                            default: d64[k] = (double)(v >> 11) / 9007199254740992.0 * 2.0;
                            }
                        }
                        dg_panel p{};
                        p.mode = mode; p.result_kind = kind; p.row_begin = r0; p.row_end = r1; p.n_cols = cols; p.n_results = cnt;
                        p.data = kind == DG_RESULT_U16 ? (const void*)d16.data() : (kind == DG_RESULT_U32 ? (const void*)d32.data() : (const void*)d64.data());
                        w.write_panel(p);
                        uint64_t k = 0;
                        for (uint64_t r = r0; r < r1; r++)
                            for (uint64_t c = mode == DG_MODE_SQUARE ? r + 1 : 0; c < cols; c++, k++) {
                                if (mode == DG_MODE_SQUARE) want += ids1[r] + "\t" + ids1[c] + "\t";
                                else if (mode == DG_MODE_RECT) want += ids1[r] + "\t" + ids2[c] + "\t";
                                else want += ids1[c] + "\t" + ids2[r] + "\t";
                                if (kind == DG_RESULT_U16) format_u32(d16[k], want);
                                else if (kind == DG_RESULT_U32) format_u32(d32[k], want);
                                else format_float12(d64[k], want);
                                want += "\n";
                            }
                        r0 = r1;
                    }
                }
                std::string got(want.size() + 16, '\0');
                const ssize_t n = ::pread(fd, &got[0], got.size(), 0);
                got.resize(n > 0 ? (size_t)n : 0);
                ::close(fd);
                ::unlink(path);
                cases++;
                if (got != want) {
                    bad++;
                    fprintf(stderr, "MISMATCH threads %d mode %d kind %d: %zu vs %zu bytes\n", threads, mode, kind, got.size(), want.size());
                }
            }
        }
    }
    printf("selftest-tsv: %d cases, %d mismatches\n", cases, bad);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    std::signal(SIGPIPE, SIG_IGN);
    if (argc == 2 && !strcmp(argv[1], "--selftest-tsv")) return selftest_tsv();
    if (argc >= 5 && !strcmp(argv[1], "--selftest-stream"))
        return selftest_stream(argv[2], strtoull(argv[3], nullptr, 10), strtoull(argv[4], nullptr, 10), argc > 5 ? atoi(argv[5]) : 0);  // a closed pipe shows up as EPIPE -> exit 0 (lib.rs:598-608)
    if (argc == 3 && !strcmp(argv[1], "--selftest-format")) return selftest_format(strtoull(argv[2], nullptr, 10));
    if (argc >= 3 && !strcmp(argv[1], "--selftest-parse")) return selftest_parse(argv[2], argc > 3 ? atoi(argv[3]) : 0);
    const Args a = parse_args(argc, argv);
    if (a.licenses) {  // main.rs:7-10
        printf("%s\n", kLicences);
        return 0;
    }
    try {
        return run(a);
    } catch (const DistanceError& e) {
        fprintf(stderr, "Error: %s\n", e.what());  // Debug form of the returned Err (main.rs:4-15)
        return 1;
    }
}
