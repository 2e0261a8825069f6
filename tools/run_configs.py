#!/usr/bin/env python
"""BASELINE.json configs 1-5 through the C ABI on ONE GPU (one JSON line per config).

  cfg1  raw   all-vs-all, 1,000 records                      (also through the `distance` CLI: FASTA in, TSV out)
  cfg2  n and n_high all-vs-all, 20,000 records              (bench.py's workload; here for the `n` measure too)
  cfg3  tn93  10,000 x 10,000 between two alignments
  cfg4  k80   1,000 resident vs 1,000,000 streamed           (a pinned pool of --pool records, cycled)
  cfg5  jc69  100,000 all-vs-all: THIS GPU's 1/8 share       (dg_run_part(part 0 of 8): what one of 8 B200s does)

Per config: kernel-only device time (CUDA events on the library's streams, operands resident, re-pack
included), end to end through the C ABI from pinned host buffers (H2D + pack + tiles + D2H + sink), the
tensor roofline fraction of the count kernel, and the CPU oracle on a bounded sample of the same inputs.
usage: run_configs.py [--only 1,3] [--cpu-budget S] [--stream-records N]"""
import argparse, ctypes as C, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import distance_b200 as dg
from distance_b200 import api, synth

W = synth.SC2_WIDTH
I8_OPS = {"n": 8, "n_high": 8, "raw": 16, "jc69": 16, "k80": 12, "tn93": 10}
PEAKS = json.load(open(os.path.join(ROOT, "profiles", "int_peaks.json")))
I8_PEAK = PEAKS["int8_tops_measured"] * 1e12
FP4_PEAK = PEAKS["fp4_tops_measured"] * 1e12


def pinned(a):
    p = api.pinned_array(a.shape, np.uint8)
    p[...] = a
    return p


def cpu_sample(measure, mode, a_codes, b_codes, budget):
    from oracle import oracle as orc
    a = orc.Alignment(a_codes)
    b = None if b_codes is None else orc.Alignment(b_codes)
    orc.prepare(measure, [a] + ([b] if b is not None else []), consensus_from=[a] if mode == "stream" else None)
    thr = os.cpu_count() or 1
    pairs, secs = orc.bench(measure, mode, a, b, thr, thr)
    per_row = pairs / thr
    total_rows = (a.n - 1) if mode == "square" else (a.n if mode == "rect" else b.n)
    rows = int(min(total_rows, max(thr, budget * (pairs / secs) / per_row))) // thr * thr or thr
    pairs, secs = orc.bench(measure, mode, a, b, rows, thr)
    return {"value": pairs / secs, "unit": "pairs/s", "cores": thr, "kind": "port",
            "sample": f"{rows} major rows ({pairs} pairs, {secs:.1f} s), oracle C port, {thr} pthreads"}


def line(cfg, measure, pairs, dev_ms, e2e_ms, extra):
    out = {"config": cfg, "measure": measure, "pairs": pairs, "width": W,
           "kernel_only": {"ms": dev_ms, "pairs_per_s": pairs / dev_ms * 1e3, "pair_sites_per_s": pairs * W / dev_ms * 1e3,
                           "int8_tops": pairs * W * I8_OPS[measure] / dev_ms * 1e3 / 1e12,
                           "frac_of_measured_int8_peak": pairs * W * I8_OPS[measure] / dev_ms * 1e3 / I8_PEAK,
                           "frac_of_measured_fp4_peak": pairs * W * I8_OPS[measure] / dev_ms * 1e3 / FP4_PEAK},
           "e2e": {"ms": e2e_ms, "pairs_per_s": pairs / e2e_ms * 1e3}}
    out.update(extra)
    print(json.dumps(out), flush=True)


def square_or_rect(cfg, measure, a, b, args, part=0, n_parts=1, u16=False):
    mode = api.DG_MODE_SQUARE if b is None else api.DG_MODE_RECT
    e = dg.Engine(measure, W)
    e.set_option(api.DG_OPT_KEEP_CODES, 1)
    if u16:
        e.set_option(api.DG_OPT_RESULT_U16, 1)
    pa = pinned(a)
    pb = None if b is None else pinned(b)
    e.load(0, pa)
    if pb is not None:
        e.load(1, pb)
    elem = (2 if u16 else 4) if measure in ("n", "n_high") else 8
    if elem == 2:
        e.set_option(api.DG_OPT_PANEL_BYTES, 128 << 20)
    plan = e.plan(mode)
    from distance_b200 import dist
    pairs = sum(p[2] for p in dist.my_panels(plan, part, n_parts))
    for _ in range(2):
        e.run_device_only(mode, part, n_parts, repack=True)
    ms = []
    for _ in range(args.steps):
        e.run_device_only(mode, part, n_parts, repack=True)
        ms.append(e.timings()["run_ms"])
    eng = e.timings()["engine"]
    t_e2e = []
    for it in range(args.steps + 1):
        t0 = time.time()
        e.load(0, pa)
        if pb is not None:
            e.load(1, pb)
        got = e.run_discard(mode, part, n_parts)
        if it:
            t_e2e.append(1e3 * (time.time() - t0))
        assert got == pairs
    # the pipelined sessions (dg_run_square_host / dg_run_rect_host): upload, tiles and D2H overlap
    t_pipe = []
    state = {"n": 0}

    def sink(user, pp):
        state["n"] += int(pp.contents.n_results)
        return 0

    cb = api.SINK_FN(sink)
    for it in range(args.steps + 1):
        state["n"] = 0
        t0 = time.time()
        if pb is None:
            e._check(e.L.dg_run_square_host(e.h, C.c_void_p(pa.ctypes.data), a.shape[0], 0, None, part, n_parts, cb, None))
        else:
            e.load(1, pb)
            e._check(e.L.dg_run_rect_host(e.h, C.c_void_p(pa.ctypes.data), a.shape[0], 0, None, part, n_parts, cb, None))
        if it:
            t_pipe.append(1e3 * (time.time() - t0))
    e.close()
    cpu = cpu_sample(measure, "square" if b is None else "rect", a, b, args.cpu_budget)
    line(cfg, measure, pairs, float(np.median(ms)), float(np.median(t_e2e)),
         {"e2e_pipelined": {"ms": float(np.median(t_pipe)), "pairs_per_s": state["n"] / float(np.median(t_pipe)) * 1e3,
                            "note": "dg_run_square_host / dg_run_rect_host; e2e above = dg_load_resident + dg_run_* (in order)"},
          "engine": int(eng), "h2d_bytes": int(a.nbytes + (0 if b is None else b.nbytes)), "d2h_bytes": pairs * elem,
          "part": f"{part}/{n_parts}", "cpu_baseline": cpu})


def cfg1(args):
    a = synth.encode_ascii(synth.make_alignment(1000, seed=20251018 + 1, ambiguity=True))
    square_or_rect("cfg1: raw all-vs-all 1,000 x 29,903", "raw", a, None, args)
    # the CLI: FASTA parse + GPU + exact {:.12} TSV text
    cli = os.path.join(ROOT, "distance_b200", "_bin", "distance")
    if os.path.exists(cli):
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            fa, out = os.path.join(td, "a.fasta"), os.path.join(td, "out.tsv")
            synth.write_fasta(fa, synth.make_alignment(1000, seed=20251018 + 1, ambiguity=True), synth.ids(1000))
            best = 1e9
            for _ in range(3):
                t0 = time.time()
                subprocess.check_call([cli, "-m", "raw", "-o", out, fa])
                best = min(best, time.time() - t0)
            print(json.dumps({"config": "cfg1 through the `distance` CLI (process start, CUDA context, FASTA parse, GPU, TSV write)",
                              "pairs": 499500, "wall_s": best, "pairs_per_s": 499500 / best, "tsv_bytes": os.path.getsize(out)}), flush=True)


def cfg2(args):
    a = synth.encode_ascii(synth.make_alignment(20000, seed=20251018 + 2, ambiguity=True))
    for m in ("n", "n_high"):
        square_or_rect("cfg2: n / n_high all-vs-all 20,000 x 29,903, 1% N/ambiguity/gaps", m, a, None, args, u16=True)


def cfg3(args):
    root = synth.make_root(W, 20251018 + 3)
    a = synth.encode_ascii(synth.make_alignment(10000, seed=20251018 + 3, ambiguity=True, root=root))
    b = synth.encode_ascii(synth.make_alignment(10000, seed=20251018 + 33, ambiguity=True, root=root))
    square_or_rect("cfg3: tn93 10,000 x 10,000 between two alignments", "tn93", a, b, args)


def cfg4(args):
    root = synth.make_root(W, 20251018 + 4)
    res = synth.encode_ascii(synth.make_alignment(1000, seed=20251018 + 4, ambiguity=True, root=root))
    pool = pinned(synth.encode_ascii(synth.make_alignment(args.pool, seed=20251018 + 44, ambiguity=True, root=root)))
    total = args.stream_records
    L = dg.load_library()
    for zero_copy in (False, True):
        e = dg.Engine("k80", W)
        e.load(0, pinned(res))
        state = {"n": 0}

        def sink(user, pp):
            state["n"] += int(pp.contents.n_results)
            return 0
        cb = api.SINK_FN(sink)
        batch = 4096
        e._check(L.dg_stream_begin(e.h, cb, None, batch))
        buf, cap = C.c_void_p(), C.c_uint64()
        if zero_copy:  # stand-in for a parser that writes straight into the pinned staging buffers
            for _ in range(2):
                e._check(L.dg_stream_buffer(e.h, C.byref(buf), C.byref(cap)))
                C.memmove(buf.value, pool.ctypes.data, int(cap.value) * W)
                e._check(L.dg_stream_push(e.h, buf, cap.value, api.DG_INPUT_PARADIS, None))
            e._check(L.dg_stream_end(e.h))
            state["n"] = 0
            e._check(L.dg_stream_begin(e.h, cb, None, batch))
        e.reset_timings()
        t0 = time.time()
        done = 0
        while done < total:
            nb = min(batch, total - done)
            if zero_copy:
                e._check(L.dg_stream_buffer(e.h, C.byref(buf), C.byref(cap)))
                e._check(L.dg_stream_push(e.h, buf, nb, api.DG_INPUT_PARADIS, None))
            else:
                off = (done % (args.pool - batch + 1))
                e._check(L.dg_stream_push(e.h, C.c_void_p(pool.ctypes.data + off * W), nb, api.DG_INPUT_PARADIS, None))
            done += nb
        e._check(L.dg_stream_end(e.h))
        wall = 1e3 * (time.time() - t0)
        tm = e.timings()
        pairs = total * 1000
        assert state["n"] == pairs
        extra = {"engine": int(tm["engine"]), "streamed_records": total, "batch": batch, "h2d_bytes": total * W,
                 "d2h_bytes": pairs * 8, "count_kernels_ms": tm["count_ms"], "pack_kernels_ms": tm["pack_ms"],
                 "host_staging_copy": not zero_copy,
                 "note": "kernel_only = sum of the per-batch device spans (pack + tiles + combine); e2e = wall of the whole "
                         "stream session incl. H2D of every batch and D2H of every result panel"}
        if zero_copy:
            extra["cpu_baseline"] = cpu_sample("k80", "stream", res, np.asarray(pool[:2048]), args.cpu_budget)
        line("cfg4: k80, 1,000 resident vs streamed records (pinned double-buffered batches)", "k80", pairs,
             tm["count_ms"] + tm["pack_ms"], wall, extra)
        e.close()


def cfg5(args):
    n = args.n5
    a = synth.encode_ascii(synth.make_alignment(n, seed=20251018 + 5, ambiguity=True))
    square_or_rect(f"cfg5: jc69 all-vs-all {n:,} x 29,903 -- one GPU's 1/8 share (panels k % 8 == 0)", "jc69", a, None, args,
                   part=0, n_parts=8)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="1,2,3,4,5")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=8.0)
    ap.add_argument("--pool", type=int, default=16384)
    ap.add_argument("--stream-records", type=int, default=1000000)
    ap.add_argument("--n5", type=int, default=100000)
    args = ap.parse_args()
    dg.load_library()
    for k in args.only.split(","):
        {"1": cfg1, "2": cfg2, "3": cfg3, "4": cfg4, "5": cfg5}[k](args)
