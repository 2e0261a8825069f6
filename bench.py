#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the pairwise-comparison hot path.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                   (the CPU reference arm, rank 0 only)

Headline workload (BASELINE.json configs[1]): `-m n_high` all-vs-all over 20,000 synthetic SARS-CoV-2-length
records (29,903 nt, 1% N / ambiguity codes / gaps).  A step = one pass of the hot path (operand pack -> count
tiles -> results) over that alignment.  For N>1 the run is WEAK-scaled: n = round(20000 * sqrt(N)) records so
every GPU keeps ~2.0e8 pairs per step; ranks own disjoint result panels (dg_run_part) and exchange nothing but
the timing reductions.

  value     : pairs/s, whole job, inputs resident in HBM, device time from CUDA events on the library's compute
              streams (dg_timings.run_ms), MAX over ranks.
  e2e       : same metric through the C ABI with HOST buffers: H2D of the codes from pinned memory, pack, tiles,
              D2H of every result panel into pinned memory and the sink callback, per step.
  sustained : the same step repeated for >= 3 s with the SM clock and throttle reasons beside it.
  configs   : BASELINE configs 1, 3, 4, 5 (and 2 again) in the same line: kernel-only ms, pairs/s, roofline
              fraction, e2e ms, engine and an in-run SAMPLED ORACLE CHECK (>= 2,000 pairs per config, NaN / inf /
              -0.0 included; outside every timed region; a mismatch aborts the run).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT = sys.stdout
MEASURE = "n_high"
BASE_N = 20000
WIDTH = 29903
SEED = 20251018 + 2
REL_TOL = 1e-12   # north_star: raw / jc69 / k80 / tn93 within 1e-12 relative; counts bit-exact
# algorithmic 32-bit lane-ops per pair-site (4 LOP3 + 1 POPC + 1 IADD per 32 sites)
OPS_PER_PAIR_SITE = {"n": 0.1875, "n_high": 0.1875, "raw": 0.28125, "jc69": 0.28125, "k80": 0.5, "tn93": 0.5}
# tensor engine: ops (2 per MAC) per pair-site of the minimal-rank schedules in tc_engine.cuh / DESIGN.md 3.1
# (SURVEY 8d budgeted 10 / 10 / 14 / 14; the 4-MAC DIFF and the W/Z factorisation need fewer)
I8_OPS_PER_PAIR_SITE = {"n": 8, "n_high": 8, "raw": 16, "jc69": 16, "k80": 12, "tn93": 10}
ENGINE_NAME = {1: "lop3_popc", 2: "tcgen05_i8", 3: "tcgen05_mxf4"}
WORKLOAD = f"config 2: -m {MEASURE} all-vs-all, 20,000 x 29,903 nt (n = round(20000*sqrt(N)) for N GPUs), 1% N/ambiguity/gaps"


def load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0=None, t1=None):
        rows = [r for ts, r in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.2)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        return self.window(t0, t1)


def make_workload(n: int):
    from distance_b200 import synth
    asc = synth.make_alignment(n, width=WIDTH, seed=SEED, ambiguity=True)
    return synth.encode_ascii(asc)


def cpu_baseline(codes, threads: int, budget_s: float):
    """The oracle (C port of the reference's per-pair loops, -t = all host threads) on a bounded
    sample: the first R major rows of the same all-vs-all, R sized for ~budget_s of CPU time."""
    from oracle import oracle as orc
    a = orc.Alignment(codes)
    orc.prepare(MEASURE, [a])
    n = a.n
    probe_rows = max(threads, 8)
    pairs, secs = orc.bench(MEASURE, "square", a, None, probe_rows, threads)
    rate = pairs / max(secs, 1e-9)
    rows = int(max(probe_rows, min(n - 1, budget_s * rate / n)))
    rows = rows // threads * threads or threads
    pairs, secs = orc.bench(MEASURE, "square", a, None, rows, threads)
    return pairs / secs, pairs, secs, rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = int(round(BASE_N * math.sqrt(args.gpus))) if args.n is None else args.n
    n_sample = min(n, 4000)
    codes = make_workload(n_sample)  # the sample only touches the first rows x all columns
    threads = os.cpu_count() or 1
    from oracle import oracle as orc
    a = orc.Alignment(codes)
    orc.prepare(MEASURE, [a])
    # size one step for ~2 s so W+K steps end within a few minutes
    pairs, secs = orc.bench(MEASURE, "square", a, None, threads, threads)
    rows = max(threads, int(2.0 * (pairs / secs) / a.n) // threads * threads)
    rows = min(rows, a.n - 1)
    for _ in range(args.warmup):
        orc.bench(MEASURE, "square", a, None, rows, threads)
    tot_pairs, tot_s = 0, 0.0
    for _ in range(args.steps):
        p, s = orc.bench(MEASURE, "square", a, None, rows, threads)
        tot_pairs += p; tot_s += s
    value = tot_pairs / tot_s
    sample = (f"first {rows} rows x {a.n} records of the synthetic alignment per step "
              f"({tot_pairs // args.steps} pairs/step), oracle C port, {threads} pthreads")
    line = {
        "impl": "reference", "metric": "pairwise distances/sec", "value": value, "unit": "pairs/s",
        "pair_sites_per_s": value * WIDTH, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak" if args.n is None else "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "measure": MEASURE, "n": n, "width": WIDTH,
                   "reference_arm_records": n_sample, "reference_arm_rows_per_step": rows,
                   "reference_arm_sample": sample,
                   "reference_arm_note": "the CPU arm times a bounded sample (the first rows of an alignment truncated to "
                                         f"{n_sample} records): CPU pairs/s does not depend on n, the per-pair cost is O(width)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated reference (C oracle), not the Rust binary: no Rust toolchain in this image",
    }
    print(json.dumps(line), file=OUT, flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# sampled oracle check (outside every timed region)
# ---------------------------------------------------------------------------------------------------------------
class ParityError(RuntimeError):
    pass


def compare_sample(measure, got, want, where):
    """counts bit-exact; floats: NaN / +-inf / +-0.0 in the same places, the rest within REL_TOL.  Returns the
    number of NaN, inf and zero (incl. -0.0) values the sample hit."""
    got, want = np.asarray(got), np.asarray(want)
    if measure in ("n", "n_high"):
        if not np.array_equal(got.astype(np.int64), want.astype(np.int64)):
            bad = np.flatnonzero(got.astype(np.int64) != want.astype(np.int64))[:5]
            raise ParityError(f"{where}: counts differ at sample positions {bad.tolist()}: got {got[bad].tolist()} want {want[bad].tolist()}")
        return {"nan": 0, "inf": 0, "zero": int((want == 0).sum())}
    got, want = got.astype(np.float64), want.astype(np.float64)
    nan_w, inf_w = np.isnan(want), np.isinf(want)
    if not np.array_equal(np.isnan(got), nan_w):
        raise ParityError(f"{where}: NaN positions differ")
    if not (np.array_equal(np.isinf(got), inf_w) and np.array_equal(got[inf_w], want[inf_w])):
        raise ParityError(f"{where}: inf positions / signs differ")
    fin = ~(nan_w | inf_w)
    g, w = got[fin], want[fin]
    zero = w == 0.0
    if not (np.array_equal(g[zero], w[zero]) and np.array_equal(np.signbit(g[zero]), np.signbit(w[zero]))):
        raise ParityError(f"{where}: zeros (sign of -0.0 included) differ")
    rel = np.abs(g[~zero] - w[~zero]) / np.abs(w[~zero])
    if rel.size and rel.max() > REL_TOL:
        raise ParityError(f"{where}: max relative error {rel.max():.3e} > {REL_TOL}")
    return {"nan": int(nan_w.sum()), "inf": int(inf_w.sum()), "zero": int(zero.sum())}


def oracle_pairs(measure, row_codes, col_codes, swap=False):
    """The oracle's per-pair functions (C port of measures.rs) for one row record against some column records.
    swap: stream mode, where the streamed (row) record is the reference's `target` (lib.rs:322-325)."""
    from oracle import oracle as orc
    if measure in ("n", "n_high"):
        return np.array([orc.snp(row_codes, c) for c in col_codes], dtype=np.int64)
    if measure == "tn93":
        rc = orc.count_bases(row_codes)
        out = []
        for c in col_codes:
            cc = orc.count_bases(c)
            out.append(orc.tn93(c, row_codes, cc, rc) if swap else orc.tn93(row_codes, c, rc, cc))
        return np.array(out)
    f = {"raw": orc.raw, "jc69": orc.jc69, "k80": orc.k80}[measure]
    return np.array([f(c, row_codes) if swap else f(row_codes, c) for c in col_codes])


def merge_hits(a, b):
    return {k: a.get(k, 0) + b.get(k, 0) for k in set(a) | set(b)}


def spike(asc, rng):
    """Three planted records so that the sampled check meets the special values: an all-N record (no compared site:
    NaN), a copy of its neighbour (identical pair: -0.0 / 0.0), a random-base record (saturated distances)."""
    n, w = asc.shape
    if n >= 16:   # at both ends: as COLUMNS the last ones meet every sampled row of an all-vs-all, whichever part owns it
        for base in (0, n - 8):
            asc[base + 3, :] = ord("N")
            asc[base + 5, :] = asc[base + 4, :]
            asc[base + 7, :] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=w)]
    return asc


def special_indices(n):
    return [3, 4, 5, 7, n - 5, n - 4, n - 3, n - 1] if n >= 16 else []


class RowGrabber:
    """A sink that keeps the result rows of a few sampled major rows (everything else is only counted)."""

    def __init__(self, api, mode, n_rows_total, n_cols, dtype, rows):
        self.api, self.mode, self.n, self.n_cols, self.dtype = api, mode, n_rows_total, n_cols, np.dtype(dtype)
        self.rows = {int(r): None for r in rows}
        self.pairs = 0
        self.cb = api.SINK_FN(self._sink)

    def _sink(self, user, pp):
        p = pp.contents
        r0, r1 = int(p.row_begin), int(p.row_end)
        self.pairs += int(p.n_results)
        isz = self.dtype.itemsize
        whole = None
        if int(p.result_kind) == self.api.DG_RESULT_U8:   # uint8 + overflow list: widen the panel once, slice below
            whole = self.api.panel_values(p)
        for r in self.rows:
            if r0 <= r < r1:
                if self.mode == self.api.DG_MODE_SQUARE:
                    off = r * (2 * self.n - r - 1) // 2 - r0 * (2 * self.n - r0 - 1) // 2
                    ln = self.n - 1 - r
                else:
                    off, ln = (r - r0) * self.n_cols, self.n_cols
                if ln and whole is not None:
                    self.rows[r] = whole[off:off + ln].copy()
                elif ln:
                    buf = (C.c_uint8 * (ln * isz)).from_address(p.data + off * isz)
                    self.rows[r] = np.frombuffer(buf, dtype=self.dtype, count=ln).copy()
                else:
                    self.rows[r] = np.zeros(0, self.dtype)
        return 0


def sample_rows_of(panels, specials, rng, per_panel=2, max_panels=4):
    """rows to check: the planted special rows that fall into this rank's panels + a few rows of a few panels"""
    rows = set()
    if not panels:
        return []
    pick = [panels[0], panels[-1]] + [panels[i] for i in rng.integers(0, len(panels), size=max(0, max_panels - 2))]
    for (r0, r1, _) in pick:
        rows.update({r0, r1 - 1})
        rows.update(int(x) for x in rng.integers(r0, r1, size=per_panel))
    for r in specials:
        if any(r0 <= r < r1 for (r0, r1, _) in panels):
            rows.add(r)
    return sorted(rows)


def check_rows(measure, mode_name, grab, row_codes_of, col_codes_of, n_cols_total, rng, per_row, label, swap=False):
    """Compare `per_row` columns of every grabbed row with the oracle."""
    hits, n_checked = {}, 0
    for r, vals in grab.rows.items():
        if vals is None:
            raise ParityError(f"{label}: sampled row {r} never reached the sink")
        first = r + 1 if mode_name == "square" else 0
        avail = n_cols_total - first
        if avail <= 0:
            continue
        k = min(per_row, avail)
        cols = np.unique(np.concatenate([rng.integers(first, n_cols_total, size=k), np.arange(first, min(n_cols_total, first + 24)),
                                         [n_cols_total - 1]]))
        specials = [c for c in special_indices(n_cols_total) if first <= c < n_cols_total]
        cols = np.unique(np.concatenate([cols, np.array(specials, dtype=cols.dtype)])) if specials else cols
        want = oracle_pairs(measure, row_codes_of(r), [col_codes_of(int(c)) for c in cols], swap=swap)
        got = vals[cols - first]
        hits = merge_hits(hits, compare_sample(measure, got, want, f"{label} row {r}"))
        n_checked += int(cols.size)
    return n_checked, hits


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs 1, 3, 4, 5 (+ 2): one entry each in the JSON line
# ---------------------------------------------------------------------------------------------------------------
def fp4_peak_tops():
    peaks = load_json(os.path.join(ROOT, "profiles", "int_peaks.json")) or {}
    return peaks.get("fp4_tops_measured") or 9000.0, peaks.get("int8_tops_measured") or 4500.0


def tensor_frac(measure, pairs, ms, n_gpus, engine):
    fp4, i8 = fp4_peak_tops()
    peak = fp4 if engine == 3 else i8
    if engine not in (2, 3) or ms <= 0:
        return None
    return pairs * WIDTH * I8_OPS_PER_PAIR_SITE[measure] / (ms * 1e-3) / 1e12 / (peak * n_gpus)


def pinned_copy(api, a):
    p = api.pinned_array(a.shape, np.uint8)
    p[...] = a
    return p


def pinned_nibbles(api, synth, asc):
    """ASCII alignment -> DG_INPUT_NIBBLE rows in pinned memory (what a parser that packs while it validates hands over),
    in row blocks on a few threads so that a 100,000-record alignment needs neither multi-GB temporaries nor 20 s."""
    from concurrent.futures import ThreadPoolExecutor
    n, w = asc.shape
    nlut = (synth.ascii_lut() >> 4).astype(np.uint8)
    out = api.pinned_array((n, (w + 1) // 2), np.uint8)

    def block(r0):
        t = nlut[asc[r0:r0 + 2048]]
        if w % 2:
            t = np.concatenate([t, np.full((t.shape[0], 1), 15, np.uint8)], axis=1)
        np.bitwise_or(t[:, 0::2], t[:, 1::2] << 4, out=out[r0:r0 + t.shape[0]])

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(block, range(0, n, 2048)))
    return out


def config_square_or_rect(cfg_id, label, measure, a_asc, b_asc, d, args, dg, api, synth, parts_of=None):
    """All-vs-all (b_asc None) or two-file run of one config on this rank's part: kernel-only, pipelined e2e, sampled
    oracle check.  Inputs are ASCII (the LUT of encoding.rs runs on the device: DG_INPUT_ASCII)."""
    rank, world = d.rank, d.world
    part, parts = parts_of if parts_of else (rank, world)
    mode = api.DG_MODE_SQUARE if b_asc is None else api.DG_MODE_RECT
    mode_name = "square" if b_asc is None else "rect"
    lut = synth.ascii_lut()
    is_int = measure in ("n", "n_high")
    eng = dg.Engine(measure, WIDTH, gpus=[d.local_rank])
    eng.set_option(api.DG_OPT_KEEP_CODES, 1)
    if is_int:
        eng.set_option(api.DG_OPT_RESULT_U16, 1)
        eng.set_option(api.DG_OPT_RESULT_U8, 1)
        eng.set_option(api.DG_OPT_PANEL_BYTES, 128 << 20)
    pa = pinned_copy(api, a_asc)
    pb = None if b_asc is None else pinned_copy(api, b_asc)
    # e2e legs: DG_INPUT_NIBBLE rows (two sites per byte), packed outside the timed region like a parser would
    na = pinned_nibbles(api, synth, a_asc)
    nb_ = None if b_asc is None else pinned_nibbles(api, synth, b_asc)
    eng.load(0, pa, input_kind=api.DG_INPUT_ASCII)
    if pb is not None:
        eng.load(1, pb, input_kind=api.DG_INPUT_ASCII)
    n_rows, n_cols = a_asc.shape[0], (a_asc.shape[0] if b_asc is None else b_asc.shape[0])
    if parts > 1:
        # panels are dealt to the parts round-robin: cut the job into >= 4 panels per part so that the shares are even
        total_bytes = (n_rows * (n_rows - 1) // 2 if b_asc is None else n_rows * n_cols) * (2 if is_int else 8)
        eng.set_option(api.DG_OPT_PANEL_BYTES, int(max(8 << 20, min((128 << 20) if is_int else (256 << 20), total_bytes // (4 * parts)))))
    plan = eng.plan(mode)
    from distance_b200 import dist as _dist
    mine = _dist.my_panels(plan, part, parts)
    my_pairs = sum(p[2] for p in mine)
    # kernel-only
    for _ in range(2):
        eng.run_device_only(mode, part, parts, repack=True)
    d.barrier()
    eng.reset_timings()
    ms = 0.0
    for _ in range(args.cfg_steps):
        eng.run_device_only(mode, part, parts, repack=True)
        ms += eng.timings()["run_ms"]
    d.barrier()
    tm = eng.timings()
    k_ms = d.max(ms / args.cfg_steps)
    pairs = int(d.sum(my_pairs))
    engine = int(tm["engine"])
    # e2e: the pipelined session from pinned host memory (alignment 1 of a two-file run is uploaded inside the step too)
    state = {"n": 0}

    def sink(user, pp):
        state["n"] += int(pp.contents.n_results)
        return 0

    cb = api.SINK_FN(sink)

    def e2e_step():
        state["n"] = 0
        if pb is None:
            eng._check(eng.L.dg_run_square_host(eng.h, C.c_void_p(na.ctypes.data), n_rows, api.DG_INPUT_NIBBLE, None, part, parts, cb, None))
        else:
            eng.load(1, nb_, input_kind=api.DG_INPUT_NIBBLE)
            eng._check(eng.L.dg_run_rect_host(eng.h, C.c_void_p(na.ctypes.data), n_rows, api.DG_INPUT_NIBBLE, None, part, parts, cb, None))
        return state["n"]

    e2e_step()
    d.barrier()
    t0 = time.time()
    for _ in range(args.cfg_e2e_steps):
        got_n = e2e_step()
    d.barrier()
    e2e_ms = d.max(1e3 * (time.time() - t0) / args.cfg_e2e_steps)
    e2e_pairs = int(d.sum(got_n))
    eng.reset_timings()
    e2e_step()
    e2e_d2h = int(eng.timings()["d2h_bytes"])   # counted by the library from the copies it issued
    # sampled oracle check through the in-order path (dg_load_resident + dg_run_part): resident again after the session
    rng = np.random.default_rng(1000 + cfg_id * 10 + rank)
    eng.load(0, pa, input_kind=api.DG_INPUT_ASCII)
    if pb is not None:
        eng.load(1, pb, input_kind=api.DG_INPUT_ASCII)
    rows = sample_rows_of(mine, special_indices(n_rows), rng)
    dtype = (np.uint16 if is_int else np.float64)
    grab = RowGrabber(api, mode, n_rows, n_cols, dtype, rows)
    eng._check(eng.L.dg_run_part(eng.h, mode, part, parts, grab.cb, None, 0))
    if grab.pairs != my_pairs:
        raise ParityError(f"{label}: the sink saw {grab.pairs} results, the plan has {my_pairs}")
    cols_asc = a_asc if b_asc is None else b_asc
    n_chk, hits = check_rows(measure, mode_name, grab, lambda r: lut[a_asc[r]], lambda c: lut[cols_asc[c]], n_cols, rng,
                             max(64, 2400 // max(1, len(rows))), label)
    n_chk = int(d.sum(n_chk))
    hits = {k: int(d.sum(hits.get(k, 0))) for k in ("inf", "nan", "zero")}   # the same collectives on every rank
    eng.close()
    return {"id": cfg_id, "workload": label, "measure": measure, "mode": mode_name, "pairs": pairs,
            "kernel_ms": k_ms, "pairs_per_s": pairs / (k_ms * 1e-3), "pair_sites_per_s": pairs * WIDTH / (k_ms * 1e-3),
            "roofline_frac": tensor_frac(measure, pairs, k_ms, world, engine), "engine": ENGINE_NAME.get(engine, "?"),
            "sm_mhz_in_kernel": tm.get("sm_mhz"),
            "e2e_ms": e2e_ms, "e2e_pairs_per_s": e2e_pairs / (e2e_ms * 1e-3),
            "h2d_bytes_per_step": int(na.nbytes + (0 if nb_ is None else nb_.nbytes)), "d2h_bytes_per_step": e2e_d2h,
            "e2e_input_kind": "DG_INPUT_NIBBLE", "e2e_result_kind": "u8 + overflow list" if is_int else "f64",
            "parity_sampled": "ok", "parity_pairs": n_chk, "special_values_hit": hits,
            "part": f"rank r runs part r of {parts}" if world > 1 or parts > 1 else "whole job"}


def config_stream(cfg_id, label, measure, d, args, dg, api, synth):
    """BASELINE config 4: 1,000 resident records against 1,000,000 streamed ones (every rank streams 1/N of them) in
    pinned double-buffered batches.  The two staging buffers of the ring are filled once with two different chunks of a
    synthetic pool (a parser would write the next batch there: dg_stream_buffer) and pushed over and over, so a step
    moves 15 GB / N (two sites per byte) up and 8 GB / N of f64 results down over PCIe without a host-side copy in the way."""
    rank, world = d.rank, d.world
    lut = synth.ascii_lut()
    rng = np.random.default_rng(1000 + cfg_id * 10 + rank)
    root = synth.make_root(WIDTH, 20251018 + 4)
    res_asc = spike(synth.make_alignment(1000, seed=20251018 + 4, ambiguity=True, root=root), rng)
    batch = 7168   # 14 row blocks x 5 column blocks x 3 accumulators = 210 work items = 2.84 rounds of the 74 CTA pairs (4,096: 1.62)
    pool_asc = spike(synth.make_alignment(2 * batch, seed=20251018 + 44 + rank, ambiguity=True, root=root), rng)
    pool_nib = api.pack_nibbles(lut[pool_asc])   # the streamed batches travel as DG_INPUT_NIBBLE rows (two sites per byte)
    nibw = pool_nib.shape[1]
    total = args.stream_records // world
    eng = dg.Engine(measure, WIDTH, gpus=[d.local_rank])
    eng.load(0, res_asc, input_kind=api.DG_INPUT_ASCII)
    L = eng.L
    st = {"n": 0, "batches": 0, "keep": {}}

    def sink(user, pp):
        p = pp.contents
        cnt = int(p.n_results)
        st["n"] += cnt
        b = st["batches"]
        if b in st["keep"] and st["keep"][b] is None:
            st["keep"][b] = np.frombuffer((C.c_uint8 * (cnt * 8)).from_address(p.data), dtype=np.float64, count=cnt).copy()
        st["batches"] += 1
        return 0

    cb = api.SINK_FN(sink)
    buf, cap = C.c_void_p(), C.c_uint64()

    def session(n_records, fill):
        st["n"], st["batches"] = 0, 0
        eng._check(L.dg_stream_begin(eng.h, cb, None, batch))
        done, k = 0, 0
        while done < n_records:
            nb = min(batch, n_records - done)
            eng._check(L.dg_stream_buffer(eng.h, C.byref(buf), C.byref(cap)))
            if fill and k < 2:   # the ring has two staging buffers per device: slot k % 2 keeps pool chunk k % 2
                C.memmove(buf.value, pool_nib[k * batch:(k + 1) * batch].ctypes.data, batch * nibw)
            eng._check(L.dg_stream_push(eng.h, buf, nb, api.DG_INPUT_NIBBLE, None))
            done += nb
            k += 1
        eng._check(L.dg_stream_end(eng.h))
        return st["n"]

    session(4 * batch, True)     # warm-up + fills both staging buffers
    d.barrier()
    eng.reset_timings()
    t0 = time.time()
    for _ in range(args.cfg_e2e_steps):
        got = session(total, False)
    d.barrier()
    wall_ms = d.max(1e3 * (time.time() - t0) / args.cfg_e2e_steps)
    tm = eng.timings()
    dev_ms = d.max((tm["count_ms"] + tm["pack_ms"]) / args.cfg_e2e_steps)
    pairs = int(d.sum(got))
    engine = int(tm["engine"])
    # sampled check: batches 0, 1 and a late one; batch k holds pool chunk k % 2
    n_b = (total + batch - 1) // batch
    st["keep"] = {k: None for k in sorted({0, 1, max(0, n_b - 2)})}
    session(total, False)
    hits, n_chk = {}, 0
    for k, vals in st["keep"].items():
        if vals is None:
            raise ParityError(f"{label}: batch {k} never reached the sink")
        rows_in = vals.size // 1000
        chunk = pool_asc[(k % 2) * batch:(k % 2) * batch + rows_in]
        for r in sorted(set([0, 3, 4, 5, 7, rows_in - 1] + [int(x) for x in rng.integers(0, rows_in, size=2)])):
            if r >= rows_in:
                continue
            cols = np.unique(np.concatenate([rng.integers(0, 1000, size=100), [0, 3, 4, 5, 7, 999]]))
            want = oracle_pairs(measure, lut[chunk[r]], [lut[res_asc[int(c)]] for c in cols], swap=True)
            hits = merge_hits(hits, compare_sample(measure, vals.reshape(rows_in, 1000)[r, cols], want, f"{label} batch {k} row {r}"))
            n_chk += int(cols.size)
    n_chk = int(d.sum(n_chk))
    hits = {k: int(d.sum(hits.get(k, 0))) for k in ("inf", "nan", "zero")}   # the same collectives on every rank
    eng.close()
    return {"id": cfg_id, "workload": label, "measure": measure, "mode": "stream", "pairs": pairs,
            "kernel_ms": dev_ms, "pairs_per_s": pairs / (dev_ms * 1e-3), "pair_sites_per_s": pairs * WIDTH / (dev_ms * 1e-3),
            "kernel_note": "sum of the per-batch device spans (operand pack + tiles + combine, CUDA events)",
            "roofline_frac": tensor_frac(measure, pairs, dev_ms, world, engine), "engine": ENGINE_NAME.get(engine, "?"),
            "sm_mhz_in_kernel": tm.get("sm_mhz"),
            "e2e_ms": wall_ms, "e2e_pairs_per_s": pairs / (wall_ms * 1e-3),
            "h2d_bytes_per_step": int(total * nibw), "d2h_bytes_per_step": int(total * 1000 * 8),
            "h2d_gbs": total * nibw / (wall_ms * 1e-3) / 1e9, "d2h_gbs": total * 8000 / (wall_ms * 1e-3) / 1e9,
            "e2e_input_kind": "DG_INPUT_NIBBLE",
            "streamed_records": total * world, "batch": batch,
            "parity_sampled": "ok", "parity_pairs": n_chk, "special_values_hit": hits,
            "part": f"every rank streams {total} records" if world > 1 else "whole job"}


def run_configs(d, args, dg, api, synth):
    out = []
    want = [int(x) for x in args.configs.split(",") if x]
    rng = np.random.default_rng(77)
    if 1 in want:
        asc = spike(synth.make_alignment(1000, seed=20251018 + 1, ambiguity=True), rng)
        out.append(config_square_or_rect(1, "config 1: -m raw all-vs-all, 1,000 x 29,903", "raw", asc, None, d, args, dg, api, synth))
    if 2 in want:
        n2 = BASE_N if args.n is None else args.n
        asc = spike(synth.make_alignment(n2, seed=SEED, ambiguity=True), rng)
        out.append(config_square_or_rect(2, f"config 2: -m n all-vs-all, {n2:,} x 29,903, 1% N/ambiguity/gaps (strong-scaled over the ranks)",
                                         "n", asc, None, d, args, dg, api, synth))
    if 3 in want:
        root = synth.make_root(WIDTH, 20251018 + 3)
        a = spike(synth.make_alignment(10000, seed=20251018 + 3, ambiguity=True, root=root), rng)
        b = spike(synth.make_alignment(10000, seed=20251018 + 33, ambiguity=True, root=root), rng)
        out.append(config_square_or_rect(3, "config 3: -m tn93 between two alignments, 10,000 x 10,000 x 29,903", "tn93", a, b, d, args, dg, api, synth))
    if 4 in want:
        out.append(config_stream(4, f"config 4: -m k80, -i 1,000 resident vs -s {args.stream_records:,} streamed records (pinned double-buffered batches)",
                                 "k80", d, args, dg, api, synth))
    if 5 in want:
        n5 = args.n5
        asc = spike(synth.make_alignment(n5, seed=20251018 + 5, ambiguity=True), rng)
        parts = max(8, d.world)
        share = "the whole job" if d.world >= 8 else f"{d.world}/8 of the job: rank r computes part r of 8 (what GPU r of 8 B200s does)"
        out.append(config_square_or_rect(5, f"config 5: -m jc69 all-vs-all, {n5:,} x 29,903 -- {share}", "jc69", asc, None, d, args, dg, api, synth,
                                         parts_of=(d.rank, parts)))
        del asc
    return out


def run_ours(args):
    # a hung collective or kernel must not sit on the GPUs until an outer limit fires: dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog, exit=True)
    import distance_b200 as dg
    from distance_b200 import api, dist, synth

    d = dist.Dist()
    rank, world = d.rank, d.world
    if args.gpus != world and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    n = int(round(BASE_N * math.sqrt(world))) if args.n is None else args.n
    dg.load_library()
    if dg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the library has no CPU fallback)")

    codes = make_workload(n)
    pinned = api.pinned_array(codes.shape, np.uint8)
    pinned[...] = codes
    # the e2e legs upload DG_INPUT_NIBBLE rows (two sites per byte: what a parser that packs while it validates hands over)
    nib = api.pack_nibbles(codes)
    pinned_nib = api.pinned_array(nib.shape, np.uint8)
    pinned_nib[...] = nib
    NIBW = nib.shape[1]
    total_pairs = n * (n - 1) // 2

    eng = dg.Engine(MEASURE, WIDTH, gpus=[d.local_rank])
    eng.set_option(api.DG_OPT_PANEL_BYTES, args.panel_bytes)
    eng.set_option(api.DG_OPT_KEEP_CODES, 1)
    if args.is_int and not args.u32_results:
        eng.set_option(api.DG_OPT_RESULT_U16, 1)   # counts <= width < 65536: lossless, half the D2H bytes
        if not args.u16_results:
            eng.set_option(api.DG_OPT_RESULT_U8, 1)   # e2e panels as uint8 + an overflow list for the counts >= 255
    if args.tile_variant:
        eng.set_option(api.DG_OPT_TILE_VARIANT, args.tile_variant)
    if args.engine:
        eng.set_option(api.DG_OPT_ENGINE, args.engine)
    eng.load(0, pinned)
    plan = eng.plan(api.DG_MODE_SQUARE)
    if world > 1:
        # ranks own the panels dg_plan_parts deals them: pick the panel size (<= the default) whose largest share is smallest
        best = None
        for pct in range(100, 59, -4):
            pb = args.panel_bytes * pct // 100
            eng.set_option(api.DG_OPT_PANEL_BYTES, pb)
            pl = eng.plan(api.DG_MODE_SQUARE)
            worst = max(sum(p[2] for p in dist.my_panels(pl, r, world)) for r in range(world))
            if best is None or worst < best[0]:
                best = (worst, pb, pl)
        args.panel_bytes, plan = best[1], best[2]
        eng.set_option(api.DG_OPT_PANEL_BYTES, args.panel_bytes)
    my_pairs = sum(p[2] for p in dist.my_panels(plan, rank, world))

    def sync_all():
        d.barrier()

    # ---- pack kernel alone (roofline_pack): a few steps with the packing serialised in front of the tiles ------
    eng.set_option(api.DG_OPT_REPACK_OVERLAP, 0)
    eng.run_device_only(api.DG_MODE_SQUARE, rank, world, repack=True)
    eng.reset_timings()
    for _ in range(3):
        eng.run_device_only(api.DG_MODE_SQUARE, rank, world, repack=True)
    pack_ms_step = eng.timings()["pack_ms"] / 3
    serial_step_ms = eng.timings()["run_ms"]
    eng.set_option(api.DG_OPT_REPACK_OVERLAP, 0 if args.no_repack_overlap else 1)

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        eng.run_device_only(api.DG_MODE_SQUARE, rank, world, repack=True)

    sampler = ClockSampler(d.local_rank)
    sampler.start()
    time.sleep(0.3)

    # ---- value: device-resident inputs, kernel-only, CUDA events --------------------------------
    sync_all()
    eng.reset_timings()
    t_wall0 = time.time()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.run_device_only(api.DG_MODE_SQUARE, rank, world, repack=True)
        dev_ms += eng.timings()["run_ms"]
    sync_all()
    t_wall1 = time.time()
    tm = eng.timings()
    step_ms = d.max(dev_ms / args.steps)
    wall_step_ms = d.max(1e3 * (t_wall1 - t_wall0) / args.steps)
    value = total_pairs / (step_ms * 1e-3)
    launches = int(d.sum(tm["pack_launches"] + tm["count_launches"]))
    count_launch_ms = tm["count_ms"] / max(tm["count_launches"], 1)
    sm_mhz_burst = tm.get("sm_mhz")

    # ---- sustained: the same step back to back for >= args.sustained_s seconds -----------------------------------------
    sustained = None
    if args.sustained_s > 0:
        sync_all()
        eng.reset_timings()
        ts0 = time.time()
        s_dev, s_steps = 0.0, 0
        while True:
            for _ in range(16):
                eng.run_device_only(api.DG_MODE_SQUARE, rank, world, repack=True)
                s_dev += eng.timings()["run_ms"]
            s_steps += 16
            # every rank runs the same number of steps: decide together
            if d.max(time.time() - ts0) >= args.sustained_s:
                break
        sync_all()
        ts1 = time.time()
        s_ms = d.max(s_dev / s_steps)
        s_wall_ms = d.max(1e3 * (ts1 - ts0) / s_steps)
        clk = sampler.window(ts0 + 0.2, ts1)
        sustained = {"seconds": ts1 - ts0, "steps": s_steps, "ms_per_step": s_ms, "value": total_pairs / (s_ms * 1e-3), "unit": "pairs/s",
                     "wall_ms_per_step": s_wall_ms, "value_by_wall_clock": total_pairs / (s_wall_ms * 1e-3),
                     "sm_mhz_in_kernel": eng.timings().get("sm_mhz"), "clocks": clk,
                     "burst_value": value, "sustained_over_burst": (total_pairs / (s_ms * 1e-3)) / value,
                     "note": "device time per step from CUDA events, steps back to back; sm_mhz_in_kernel = clock64 / globaltimer inside "
                             "the GEMM launches (nvidia-smi's 200 ms samples are in `clocks`)"}

    # ---- e2e: host buffers through the C ABI ----------------------------------------------------
    # The pipelined session (dg_square_*): chunks of the alignment go up highest records first, so the PCIe upload,
    # packing + tiles and the D2H of finished panels overlap; the sink reads every panel in pinned host memory.
    #   N = 1 : dg_run_square_host on the pinned host alignment.
    #   N > 1 : every code byte crosses PCIe ONCE: each rank uploads 1/N of the alignment from pinned host memory, ONE
    #           NCCL all-gather over NVLink completes it on every GPU, and the session takes its chunks from that device
    #           buffer (per-chunk collectives were faster at N = 2 but tie every rank's progress to every other rank's
    #           host; one collective per step keeps the ranks independent).
    e2e_state = {"n": 0, "acc": 0}

    def e2e_sink(user, pp):
        p = pp.contents
        e2e_state["n"] += int(p.n_results)
        if p.n_results:
            e2e_state["acc"] ^= C.c_uint32.from_address(p.data).value
        return 0

    e2e_cb = api.SINK_FN(e2e_sink)
    if world > 1:
        import torch
        import torch.distributed as td
        # The session takes its chunks highest records first, so the alignment is cut into BANDS of whole chunks, top down;
        # every rank uploads 1/N of each band from pinned host memory over its own PCIe link and one NCCL all-gather per
        # band (NVLink) completes it on every GPU.  No host synchronisation: a CUDA event behind each all-gather is handed
        # to dg_square_push (ready_event), so band b + 1 is uploaded and gathered while the tiles of band b run.
        eng.square_begin(n, e2e_cb, rank, world, input_kind=api.DG_INPUT_NIBBLE)
        chunks = eng.square_plan()                       # [(lo, hi)] in push order (descending)
        for lo, hi in chunks:
            eng.square_push(pinned_nib.ctypes.data + lo * NIBW, -1, lo, hi)   # finish this planning session (a warm-up step)
        eng.square_end()
        n_bands = max(1, min(args.e2e_bands, len(chunks)))
        per_band = -(-len(chunks) // n_bands)
        bands = []
        for b0 in range(0, len(chunks), per_band):
            grp = chunks[b0:b0 + per_band]
            blo, bhi = grp[-1][0], grp[0][1]             # records [blo, bhi)
            pb = -(-(bhi - blo) // world)                # records of one rank's piece
            plo, phi = min(bhi, blo + rank * pb), min(bhi, blo + (rank + 1) * pb)
            host_piece = torch.full((pb, NIBW), 255, dtype=torch.uint8).pin_memory()
            host_piece[:phi - plo] = torch.from_numpy(np.ascontiguousarray(nib[plo:phi]))
            bands.append({"chunks": grp, "blo": blo, "host": host_piece,
                          "dev": torch.empty((pb, NIBW), dtype=torch.uint8, device=d.device),
                          "gathered": torch.empty((world * pb, NIBW), dtype=torch.uint8, device=d.device),
                          "ev": torch.cuda.Event()})
        in_stream = torch.cuda.Stream(device=d.device)

        def e2e_step():
            e2e_state["n"] = 0
            with torch.cuda.stream(in_stream):
                for bd in bands:
                    bd["dev"].copy_(bd["host"], non_blocking=True)        # 1/N of the band over this rank's PCIe link
                    td.all_gather_into_tensor(bd["gathered"], bd["dev"])  # the rest over NVLink
                    bd["ev"].record(in_stream)
            eng.square_begin(n, e2e_cb, rank, world, input_kind=api.DG_INPUT_NIBBLE)
            for bd in bands:
                base = bd["gathered"].data_ptr()
                for lo, hi in bd["chunks"]:
                    lo2, hi2 = eng.square_next()
                    assert (lo2, hi2) == (lo, hi)
                    eng.square_push(base + (lo - bd["blo"]) * NIBW, d.local_rank, lo, hi, ready_event=bd["ev"].cuda_event)
            eng.square_end()
            return e2e_state["n"]
    else:
        def e2e_step():
            return eng.square_pipelined_discard(pinned_nib, rank, world, input_kind=api.DG_INPUT_NIBBLE)

    for _ in range(2):
        got = e2e_step()
        # the session cuts the triangle into its own (smaller) panels, part k % world per rank: together they cover it once
        assert int(d.sum(got)) == total_pairs, (got, total_pairs)
    sync_all()
    t0 = time.time()
    for _ in range(args.steps):
        e2e_step()
    sync_all()
    t1 = time.time()
    if args.trace_e2e:   # one more step with the library's device timeline of the session on stderr (rank 0)
        if rank == 0:
            os.environ["DG_TRACE"] = "1"
        e2e_step()
        os.environ.pop("DG_TRACE", None)
        sync_all()
    e2e_step_ms = d.max(1e3 * (t1 - t0) / args.steps)
    e2e_value = total_pairs / (e2e_step_ms * 1e-3)
    eng.reset_timings()
    e2e_step()
    e2e_d2h_bytes = int(d.sum(eng.timings()["d2h_bytes"]))   # counted by the library from the copies it issued (all ranks)

    # the in-order path (dg_load_resident, then dg_run_part: panels reach the sink in the reference's output order)
    e2e_inorder_ms = None
    if world == 1:
        for _ in range(2):
            eng.load(0, pinned_nib, input_kind=api.DG_INPUT_NIBBLE); eng.run_discard(api.DG_MODE_SQUARE, rank, world)
        t0i = time.time()
        for _ in range(max(3, args.steps // 3)):
            eng.load(0, pinned_nib, input_kind=api.DG_INPUT_NIBBLE)
            assert eng.run_discard(api.DG_MODE_SQUARE, rank, world) == my_pairs
        e2e_inorder_ms = 1e3 * (time.time() - t0i) / max(3, args.steps // 3)
        t1 = time.time()
    clocks = sampler.window(t_wall0, t_wall1)
    clocks["sm_mhz_in_kernel"] = sm_mhz_burst
    clocks["note"] = ("sm_mhz = median of nvidia-smi's 200 ms samples over the timed region; sm_mhz_in_kernel = clock64 / globaltimer "
                      "around the GEMM launches of the timed steps (CTA 0), which sees the dips the 200 ms samples miss")

    # ---- headline parity: sampled rows of this rank's panels against the oracle (outside the timed regions) ---------------
    rng = np.random.default_rng(2000 + rank)
    eng.load(0, pinned)
    mine = dist.my_panels(plan, rank, world)
    rows = sample_rows_of(mine, (), rng)
    grab = RowGrabber(api, api.DG_MODE_SQUARE, n, n, np.uint16 if (args.is_int and not args.u32_results) else (np.uint32 if args.is_int else np.float64), rows)
    eng._check(eng.L.dg_run_part(eng.h, api.DG_MODE_SQUARE, rank, world, grab.cb, None, 0))
    assert grab.pairs == my_pairs
    head_chk, head_hits = check_rows(MEASURE, "square", grab, lambda r: codes[r], lambda c: codes[c], n, rng,
                                     max(64, 2400 // max(1, len(rows))), "headline")
    head_chk = int(d.sum(head_chk))

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------------
    ops = my_pairs * WIDTH * OPS_PER_PAIR_SITE[MEASURE]
    count_ms_step = tm["count_ms"] / args.steps
    run_ms_step = dev_ms / args.steps
    peaks = load_json(os.path.join(ROOT, "profiles", "int_peaks.json")) or {}
    measured = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    achieved = ops / (run_ms_step * 1e-3) / 1e12
    ops_per_word = OPS_PER_PAIR_SITE[MEASURE] * 32
    if "lop3_lane_ops_per_s" in peaks:
        # Issue ceiling of this instruction mix from the MEASURED per-pipe peaks: LOP3 issues on the ALU
        # pipe, POPC on the XU pipe, the accumulate (IMAD.IADD) on the FMA pipe; the slowest pipe bounds
        # the word-pair rate, and peak = that rate x the algorithmic ops per word.
        mix = peaks["per_word"][MEASURE]
        words_per_s = min(peaks["lop3_lane_ops_per_s"] / mix["lop3"], peaks["popc_lane_ops_per_s"] / mix["popc"],
                          peaks["imad_lane_ops_per_s"] / mix["iadd"])
        peak = words_per_s * ops_per_word / 1e12
        peak_src = ("measured (tools/ubench_int, profiles/int_peaks.json): min over pipes of "
                    "LOP3 1.85e13/4, POPC 4.39e12/1, IMAD 3.68e13/1 word-pairs/s x 6 algorithmic ops/word")
        frac_lop3 = achieved * 1e12 / peaks["lop3_lane_ops_per_s"]
    else:
        mhz = clocks.get("sm_mhz") or measured.get("sm_max_mhz", 1965.0)
        peak = 16 * 148 * mhz * 1e6 * ops_per_word / 1e12
        peak_src = f"nominal 16 word-pairs/clk/SM x 148 SM x {mhz:.0f} MHz (no measured int peaks file)"
        frac_lop3 = achieved / (64 * 148 * mhz * 1e6 / 1e12)
    roofline_lop3 = {
        "bound": "int_issue", "kernel": "count_tile_kernel<FAM_SNP>", "achieved": achieved, "peak": peak,
        "unit": "Tlaneop/s", "frac": achieved / peak, "traffic": peaks.get("count_kernel_dram_bytes_per_launch"),
        "ops_per_pair_site": OPS_PER_PAIR_SITE[MEASURE], "peak_source": peak_src,
        "frac_vs_lop3_peak_alone": frac_lop3,
        "word_pairs_per_clk_per_sm": my_pairs * math.ceil(WIDTH / 32) / (run_ms_step * 1e-3) / 148 / ((clocks.get("sm_mhz") or 1965.0) * 1e6),
        "avg_launch_ms": count_launch_ms, "count_ms_per_step": count_ms_step,
        "note": "integer-issue bound (LOP3 on the ALU pipe and POPC on the XU pipe saturate together), not HBM or "
                "tensor bound (SURVEY 8d); consecutive panel launches overlap on two streams, so achieved uses "
                "the device time of the whole step; traffic = DRAM bytes of one profiled launch (see profiles/)",
    }
    engine_id = int(tm.get("engine", 1))
    if engine_id in (2, 3):
        # tcgen05 engines: executed ALGORITHMIC tensor ops (2 per MAC) of this rank's launches /
        # device time of the step, against the MEASURED issue rate of the instruction kind that ran.
        i8ops = my_pairs * WIDTH * float(I8_OPS_PER_PAIR_SITE[MEASURE])
        ach = i8ops / (run_ms_step * 1e-3) / 1e12
        if engine_id == 3:
            pk8 = peaks.get("fp4_tops_measured")
            src8 = ("measured: tools/ubench_fp4 (profiles/ubench_fp4_r01.json), tcgen05.mma kind::mxf4.block_scale M128 N256 K64 SS "
                    "on 148 SMs; nominal dense fp4 = 9,000 TOPS")
            kname, unit = "tc_gemm_kernel<FP4> (tcgen05.mma kind::mxf4.block_scale, unit scales, TMA, TMEM)", "TOP/s (fp4; the spec's TFLOP/s slot)"
            if not pk8:
                pk8, src8 = 9000.0, "nominal dense fp4 of B200_PROFILING.md (no measured file)"
        else:
            pk8 = peaks.get("int8_tops_measured")
            src8 = "measured: tools/ubench_tc (profiles/ubench_tc_r01.json), tcgen05.mma kind::i8 M128 N256 K32 SS on 148 SMs"
            kname, unit = "tc_gemm_kernel (tcgen05.mma kind::i8, TMA, TMEM)", "TOP/s (int8; the spec's TFLOP/s slot)"
            if not pk8:
                pk8 = 2.0 * measured.get("bf16_tflops", 1590.0)
                src8 = "2 x the measured bf16 cuBLAS figure of MEASURED_PEAKS.json (kind::i8 is nominally 2 x bf16)"
        roofline = {
            "bound": "tensor", "kernel": kname, "achieved": ach, "peak": pk8,
            "unit": unit, "frac": ach / pk8,
            "frac_of_int8_peak": ach / peaks["int8_tops_measured"] if peaks.get("int8_tops_measured") else None,
            # context: MEASURED_PEAKS.json only carries dense bf16 (cuBLAS); nominal fp4 = 4 x bf16
            "measured_peaks_json_bf16_tflops": measured.get("bf16_tflops"),
            "frac_of_4x_measured_bf16": (ach / (4.0 * measured["bf16_tflops"])) if measured.get("bf16_tflops") else None,
            "traffic": peaks.get("tc_kernel_dram_bytes_per_launch"),
            "ops_per_pair_site": I8_OPS_PER_PAIR_SITE[MEASURE], "peak_source": src8,
            "frac_in_survey_units": ach / pk8 * 10.0 / I8_OPS_PER_PAIR_SITE[MEASURE],
            # the same achieved rate against the hardware limit AT THE CLOCK THE KERNEL SAW (148 SMs x 16,384 fp4 / 8,192 int8 MAC
            # per clock): the chip runs at its power cap, and the clock it settles at varies by several percent between runs
            "frac_of_hw_peak_at_in_kernel_clock": (ach * 1e12 / (148 * (16384 if engine_id == 3 else 8192) * 2 * sm_mhz_burst * 1e6))
            if sm_mhz_burst else None,
            "padded_frac": ach / pk8 * (math.ceil(WIDTH / 128) * 128) / WIDTH,
            "avg_launch_ms": count_launch_ms, "count_ms_per_step": count_ms_step,
            "note": "achieved = executed algorithmic tensor ops (4 MAC per pair-site: DIFF as a rank-4 bilinear form, "
                    "DESIGN.md 3.1) / device time of the whole step (operand re-pack included); frac_in_survey_units "
                    "scores the same time against SURVEY 8d's 5-MAC budget.  padded_frac counts the 49 zero-padded "
                    "sites per K block as work done.  Engine chosen automatically (DG_OPT_ENGINE=0): E2M1 operands "
                    "with unit block scales and fp32 accumulation are exact for these integer sums (tools/ubench_fp4, "
                    "parity suite; widths above 2^22 - 256 sites route to kind::i8), so the fp4 tensor rate applies; "
                    "frac_of_int8_peak scores the same ops against kind::i8.",
        }
    else:
        roofline = roofline_lop3
    hbm = measured.get("hbm_gbs")
    if engine_id == 3:
        pack_bytes = n * WIDTH + n * 8 * math.ceil(WIDTH / 256) * 128  # read 1 B/site, write 8 E2M1 planes (U, V), 2 sites per byte
        pack_kernel = "pack_ops_kernel<FP4>"
    elif engine_id == 2:
        pack_bytes = n * WIDTH + n * 8 * math.ceil(WIDTH / 128) * 128  # read 1 B/site, write 8 int8 planes (U, V)
        pack_kernel = "pack_ops_kernel<int8>"
    else:
        pack_bytes = n * WIDTH + n * math.ceil(WIDTH / 32) * 16  # read 1 B/site, write the 4 core bit-planes
        pack_kernel = "pack_planes_kernel"
    roofline_pack = {"bound": "hbm", "kernel": pack_kernel, "achieved": pack_bytes / (pack_ms_step * 1e-3) / 1e9,
                     "peak": hbm, "unit": "GB/s", "frac": (pack_bytes / (pack_ms_step * 1e-3) / 1e9 / hbm) if hbm else None,
                     "ms": pack_ms_step,
                     "note": "timed alone (DG_OPT_REPACK_OVERLAP=0: %.3f ms per step with the packing serialised in front of the tiles); in the "
                             "timed steps the packing of the next panel's records overlaps the tiles of the current one" % serial_step_ms}
    eng.close()
    del pinned

    # ---- the other BASELINE configs ------------------------------------------------------------------------------------------
    configs = []
    if args.configs:
        configs = run_configs(d, args, dg, api, synth)
    sampler.stop()

    line = None
    if rank == 0:
        cb_value, cb_pairs, cb_secs, cb_rows = (None, 0, 0.0, 0)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cb_value, cb_pairs, cb_secs, cb_rows = cpu_baseline(codes, threads, args.cpu_budget)
            cpu = {"value": cb_value, "unit": "pairs/s", "cores": threads, "kind": "port",
                   "sample": f"first {cb_rows} rows x {n} records of the same alignment ({cb_pairs} pairs, {cb_secs:.1f} s), "
                             f"oracle C port of measures.rs snp, {threads} pthreads"}
        else:
            cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "port", "sample": "skipped (N>1 or --no-cpu-baseline)"}
        line = {
            "metric": "pairwise distances/sec", "value": value, "unit": "pairs/s",
            "pair_sites_per_s": value * WIDTH, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "wall_ms_per_step": wall_step_ms, "higher_is_better": True, "scaling": "weak" if args.n is None else "strong",
            "vs_baseline": None, "dtype": {3: "fp4 (E2M1 planes, unit UE8M0 block scales, fp32 accumulation: exact integer sums; tcgen05 kind::mxf4)",
                                          2: "i8 (int8 planes, int32 accumulation, tcgen05 kind::i8)"}.get(engine_id, "u32 bit-planes (LOP3+POPC)"), "data": "synthetic",
            "config": {"workload": WORKLOAD, "measure": MEASURE, "n": n, "width": WIDTH, "pairs_per_step": total_pairs,
                       "weak_scaling": "n = round(20000*sqrt(N)) so pairs per GPU stay ~2.0e8",
                       "panel_bytes": args.panel_bytes, "panels": len(plan),
                       "l2": "inputs larger than L2 (operand planes %.0f MB vs 126 MB L2)" % (n * 8 * 14976 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": e2e_step_ms,
                    "h2d_bytes_per_step": int(n * NIBW),
                    "input_kind": "DG_INPUT_NIBBLE: two sites per byte (the possibility half of the Paradis code), packed by the host before "
                                  "the timed region like a parser would; unpacked to code bytes on the device",
                    "input_path": ("the alignment goes up in %d bands of chunks, highest records first: every rank uploads 1/N of a band from pinned "
                                   "host memory, one NCCL all-gather over NVLink completes it on every GPU, a CUDA event behind it releases the band's "
                                   "chunks to the pipelined session (dg_square_push ready_event): no host synchronisation; panels in completion order"
                                   % args.e2e_bands)
                    if world > 1 else "pipelined session (dg_run_square_host) from pinned host memory: upload, tiles and D2H overlap; "
                                      "panels reach the sink in completion order (descending rows)",
                    "in_order_ms_per_step": e2e_inorder_ms,
                    "in_order_note": "dg_load_resident + dg_run_square: same bytes, panels in the reference's output order (no overlap of upload and tiles)",
                    "d2h_bytes_per_step": e2e_d2h_bytes,
                    "result_type": "f64" if not args.is_int else ("u32" if args.u32_results else ("u16 (DG_OPT_RESULT_U16)" if args.u16_results else
                                   "u8 + overflow list for counts >= 255 (DG_OPT_RESULT_U8; a panel with > 16,384 such pairs arrives as u16)"))},
            "gpu_launches": launches, "engine": ENGINE_NAME.get(engine_id, "?"),
            "parity_sampled": "ok", "parity_pairs": head_chk,
            "roofline": roofline, "roofline_pack": roofline_pack, "cpu_baseline": cpu, "clocks": clocks,
            "sustained": sustained, "configs": configs,
        }
        print(json.dumps(line), file=OUT, flush=True)
    d.close()
    return 0


def main():
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner under torchrun) are sent to
    # stderr, and the line goes out through a private duplicate of the original stdout.
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", "--n", dest="n", type=int, default=None,
                    help="record count of the alignment (fixed total work: strong scaling); default 20000*sqrt(N)")
    ap.add_argument("--panel-bytes", type=int, default=None,
                    help="result panel size (default: 6.7e7 results per panel = 128 MiB of uint16 / 256 MiB of uint32)")
    ap.add_argument("--tile-variant", type=int, default=0)
    ap.add_argument("--engine", type=int, default=0, help="DG_OPT_ENGINE: 0 auto, 1 LOP3+POPC, 2 tcgen05 int8, 3 tcgen05 fp4")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--watchdog", type=int, default=900, help="seconds after which a stuck run dumps its stacks and exits")
    ap.add_argument("--trace-e2e", action="store_true", help="after the timed steps, print rank 0's device timeline of one e2e session (DG_TRACE)")
    ap.add_argument("--no-repack-overlap", action="store_true", help="pack every operand plane before the first tile (DG_OPT_REPACK_OVERLAP=0)")
    ap.add_argument("--u32-results", action="store_true", help="keep n / n_high panels as uint32 (default: uint8 + overflow list)")
    ap.add_argument("--u16-results", action="store_true", help="deliver n / n_high panels as uint16 (default: uint8 + overflow list)")
    ap.add_argument("--measure", default="n_high", choices=sorted(OPS_PER_PAIR_SITE),
                    help="default n_high = BASELINE config 2 (the driver's workload); jc69 + --n 100000 = config 5")
    ap.add_argument("--configs", default="1,2,3,4,5", help="BASELINE configs measured next to the headline (\"\" = none)")
    ap.add_argument("--no-configs", dest="configs", action="store_const", const="")
    ap.add_argument("--cfg-steps", type=int, default=3, help="kernel-only steps per extra config")
    ap.add_argument("--cfg-e2e-steps", type=int, default=2, help="e2e steps per extra config")
    ap.add_argument("--stream-records", type=int, default=1000000, help="config 4: streamed records per step (whole job)")
    ap.add_argument("--n5", type=int, default=100000, help="config 5: records of the all-vs-all")
    ap.add_argument("--e2e-bands", type=int, default=4, help="N > 1: all-gathers per e2e step (bands of upload chunks, top down)")
    ap.add_argument("--sustained-s", type=float, default=3.0, help="seconds of back-to-back steps for the `sustained` record (0 = skip)")
    args = ap.parse_args()
    global MEASURE, WORKLOAD
    if args.measure != MEASURE or args.n is not None:
        MEASURE = args.measure
        WORKLOAD = (f"-m {MEASURE} all-vs-all, {args.n or BASE_N:,} x 29,903 nt, 1% N/ambiguity/gaps"
                    + (" (BASELINE config 5)" if MEASURE == "jc69" and args.n == 100000 else ""))
    args.is_int = MEASURE in ("n", "n_high")
    if not args.is_int:
        args.u32_results = True   # f64 panels; the uint16 option only exists for n / n_high
    if args.panel_bytes is None:
        args.panel_bytes = (256 << 20) if args.u32_results else (128 << 20)
    if args.impl == "reference":
        return run_reference(args)
    try:
        return run_ours(args)
    except ParityError as e:
        print(f"bench.py: PARITY MISMATCH against the oracle: {e}", file=sys.stderr, flush=True)
        return 3


if __name__ == "__main__":
    sys.exit(main())
