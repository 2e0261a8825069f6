#!/usr/bin/env python
"""End-to-end step time of the pipelined session (dg_run_square_host) vs the in-order path
(dg_load_resident + dg_run_square) over panel / chunk settings.  usage: e2e_sweep.py [--n N] [--measure M]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import distance_b200 as dg
from distance_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--measure", default="n_high")
ap.add_argument("--reps", type=int, default=6)
a = ap.parse_args()
codes = synth.encode_ascii(synth.make_alignment(a.n, seed=20251018 + 2, ambiguity=True))
pinned = api.pinned_array(codes.shape, np.uint8)
pinned[...] = codes
pairs = a.n * (a.n - 1) // 2
is_int = a.measure in ("n", "n_high")


def timed(f, reps):
    f(); f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); t.append(1e3 * (time.perf_counter() - t0))
    return min(t), float(np.median(t))


with dg.Engine(a.measure, synth.SC2_WIDTH) as e:
    if is_int:
        e.set_option(api.DG_OPT_RESULT_U16, 1)
    e.set_option(api.DG_OPT_PANEL_BYTES, (128 << 20) if is_int else (256 << 20))

    def classic():
        e.load(0, pinned)
        assert e.run_discard() == pairs
    best, med = timed(classic, a.reps)
    print(json.dumps({"path": "in-order (load + run_square)", "ms_min": best, "ms_median": med, "pairs_per_s": pairs / med * 1e3}), flush=True)
    for panels in (12, 24, 32, 48):
        for chunk_mb in (0, 16, 32):
            e.set_option(api.DG_OPT_PIPE_PANELS, panels)
            e.set_option(api.DG_OPT_PIPE_CHUNK_BYTES, chunk_mb << 20)

            def piped():
                assert e.square_pipelined_discard(pinned) == pairs
            best, med = timed(piped, a.reps)
            print(json.dumps({"path": "pipelined", "panels": panels, "chunk_mb": chunk_mb, "ms_min": best, "ms_median": med,
                              "pairs_per_s": pairs / med * 1e3}), flush=True)
