#!/usr/bin/env python
"""Device timeline of one pipelined all-vs-all session (DG_TRACE=1 prints it from the library).
usage: DG_TRACE=1 python tools/e2e_trace.py [--n N] [--measure M] [--panels P] [--chunk-mb C]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import distance_b200 as dg
from distance_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--measure", default="n_high")
ap.add_argument("--panels", type=int, default=24)
ap.add_argument("--chunk-mb", type=int, default=0)
a = ap.parse_args()
codes = synth.encode_ascii(synth.make_alignment(a.n, seed=20251018 + 2, ambiguity=True))
pin = api.pinned_array(codes.shape, np.uint8); pin[...] = codes
e = dg.Engine(a.measure, synth.SC2_WIDTH)
if a.measure in ("n", "n_high"):
    e.set_option(api.DG_OPT_RESULT_U16, 1)
e.set_option(api.DG_OPT_PIPE_PANELS, a.panels)
e.set_option(api.DG_OPT_PIPE_CHUNK_BYTES, a.chunk_mb << 20)
for it in range(3):
    t0 = time.perf_counter(); got = e.square_pipelined_discard(pin); t1 = time.perf_counter()
    print(f"iter {it}: pipelined session {1e3*(t1-t0):.2f} ms, {got} pairs", file=sys.stderr, flush=True)
