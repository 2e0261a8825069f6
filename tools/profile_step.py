#!/usr/bin/env python
"""One device-only step of a workload, for ncu / timing experiments.
usage: profile_step.py [--measure M] [--n N] [--variant V] [--steps K] [--mode square|rect]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import distance_b200 as dg
from distance_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--measure", default="n_high")
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--panel-bytes", type=int, default=128 << 20)
ap.add_argument("--u16", action="store_true", help="uint16 result panels (n / n_high), like bench.py")
a = ap.parse_args()
codes = synth.encode_ascii(synth.make_alignment(a.n, seed=20251018 + 2, ambiguity=True))
e = dg.Engine(a.measure, synth.SC2_WIDTH)
e.set_option(api.DG_OPT_PANEL_BYTES, a.panel_bytes)
e.set_option(api.DG_OPT_KEEP_CODES, 1)
if a.u16:
    e.set_option(api.DG_OPT_RESULT_U16, 1)
if a.variant:
    e.set_option(api.DG_OPT_TILE_VARIANT, a.variant)
e.load(0, codes)
e.run_device_only(repack=True)
e.reset_timings()
ms = []
for _ in range(a.steps):
    e.run_device_only(repack=True)
    ms.append(e.timings()["run_ms"])
pairs = a.n * (a.n - 1) // 2
t = e.timings()
print(json.dumps({"measure": a.measure, "n": a.n, "variant": a.variant, "run_ms": ms,
                  "pairs_per_s": pairs / (min(ms) * 1e-3), "pair_sites_per_s": pairs * synth.SC2_WIDTH / (min(ms) * 1e-3),
                  "word_pairs_per_clk_per_sm_at_1965": pairs * 936 / (min(ms) * 1e-3) / 148 / 1.965e9,
                  "pack_ms": t["pack_ms"] / a.steps, "count_ms_sum": t["count_ms"] / a.steps}))
