#!/usr/bin/env python
"""tcgen05 engine (DG_OPT_ENGINE=2) against the LOP3 engine on the same inputs; then a timing run."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import distance_b200 as dg
from distance_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--big", type=int, default=0, help="records of the 29,903-nt timing run (0 = skip)")
ap.add_argument("--measure", default="n_high")
ap.add_argument("--variant", type=int, default=0)
a = ap.parse_args()

def run(codes, engine, mode="square", codes_b=None, panel=None):
    e = dg.Engine(a.measure, codes.shape[1])
    e.set_option(api.DG_OPT_ENGINE, engine)
    if engine == 2 and a.variant:
        e.set_option(api.DG_OPT_TILE_VARIANT, a.variant)
    if panel:
        e.set_option(api.DG_OPT_PANEL_BYTES, panel)
    e.load(0, codes)
    if codes_b is not None:
        e.load(1, codes_b)
    out = e.run_square() if mode == "square" else e.run_rect()
    e.close()
    return out

ok = True
rng = np.random.default_rng(1)
for (n, w, amb) in [(2, 5, 0.3), (130, 100, 0.3), (300, 1000, 0.5), (700, 333, 0.05), (1000, 4097, 0.9)]:
    codes = synth.random_codes(rng, n, w, p_ambig=amb)
    ref = run(codes, 1)
    got = run(codes, 2, panel=max(4096, 512 * n * 4))
    bad = int((ref != got).sum())
    print(f"square n={n} w={w} amb={amb}: mismatches {bad} / {ref.size}", flush=True)
    if bad:
        i = np.flatnonzero(ref != got)[:5]
        print("   first:", i, ref[i], got[i])
        ok = False
ca, cb = synth.random_codes(rng, 150, 777, 0.2), synth.random_codes(rng, 519, 777, 0.2)
ref, got = run(ca, 1, "rect", cb), run(ca, 2, "rect", cb)
print("rect 150x519: mismatches", int((ref != got).sum()), flush=True)
ok &= bool((ref == got).all())
print("PARITY", "OK" if ok else "FAILED", flush=True)

if a.big and ok:
    codes = synth.encode_ascii(synth.make_alignment(a.big, seed=20251018 + 2, ambiguity=True))
    res = {}
    for eng in (1, 2):
        e = dg.Engine(a.measure, synth.SC2_WIDTH)
        e.set_option(api.DG_OPT_ENGINE, eng)
        if eng == 2 and a.variant:
            e.set_option(api.DG_OPT_TILE_VARIANT, a.variant)
        e.load(0, codes)
        e.run_device_only()
        e.reset_timings()
        ms = []
        for _ in range(3):
            e.run_device_only()
            ms.append(e.timings()["run_ms"])
        if eng == 1:
            ref = e.run_square()
        else:
            got = e.run_square()
        res[eng] = min(ms)
        e.close()
    pairs = a.big * (a.big - 1) // 2
    print(json.dumps({"n": a.big, "lop3_ms": res[1], "tc_ms": res[2], "lop3_pairs_per_s": pairs / res[1] * 1e3,
                      "tc_pairs_per_s": pairs / res[2] * 1e3, "mismatches": int((ref != got).sum())}))
