import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import distance_b200 as dg
from distance_b200 import api, synth
n = 20000
codes = synth.encode_ascii(synth.make_alignment(n, seed=20251018 + 2, ambiguity=True))
pin = api.pinned_array(codes.shape, np.uint8); pin[...] = codes
e = dg.Engine("n_high", synth.SC2_WIDTH)
for it in range(3):
    t0 = time.time(); e.load(0, pin); t1 = time.time(); got = e.run_discard(); t2 = time.time()
    print(f"iter {it}: load {1e3*(t1-t0):.1f} ms, run {1e3*(t2-t1):.1f} ms, timings {e.timings()}", flush=True)
