#!/usr/bin/env python
"""One BASELINE config, kernel-only, a few steps: the command `ncu` wraps for launch lists and captures.

usage: profile_cfg.py --cfg 3 [--n N] [--steps K] [--engine E] [--parts P]
  cfg 1 raw square n=1000 | 2 n_high square n=20000 (u16) | 3 tn93 rect n x n (10000) |
  cfg 4 k80 stream: 1000 resident, --n streamed records (default 32768) | 5 jc69 square, part 0 of --parts
Prints one JSON line with the device time per step (CUDA events of the library)."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import distance_b200 as dg
from distance_b200 import api, synth

W = synth.SC2_WIDTH


def pinned(a):
    p = api.pinned_array(a.shape, np.uint8)
    p[...] = a
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--measure", default=None)
    ap.add_argument("--width", type=int, default=W)
    args = ap.parse_args()
    dg.load_library()
    cfg = args.cfg
    measure = args.measure or {1: "raw", 2: "n_high", 3: "tn93", 4: "k80", 5: "jc69"}[cfg]
    n = args.n or {1: 1000, 2: 20000, 3: 10000, 4: 32768, 5: 100000}[cfg]
    width = args.width
    seed = 20251018 + cfg
    out = {"cfg": cfg, "measure": measure, "n": n, "width": width}
    e = dg.Engine(measure, width)
    if args.engine:
        e.set_option(api.DG_OPT_ENGINE, args.engine)
    e.set_option(api.DG_OPT_KEEP_CODES, 1)
    if cfg == 4:
        root = synth.make_root(width, seed)
        res = synth.encode_ascii(synth.make_alignment(1000, width=width, seed=seed, ambiguity=True, root=root))
        batch = 4096
        pool = pinned(synth.encode_ascii(synth.make_alignment(batch, width=width, seed=seed + 40, ambiguity=True, root=root)))
        e.load(0, pinned(res))
        state = {"n": 0}

        def sink(user, pp):
            state["n"] += int(pp.contents.n_results)
            return 0
        cb = api.SINK_FN(sink)
        L = e.L
        for it in range(args.warmup + 1):
            e._check(L.dg_stream_begin(e.h, cb, None, batch))
            e.reset_timings()
            t0 = time.time()
            for _ in range(max(1, n // batch)):
                e._check(L.dg_stream_push(e.h, C.c_void_p(pool.ctypes.data), batch, api.DG_INPUT_PARADIS, None))
            e._check(L.dg_stream_end(e.h))
            wall = 1e3 * (time.time() - t0)
        tm = e.timings()
        out.update({"pairs": state["n"] // (args.warmup + 1), "count_ms": tm["count_ms"], "pack_ms": tm["pack_ms"], "wall_ms": wall,
                    "engine": int(tm["engine"])})
    else:
        mode = api.DG_MODE_RECT if cfg == 3 else api.DG_MODE_SQUARE
        if cfg == 3:
            root = synth.make_root(width, seed)
            a = synth.encode_ascii(synth.make_alignment(n, width=width, seed=seed, ambiguity=True, root=root))
            b = synth.encode_ascii(synth.make_alignment(n, width=width, seed=seed + 30, ambiguity=True, root=root))
        else:
            a = synth.encode_ascii(synth.make_alignment(n, width=width, seed=seed, ambiguity=True))
            b = None
        if cfg == 2:
            e.set_option(api.DG_OPT_RESULT_U16, 1)
            e.set_option(api.DG_OPT_PANEL_BYTES, 128 << 20)
        e.load(0, pinned(a))
        if b is not None:
            e.load(1, pinned(b))
        part, parts = (0, args.parts) if cfg == 5 else (0, 1)
        plan = e.plan(mode)
        from distance_b200 import dist
        pairs = sum(p[2] for p in dist.my_panels(plan, part, parts))
        for _ in range(args.warmup):
            e.run_device_only(mode, part, parts, repack=True)
        ms = []
        for _ in range(args.steps):
            e.reset_timings()
            e.run_device_only(mode, part, parts, repack=True)
            tm = e.timings()
            ms.append(tm["run_ms"])
        out.update({"pairs": pairs, "panels": len(plan), "run_ms": ms, "count_ms": tm["count_ms"], "pack_ms": tm["pack_ms"],
                    "engine": int(tm["engine"]), "count_launches": int(tm["count_launches"])})
    e.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
