#!/usr/bin/env python
"""End to end through the `distance` command line (process start, FASTA parse, GPU, ordered TSV write), the figure
north_star asks for next to the kernel-only one.  Writes the synthetic FASTA to /dev/shm (or --dir), runs the CLI with
the TSV going to /dev/null and to a file, and times the oracle-side restatement of the reference's own e2e on a sample.
usage: cli_e2e.py [--n 20000] [--measure n_high] [--dir /dev/shm]"""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distance_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--measure", default="n_high")
ap.add_argument("--dir", default="/dev/shm")
ap.add_argument("--seed", type=int, default=20251018 + 2)
ap.add_argument("--null-only", action="store_true", help="TSV to /dev/null only (large runs: the text would not fit a tmpfs)")
ap.add_argument("--stream", type=int, default=0,
                help="config 4 instead: -m k80 -i <1,000 resident records> -s <this many streamed records> (a pool of 20,000 records repeated)")
a = ap.parse_args()
cli = os.path.join(ROOT, "distance_b200", "_bin", "distance")
if a.stream:
    import numpy as np
    root = synth.make_root(synth.SC2_WIDTH, 20251018 + 4)
    res = synth.make_alignment(1000, seed=20251018 + 4, ambiguity=True, root=root)
    pool = synth.make_alignment(20000, seed=20251018 + 44, ambiguity=True, root=root)
    fr, fs = os.path.join(a.dir, "dg_cfg4_resident.fasta"), os.path.join(a.dir, "dg_cfg4_stream.fasta")
    with open(fr, "wb") as f:
        for i, r in enumerate(res):
            f.write(b">r%06d\n" % i); f.write(r.tobytes()); f.write(b"\n")
    t0 = time.time()
    with open(fs, "wb") as f:
        k = 0
        while k < a.stream:
            for r in pool[:min(len(pool), a.stream - k)]:
                f.write(b">q%07d\n" % k); f.write(r.tobytes()); f.write(b"\n"); k += 1
    gen_s = time.time() - t0
    pairs = 1000 * a.stream
    out = {"config": "cfg4 through the CLI: -m k80 -i 1,000 resident -s %d streamed" % a.stream, "pairs": pairs,
           "stream_fasta_bytes": os.path.getsize(fs), "cores": os.cpu_count(), "fasta_write_s": gen_s, "runs": []}
    for rep in range(2):
        t0 = time.time()
        with open("/dev/null", "wb") as so:
            p = subprocess.run([cli, "-m", "k80", "-i", fr, "-s", fs], stdout=so, stderr=subprocess.PIPE, env=dict(os.environ, DG_TRACE="1"))
        wall = time.time() - t0
        out["runs"].append({"tsv_to": "/dev/null", "rc": p.returncode, "wall_s": wall, "pairs_per_s": pairs / wall,
                            "stream_gb_per_s": os.path.getsize(fs) / wall / 1e9,
                            "phases": [l for l in p.stderr.decode().splitlines() if l.startswith("[distance]")]})
    os.unlink(fr); os.unlink(fs)
    print(json.dumps(out))
    sys.exit(0)
fa = os.path.join(a.dir, f"dg_cli_e2e_{a.n}.fasta")
t0 = time.time()
asc = synth.make_alignment(a.n, seed=a.seed, ambiguity=True)
with open(fa, "wb") as f:
    for i, r in enumerate(asc):
        f.write(b">s%06d\n" % i); f.write(r.tobytes()); f.write(b"\n")
gen_s = time.time() - t0
pairs = a.n * (a.n - 1) // 2
out = {"n": a.n, "measure": a.measure, "pairs": pairs, "fasta_bytes": os.path.getsize(fa), "cores": os.cpu_count(), "runs": []}
tsv = os.path.join(a.dir, "dg_cli_e2e.tsv")
for sink in (("/dev/null",) if a.null_only else ("/dev/null", tsv)):
    for rep in range(2):
        env = dict(os.environ, DG_TRACE="1")
        t0 = time.time()
        with open(sink, "wb") as so:
            p = subprocess.run([cli, "-m", a.measure, fa], stdout=so, stderr=subprocess.PIPE, env=env)
        wall = time.time() - t0
        phases = [l for l in p.stderr.decode().splitlines() if l.startswith("[distance]") or l.startswith("[load_fasta]")]
        out["runs"].append({"tsv_to": "file in " + a.dir if sink == tsv else sink, "rc": p.returncode, "wall_s": wall,
                            "pairs_per_s": pairs / wall, "tsv_bytes": os.path.getsize(sink) if sink == tsv else None, "phases": phases})
    if sink == tsv and os.path.exists(tsv):
        os.unlink(tsv)
os.unlink(fa)
print(json.dumps(out))
