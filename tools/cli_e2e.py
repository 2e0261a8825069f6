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
a = ap.parse_args()
cli = os.path.join(ROOT, "distance_b200", "_bin", "distance")
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
for sink in ("/dev/null", tsv):
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
