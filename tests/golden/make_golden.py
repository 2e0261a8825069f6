#!/usr/bin/env python
"""Regenerates the committed TSV fixtures of tests/golden/ from the CPU oracle (the C restatement of the reference's
measures.rs / fastaio.rs / lib.rs, itself pinned to the reference's own golden vectors in tests/test_oracle_golden.py).

    python tests/golden/make_golden.py          # rewrites golden_a.fasta, golden_b.fasta and golden_<mode>_<measure>.tsv

Inputs: two small alignments (16 and 9 records, 97 sites) that exercise every code of encoding.rs:7-38 -- the four
bases in both cases, all ten partial ambiguity codes, N, '-' and '?' -- plus an identical pair (jc69 / k80 print
-0.000000000000, tn93 0.000000000000), a pair with no comparable site (NaN), a saturated pair (NaN) and a pair at exactly p = 3/4 (jc69 inf).
Modes: square = `distance -m M a` (lib.rs:390-399), rect = `distance -m M a b` (lib.rs:401-409),
stream = `distance -m M -i a -s b` (lib.rs:269-365; tn93 counts raw upper-case bases of streamed records only,
fastaio.rs:139-142)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
MEASURES = ["n", "n_high", "raw", "jc69", "k80", "tn93"]
WIDTH = 97


def alignments():
    rng = np.random.default_rng(20251018)
    letters = np.frombuffer(b"ACGTACGTACGTACGTACGTacgtRYMWSKVHDBNrykmn-?", dtype=np.uint8)
    root = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, WIDTH)]

    def mutate(p_sub, p_any):
        s = root.copy()
        m = rng.random(WIDTH) < p_sub
        s[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        m = rng.random(WIDTH) < p_any
        s[m] = letters[rng.integers(0, letters.size, int(m.sum()))]
        return s

    a = [mutate(0.05, 0.08) for _ in range(10)]
    a.append(a[3].copy())                                              # identical pair: -0.0 / 0.0
    a.append(np.full(WIDTH, ord("N"), dtype=np.uint8))                 # nothing comparable: NaN
    far = root.copy()                                                  # every site a transversion: saturated
    far = np.array([{65: 67, 67: 65, 71: 84, 84: 71}[int(c)] for c in far], dtype=np.uint8)
    a.append(far)
    a.append(np.frombuffer((b"RYMWSKVHDBNrymwskvhdbn-?" * 5)[:WIDTH], dtype=np.uint8).copy())
    a.append(root.copy())                                              # with the next record: 72 of 96 compared sites differ,
    q = root.copy()                                                    # p = 3/4 exactly -> jc69 prints inf (measures.rs:76)
    q[:72] = np.array([{65: 71, 71: 65, 67: 84, 84: 67}[int(c)] for c in root[:72]], dtype=np.uint8)
    q[96] = ord("N")
    a.append(q)
    b = [mutate(0.08, 0.10) for _ in range(8)] + [a[0].copy()]
    ids_a = [f"a{i:02d}|sample/{i}" for i in range(len(a))]
    ids_b = [f"b{i:02d}" for i in range(len(b))]
    return ids_a, np.stack(a), ids_b, np.stack(b)


def fasta(ids, rows):
    return b"".join(b">" + i.encode() + b" a description\n" + r.tobytes() + b"\n" for i, r in zip(ids, rows))


def golden_texts():
    from distance_b200 import synth
    from oracle import oracle as orc
    orc.build()
    ids_a, asc_a, ids_b, asc_b = alignments()
    ca, cb = synth.encode_ascii(asc_a), synth.encode_ascii(asc_b)
    out = {"golden_a.fasta": fasta(ids_a, asc_a), "golden_b.fasta": fasta(ids_b, asc_b)}
    for m in MEASURES:
        is_int = m in ("n", "n_high")
        a = orc.Alignment(ca)
        orc.prepare(m, [a])
        v, _ = orc.run(m, "square", a)
        out[f"golden_square_{m}.tsv"] = orc.tsv(ids_a, ids_a, "square", v, is_int).encode()
        a, b = orc.Alignment(ca), orc.Alignment(cb)
        orc.prepare(m, [a, b])
        v, _ = orc.run(m, "rect", a, b)
        out[f"golden_rect_{m}.tsv"] = orc.tsv(ids_a, ids_b, "rect", v, is_int).encode()
        # streamed records: tn93 counts raw upper-case A/T/G/C only (fastaio.rs:139-142)
        acgt_b = None
        if m == "tn93":
            acgt_b = np.stack([orc.encode_count_bases(r.tobytes())[1] for r in asc_b]).astype(np.uint64)
        a, b = orc.Alignment(ca), orc.Alignment(cb, acgt_b)
        orc.prepare(m, [a, b], consensus_from=[a])
        v, _ = orc.run(m, "stream", a, b)
        out[f"golden_stream_{m}.tsv"] = orc.tsv(ids_a, ids_b, "stream", v, is_int).encode()
    return out


if __name__ == "__main__":
    for name, data in golden_texts().items():
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(data)
        print(name, len(data), "bytes")
