"""Deterministic synthetic alignments for tests and bench.py.

SARS-CoV-2-like: width 29,903, root with base composition A .299 / C .184 / G .196 / T .321, every
record = root + substitutions at rate mu (ts:tv 2:1).  The "1% N / ambiguity + gaps" mix is 0.8% N in
runs (geometric, mean 200, amplicon-dropout like), 0.1% IUPAC 2-/3-fold codes, 0.1% '-' incl. 5'/3'
terminal runs.  PRNG: numpy PCG64 seeded with `seed` (bench uses 20251018 + config number).
Returns upper-case ASCII (n x width uint8); `encode_ascii` maps it to Paradis codes with the table
of the reference's encoding.rs:4-41.
"""
from __future__ import annotations

import numpy as np

SC2_WIDTH = 29903
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_TRANSITION = {ord("A"): ord("G"), ord("G"): ord("A"), ord("C"): ord("T"), ord("T"): ord("C")}
_IUPAC_PARTIAL = np.frombuffer(b"RYMWSKVHDB", dtype=np.uint8)


def ascii_lut() -> np.ndarray:
    lut = np.zeros(256, dtype=np.uint8)
    table = {"A": 136, "G": 72, "C": 40, "T": 24, "R": 192, "M": 160, "W": 144, "S": 96, "K": 80,
             "Y": 48, "V": 224, "H": 176, "D": 208, "B": 112, "N": 240}
    for ch, v in table.items():
        lut[ord(ch)] = v
        lut[ord(ch.lower())] = v
    lut[ord("-")] = 244
    lut[ord("?")] = 242
    return lut


def encode_ascii(ascii_codes: np.ndarray) -> np.ndarray:
    return ascii_lut()[ascii_codes]


def make_root(width: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng([seed, 0xC0FFEE])
    return _BASES[rng.choice(4, size=width, p=[0.299, 0.184, 0.196, 0.321])]


def make_alignment(n: int, width: int = SC2_WIDTH, seed: int = 20251018, mu: float = 1e-3,
                   ambiguity: bool = False, root: np.ndarray | None = None,
                   n_rate: float = 0.008, iupac_rate: float = 0.001, gap_rate: float = 0.001) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if root is None:
        root = make_root(width, seed)
    aln = np.tile(root, (n, 1))
    # substitutions: ts with prob 2/3, else one of the two transversions
    k = rng.binomial(n * width, mu) if n * width else 0
    if k:
        r = rng.integers(0, n, size=k)
        c = rng.integers(0, width, size=k)
        old = aln[r, c]
        ts = np.vectorize(_TRANSITION.get, otypes=[np.uint8])(old)
        is_ts = rng.random(k) < (2.0 / 3.0)
        pur = (old == ord("A")) | (old == ord("G"))
        tv_choice = rng.integers(0, 2, size=k)
        tv = np.where(pur, np.where(tv_choice == 0, ord("C"), ord("T")),
                      np.where(tv_choice == 0, ord("A"), ord("G"))).astype(np.uint8)
        aln[r, c] = np.where(is_ts, ts, tv)
    if ambiguity:
        # N runs
        runs = rng.poisson(n * width * n_rate / 200.0)
        rr = rng.integers(0, n, size=runs)
        start = rng.integers(0, width, size=runs)
        length = rng.geometric(1.0 / 200.0, size=runs)
        for a, s, l in zip(rr.tolist(), start.tolist(), length.tolist()):
            aln[a, s:s + l] = ord("N")
        # partial IUPAC codes, uniformly
        k = rng.binomial(n * width, iupac_rate)
        aln[rng.integers(0, n, size=k), rng.integers(0, width, size=k)] = _IUPAC_PARTIAL[rng.integers(0, 10, size=k)]
        # gaps: terminal runs on half the records, the rest scattered
        lead = rng.geometric(1.0 / 10.0, size=n) * (rng.random(n) < 0.5)
        trail = rng.geometric(1.0 / 10.0, size=n) * (rng.random(n) < 0.5)
        for a in range(n):
            if lead[a]:
                aln[a, :lead[a]] = ord("-")
            if trail[a]:
                aln[a, width - trail[a]:] = ord("-")
        k = rng.binomial(n * width, max(gap_rate - 10.0 / width, 0.0))
        aln[rng.integers(0, n, size=k), rng.integers(0, width, size=k)] = ord("-")
    return aln


def random_codes(rng: np.random.Generator, n: int, width: int, p_ambig: float = 0.15) -> np.ndarray:
    """Adversarial test input: every one of the 17 Paradis codes at high frequency."""
    codes = np.array([136, 72, 40, 24, 192, 160, 144, 96, 80, 48, 224, 176, 208, 112, 240, 244, 242], np.uint8)
    p = np.array([(1 - p_ambig) / 4] * 4 + [p_ambig / 13] * 13)
    return codes[rng.choice(17, size=(n, width), p=p)]


def ids(n: int, prefix: str = "s") -> list[str]:
    return [f"{prefix}{i:06d}" for i in range(n)]


def write_fasta(path: str, ascii_codes: np.ndarray, names: list[str]) -> None:
    with open(path, "wb") as f:
        for name, row in zip(names, ascii_codes):
            f.write(b">" + name.encode() + b"\n" + row.tobytes() + b"\n")
