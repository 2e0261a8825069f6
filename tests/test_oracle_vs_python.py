"""The C oracle against an independent pure-Python transcription of measures.rs (tests/py_measures.py): exact equality
of every integer and of every f64 bit pattern (NaN positions, the sign of zero, infinities) on random pairs over all 17
codes and on the crafted special cases.  No GPU."""
import math
import struct

import numpy as np
import pytest

import py_measures as pm

CODES = np.array([136, 72, 40, 24, 192, 160, 144, 96, 80, 48, 224, 176, 208, 112, 240, 244, 242], dtype=np.uint8)


def same_f64(a: float, b: float) -> bool:
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    return struct.pack("<d", a) == struct.pack("<d", b)


def random_pair(rng, width, p_ambig):
    base = CODES[rng.integers(0, 4, width)]
    q, t = base.copy(), base.copy()
    for s in (q, t):
        m = rng.random(width) < 0.15
        s[m] = CODES[rng.integers(0, 4, int(m.sum()))]
        m = rng.random(width) < p_ambig
        s[m] = CODES[rng.integers(0, 17, int(m.sum()))]
    return q, t


@pytest.mark.parametrize("p_ambig", [0.0, 0.05, 0.5, 1.0])
def test_c_oracle_equals_python_transcription(oracle, p_ambig):
    rng = np.random.default_rng(int(p_ambig * 100) + 7)
    for _ in range(120):
        width = int(rng.integers(1, 160))
        q, t = random_pair(rng, width, p_ambig)
        ql, tl = q.tolist(), t.tolist()
        assert oracle.snp(q, t) == pm.snp(ql, tl)
        assert same_f64(oracle.raw(q, t), pm.raw(ql, tl))
        assert same_f64(oracle.jc69(q, t), pm.jc69(ql, tl))
        assert same_f64(oracle.k80(q, t), pm.k80(ql, tl))
        qc, tc = oracle.count_bases(q), oracle.count_bases(t)
        assert same_f64(oracle.tn93(q, t, qc, tc), pm.tn93(ql, tl, qc, tc))
        # -m n: differences against a consensus of a small alignment that holds both records
        aln = np.stack([q, t] + [random_pair(rng, width, p_ambig)[0] for _ in range(3)])
        cons = oracle.consensus(aln)
        qd, td = oracle.get_differences(q, cons), oracle.get_differences(t, cons)
        got = oracle.snp_consensus(q, t, qd, td)
        assert got == pm.snp_consensus(ql, tl, [int(x) for x in qd], [int(x) for x in td])
        assert got == pm.snp(ql, tl)    # n == n_high


def test_special_cases_agree(oracle):
    A, G, C, T, N = 136, 72, 40, 24, 240
    cases = [
        ([A] * 8, [A] * 8),                      # identical: jc69 / k80 -0.0, tn93 +0.0
        ([N] * 8, [A] * 8),                      # nothing comparable: NaN everywhere
        ([A, A, A, A], [G, G, G, A]),            # p = 3/4: jc69 +inf
        ([A, A, A, A], [C, C, C, C]),            # all transversions: k80 NaN, p = 1
        ([A, A, C, C], [G, G, T, T]),            # all transitions
        ([A, C, G, T] * 4, [C, A, T, G] * 4),
    ]
    for q, t in cases:
        qa, ta = np.array(q, dtype=np.uint8), np.array(t, dtype=np.uint8)
        assert same_f64(oracle.raw(qa, ta), pm.raw(q, t))
        assert same_f64(oracle.jc69(qa, ta), pm.jc69(q, t))
        assert same_f64(oracle.k80(qa, ta), pm.k80(q, t))
        qc, tc = oracle.count_bases(qa), oracle.count_bases(ta)
        assert same_f64(oracle.tn93(qa, ta, qc, tc), pm.tn93(q, t, qc, tc))
    assert math.copysign(1.0, pm.jc69([A] * 8, [A] * 8)) == -1.0 and pm.jc69([A] * 8, [A] * 8) == 0.0
    assert math.copysign(1.0, pm.tn93([A, C, G, T], [A, C, G, T], (1, 1, 1, 1), (1, 1, 1, 1))) == 1.0
    assert pm.jc69([A, A, A, A], [G, G, G, A]) == math.inf
