"""Host-side check of the tensor-engine algebra (distance_b200/csrc/tc_engine.cuh): a numpy restatement
of the int8 plane values and of the per-count plane-pair schedules, compared with the oracle's per-site
categories over ALL 17 x 17 Paradis code pairs.  No GPU needed: this pins the maths the CUDA kernels
implement (minimal-rank bilinear forms + the both-partial repair), not the kernels themselves."""
import itertools

import numpy as np
import pytest

CODES = [136, 72, 40, 24, 192, 160, 144, 96, 80, 48, 224, 176, 208, 112, 240, 244, 242]


def planes(c: int) -> dict:
    """int8 value of every plane for Paradis code c (mirror of plane_word())."""
    A, G, C, T, K = (c >> 7) & 1, (c >> 6) & 1, (c >> 5) & 1, (c >> 4) & 1, (c >> 3) & 1
    e = A + G + C + T - 1
    p = {"UA": 1 - A, "UG": 1 - G, "UC": 1 - C, "UT": 1 - T,
         "VA": 3 * A - e, "VG": 3 * G - e, "VC": 3 * C - e, "VT": 3 * T - e,
         "KA": A & K, "KG": G & K, "KC": C & K, "KT": T & K,
         "PURK": (A | G) & K, "PYRK": (C | T) & K, "K": K,
         "W": (A & K) - (G & K), "Z": (C & K) - (T & K),
         "PURC": (A | G) & (1 - (C | T)), "PYRC": (C | T) & (1 - (A | G))}
    assert all(-128 <= v <= 127 for v in p.values())
    return p


def pp_corr(ma: int, mb: int) -> int:
    """mirror of tc::pp_corr (ma = U-side nibble, mb = V-side nibble; bit3=A .. bit0=T)."""
    nma = ~ma & 15
    cb = bin(mb).count("1")
    got = bin(nma & mb).count("1") * (4 - cb) + bin(nma & ~mb & 15).count("1") * (1 - cb)
    return (0 if (ma & mb) else 3) - got


def is_partial(c):
    return (c & 8) == 0 and (c & 0xF0) != 0xF0


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def test_padding_and_n_like_codes_are_zero_in_every_plane():
    for c in (240, 244, 242):
        assert all(v == 0 for v in planes(c).values()), c


def test_every_count_over_all_code_pairs(oracle):
    for q, t in itertools.product(CODES, CODES):
        pq, pt = planes(q), planes(t)
        cnt = oracle.pair_counts(np.array([q], np.uint8), np.array([t], np.uint8))
        # n / n_high / raw: 3 DIFF = sum_b U_b(q) V_b(t)  (+ repair when both codes are partial)
        acc = sum(pq["U" + b] * pt["V" + b] for b in "AGCT")
        if is_partial(q) and is_partial(t):
            acc += pp_corr(q >> 4, t >> 4)
        assert acc == 3 * cnt["snp"], (q, t, acc)
        assert cnt["raw_n"] == cnt["snp"]
        same = sum(pq["K" + b] * pt["K" + b] for b in "AGCT")
        assert same == cnt["raw_d"] - cnt["raw_n"], (q, t)
        # k80: CS = SAME + ts, X = SAME - ts, tv
        cs = pq["PURK"] * pt["PURK"] + pq["PYRK"] * pt["PYRK"]
        x = pq["W"] * pt["W"] + pq["Z"] * pt["Z"]
        tv = pq["PURC"] * pt["PYRC"] + pq["PYRC"] * pt["PURC"]
        k_same = cnt["k80_L"] - cnt["k80_ts"] - cnt["k80_tv"]
        assert (cs + x) % 2 == 0 and (cs + x) // 2 == k_same and (cs - x) // 2 == cnt["k80_ts"], (q, t)
        assert tv == cnt["k80_tv"], (q, t)
        # tn93: L, PP = SP + P1, YY = SY + P2, WW = SP - P1, ZZ = SY - P2
        L = pq["K"] * pt["K"]
        pp, yy = pq["PURK"] * pt["PURK"], pq["PYRK"] * pt["PYRK"]
        ww, zz = pq["W"] * pt["W"], pq["Z"] * pt["Z"]
        sp, p1, sy, p2 = (pp + ww) // 2, (pp - ww) // 2, (yy + zz) // 2, (yy - zz) // 2
        assert L == cnt["tn93_L"] and L - sp - sy == cnt["tn93_d"], (q, t)
        assert p1 == cnt["tn93_P1"] and p2 == cnt["tn93_P2"], (q, t)


def test_only_both_partial_pairs_need_the_repair(oracle):
    for q, t in itertools.product(CODES, CODES):
        pq, pt = planes(q), planes(t)
        acc = sum(pq["U" + b] * pt["V" + b] for b in "AGCT")
        snp = oracle.pair_counts(np.array([q], np.uint8), np.array([t], np.uint8))["snp"]
        if not (is_partial(q) and is_partial(t)):
            assert acc == 3 * snp, (q, t)


def test_random_alignment_sums(oracle):
    rng = np.random.default_rng(0)
    codes = np.array(CODES, np.uint8)[rng.integers(0, 17, size=(2, 4000))]
    q, t = codes
    P = {k: np.array([planes(int(c))[k] for c in CODES]) for k in planes(136)}
    idx = {c: i for i, c in enumerate(CODES)}
    qi, ti = np.array([idx[int(c)] for c in q]), np.array([idx[int(c)] for c in t])
    acc = sum(int((P["U" + b][qi] * P["V" + b][ti]).sum()) for b in "AGCT")
    acc += sum(pp_corr(int(a) >> 4, int(b) >> 4) for a, b in zip(q, t) if is_partial(int(a)) and is_partial(int(b)))
    assert acc % 3 == 0 and acc // 3 == oracle.pair_counts(q, t)["snp"]


E2M1 = {0: 0x0, 1: 0x2, 2: 0x4, 3: 0x5, -1: 0xA, -2: 0xC}


def test_fp4_nibble_formulas_match_the_plane_values():
    """Mirror of nib_v / nib_pm / plane_nib8 (bit formulas on 0/1 lanes): the nibble must be the E2M1 code of
    the int8 plane value for every code and plane."""
    for c in CODES:
        A, G, C, T, K = (c >> 7) & 1, (c >> 6) & 1, (c >> 5) & 1, (c >> 4) & 1, (c >> 3) & 1
        E = A + G + C + T - 1
        e0, e1 = E & 1, (E >> 1) & 1
        x, y, nz = e1 & (1 - e0), e0 & (1 - e1), e0 | e1

        def nib_v(b):
            b3 = nz & (1 - b)
            b2 = (b & (1 - e1)) | ((1 - b) & x)
            b1 = (b & x) | ((1 - b) & y)
            b0 = b & (1 - nz)
            return b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)

        def nib_pm(pos, neg):
            return ((pos | neg) << 1) | (neg << 3)

        p = planes(c)
        got = {"UA": (A ^ 1) << 1, "UG": (G ^ 1) << 1, "UC": (C ^ 1) << 1, "UT": (T ^ 1) << 1,
               "VA": nib_v(A), "VG": nib_v(G), "VC": nib_v(C), "VT": nib_v(T),
               "KA": (A & K) << 1, "KG": (G & K) << 1, "KC": (C & K) << 1, "KT": (T & K) << 1,
               "PURK": ((A | G) & K) << 1, "PYRK": ((C | T) & K) << 1, "K": K << 1,
               "W": nib_pm(A & K, G & K), "Z": nib_pm(C & K, T & K),
               "PURC": ((A | G) & (1 - (C | T))) << 1, "PYRC": ((C | T) & (1 - (A | G))) << 1}
        for k, v in p.items():
            assert got[k] == E2M1[v], (c, k, v, got[k])
