"""A second, independent restatement of the reference's per-pair measures in pure Python (slow; small cases only),
written straight from /root/reference/src/measures.rs without looking at oracle/distance_oracle.c.  The C oracle and this
transcription must agree exactly (tests/test_oracle_vs_python.py): two independent readings of the Rust source guard
against a shared misreading (operator precedence, branch order, f64 expression order).

Rust semantics kept on purpose:
  * `q & t < 16` parses as `(q & t) < 16` and `q & 8 == 8` as `(q & 8) == 8` (measures.rs:17, 60, 62);
  * usize / usize as f64 divisions follow IEEE 754 (x / 0 = inf or NaN, never an exception);
  * f64::ln of 0 is -inf, of a negative number NaN; f64::sqrt of a negative number NaN;
  * `query.differences[start..].binary_search(idx)` returns a position RELATIVE to the slice, and `start = pos` stores it
    unchanged (measures.rs:40-43)."""
import bisect
import math


def _div(a: float, b: float) -> float:
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * (math.copysign(1.0, b))
    return a / b


def _ln(x: float) -> float:
    if x != x or x < 0.0:
        return math.nan
    if x == 0.0:
        return -math.inf
    return math.log(x)


def _sqrt(x: float) -> float:
    if x != x or x < 0.0:
        return math.nan
    return math.sqrt(x)


def snp(q, t) -> int:  # measures.rs:14-23
    d = 0
    for i in range(len(t)):
        if (q[i] & t[i]) < 16:
            d += 1
    return d


def snp_consensus(q, t, q_diffs, t_diffs) -> int:  # measures.rs:28-53
    d = 0
    for idx in q_diffs:
        if (q[idx] & t[idx]) < 16:
            d += 1
    start = 0
    for idx in t_diffs:
        sub = q_diffs[start:]
        pos = bisect.bisect_left(sub, idx)
        if pos < len(sub) and sub[pos] == idx:
            start = pos
            continue
        if (q[idx] & t[idx]) < 16:
            d += 1
    return d


def raw(q, t) -> float:  # measures.rs:56-69
    d = n = 0
    for i in range(len(t)):
        if (q[i] & 8) == 8 and q[i] == t[i]:
            d += 1
        elif (q[i] & t[i]) < 16:
            d += 1
            n += 1
    return _div(float(n), float(d))


def jc69(q, t) -> float:  # measures.rs:72-77
    p = raw(q, t)
    return -0.75 * _ln(1.0 - (4.0 / 3.0) * p)


def k80(q, t) -> float:  # measures.rs:80-113
    count_l = ts = tv = 0
    for i in range(len(t)):
        if (q[i] & 8) == 8 and q[i] == t[i]:
            count_l += 1
        elif (q[i] & t[i]) < 16:
            if (q[i] & 55) == 0 and (t[i] & 55) == 0:
                ts += 1
                count_l += 1
            elif (q[i] & 199) == 0 and (t[i] & 199) == 0:
                ts += 1
                count_l += 1
            elif ((q[i] & 55) == 0 and (t[i] & 199) == 0) or ((q[i] & 199) == 0 and (t[i] & 55) == 0):
                tv += 1
                count_l += 1
    p = _div(float(ts), float(count_l))
    qq = _div(float(tv), float(count_l))
    return -0.5 * _ln((1.0 - 2.0 * p - qq) * _sqrt(1.0 - 2.0 * qq))


def tn93(q, t, qc, tc) -> float:  # measures.rs:116-193; qc / tc = (count_A, count_T, count_G, count_C)
    qa, qt, qg, qcc = (int(x) for x in qc)
    ta, tt, tg, tcc = (int(x) for x in tc)
    big_l = float(qa + qt + qg + qcc + ta + tt + tg + tcc)
    g_a = _div(float(ta) + float(qa), big_l)
    g_c = _div(float(tcc) + float(qcc), big_l)
    g_g = _div(float(tg) + float(qg), big_l)
    g_t = _div(float(tt) + float(qt), big_l)
    g_r = _div(float(ta) + float(qa) + float(tg) + float(qg), big_l)
    g_y = _div(float(tcc) + float(qcc) + float(tt) + float(qt), big_l)
    k1 = _div(2.0 * g_a * g_g, g_r)
    k2 = _div(2.0 * g_t * g_c, g_y)
    k3 = 2.0 * (g_r * g_y - _div(g_a * g_g * g_y, g_r) - _div(g_t * g_c * g_r, g_y))
    p1 = p2 = cd = cl = 0
    for i in range(len(t)):
        if (q[i] & 8) == 8 and q[i] == t[i]:
            cl += 1
        elif (q[i] & t[i]) < 16 and (q[i] & 8) == 8 and (t[i] & 8) == 8:
            cd += 1
            cl += 1
            if (q[i] | t[i]) == 200:
                p1 += 1
            elif (q[i] | t[i]) == 56:
                p2 += 1
    big_p1 = _div(float(p1), float(cl))
    big_p2 = _div(float(p2), float(cl))
    big_q = _div(float(cd - (p1 + p2)), float(cl))
    w1 = 1.0 - _div(big_p1, k1) - _div(big_q, 2.0 * g_r)
    w2 = 1.0 - _div(big_p2, k2) - _div(big_q, 2.0 * g_y)
    w3 = 1.0 - _div(big_q, 2.0 * g_r * g_y)
    d = -k1 * _ln(w1) - k2 * _ln(w2) - k3 * _ln(w3)
    if d == 0.0:
        d = 0.0
    return d
