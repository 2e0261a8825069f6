"""host/measures.hpp (the f64 expressions the C++ host evaluates on DG_RESULT_COUNTS16 panels with the platform's libm, so that
the TSV text of raw / jc69 / k80 / tn93 is byte-identical to the reference's) against the oracle's measures.rs restatement:
the same counts must give the same BITS, NaN / inf / -0.0 cases included.  No GPU: the header is compiled into a small shim with
the host's own flags (-ffp-contract=off) and called through ctypes."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM = r"""
#include "measures.hpp"
extern "C" {
double h_raw(uint32_t n, uint32_t same) { return host::raw_from_counts(n, same); }
double h_jc69(uint32_t n, uint32_t same) { return host::jc69_from_counts(n, same); }
double h_k80(uint32_t same, uint32_t e, uint32_t tv) { return host::k80_from_counts(same, e, tv); }
double h_tn93(uint32_t L, uint32_t d, uint32_t p1, uint32_t p2, const uint32_t* q, const uint32_t* t) {
    return host::tn93_from_counts(L, d, p1, p2, q, t);
}
}
"""


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    cxx = os.environ.get("CXX") or shutil.which("g++") or "/opt/gcc/bin/g++"
    d = tmp_path_factory.mktemp("measures_shim")
    src, so = d / "shim.cpp", d / "libshim.so"
    src.write_text(SHIM)
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "distance_b200", "csrc", "host"), "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    for f, args in (("h_raw", 2), ("h_jc69", 2), ("h_k80", 3)):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_uint32] * args
    L.h_tn93.restype = C.c_double
    L.h_tn93.argtypes = [C.c_uint32] * 4 + [C.c_void_p, C.c_void_p]
    return L


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def bits(x: float) -> int:
    return int(np.float64(x).view(np.uint64))


def pairs():
    from distance_b200 import synth
    rng = np.random.default_rng(77)
    width = 400
    out = []
    for p_ambig, mu in ((0.0, 0.02), (0.05, 0.1), (0.3, 0.5), (0.9, 0.5)):
        a = synth.random_codes(rng, 40, width, p_ambig=p_ambig)
        b = a.copy()
        flip = rng.random(a.shape) < mu
        b[flip] = synth.random_codes(rng, 40, width, p_ambig=p_ambig)[flip]
        out += [(a[i], b[i]) for i in range(40)]
    A, G, Cc, T, N = 136, 72, 40, 24, 240
    const = lambda v: np.full(width, v, np.uint8)
    half = lambda u, v: np.concatenate([np.full(width // 2, u, np.uint8), np.full(width - width // 2, v, np.uint8)])
    out += [(const(N), const(N)), (const(A), const(A)), (const(A), const(T)), (const(A), const(G)), (const(Cc), const(T)),
            (half(A, Cc), half(G, T)), (half(A, N), half(A, N)), (half(A, G), half(G, A)), (half(A, T), half(T, A)),
            (const(A), half(A, T)), (const(192), const(48))]     # R against Y: two partial codes everywhere
    return out


def test_host_epilogues_give_the_oracles_bits(shim, oracle):
    seen = {"nan": 0, "inf": 0, "negzero": 0, "finite": 0}
    for q, t in pairs():
        c = oracle.pair_counts(q, t)
        n, same = c["raw_n"], c["raw_d"] - c["raw_n"]
        want = {"raw": oracle.raw(q, t), "jc69": oracle.jc69(q, t), "k80": oracle.k80(q, t)}
        got = {"raw": shim.h_raw(n, same), "jc69": shim.h_jc69(n, same),
               "k80": shim.h_k80(c["k80_L"] - c["k80_ts"] - c["k80_tv"], c["k80_ts"] + c["k80_tv"], c["k80_tv"])}
        qc = np.ascontiguousarray(oracle.count_bases(q), dtype=np.uint32)
        tc = np.ascontiguousarray(oracle.count_bases(t), dtype=np.uint32)
        want["tn93"] = oracle.tn93(q, t, qc.astype(np.uint64), tc.astype(np.uint64))
        got["tn93"] = shim.h_tn93(c["tn93_L"], c["tn93_d"], c["tn93_P1"], c["tn93_P2"], qc.ctypes.data, tc.ctypes.data)
        for m in want:
            w, g = want[m], got[m]
            if np.isnan(w):
                assert np.isnan(g), (m, c)
                seen["nan"] += 1
            else:
                assert bits(w) == bits(g), (m, w, g, c)
                seen["inf" if np.isinf(w) else "negzero" if (w == 0 and np.signbit(w)) else "finite"] += 1
    assert seen["nan"] and seen["inf"] and seen["negzero"] and seen["finite"] > 300, seen
