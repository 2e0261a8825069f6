"""ctypes loader for the CPU oracle (oracle/distance_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/ (incl. tests/golden/make_golden.py),
__graft_entry__.smoke() and the timed CPU baselines (the cpu_baseline /
`--impl reference` legs of bench.py and the same bounded-sample baseline of the
measurement script tools/run_configs.py).  Nothing under distance_b200/ imports
this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MEASURES = {"n": 0, "n_high": 1, "raw": 2, "jc69": 3, "k80": 4, "tn93": 5}
INT_MEASURES = ("n", "n_high")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "distance_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Aln(C.Structure):
    _fields_ = [
        ("seqs", C.c_void_p),
        ("n", C.c_uint64),
        ("w", C.c_uint64),
        ("acgt", C.c_void_p),
        ("diff_off", C.c_void_p),
        ("diff_idx", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, u64p = C.c_void_p, C.c_void_p
        L.or_encoding_array.argtypes = [u8p]
        L.or_encode.argtypes = [u8p, C.c_uint64, u8p]
        L.or_encode.restype = C.c_int64
        L.or_count_bases.argtypes = [u8p, C.c_uint64, u64p]
        L.or_encode_count_bases.argtypes = [u8p, C.c_uint64, u8p, u64p]
        L.or_encode_count_bases.restype = C.c_int64
        L.or_get_differences.argtypes = [u8p, u8p, C.c_uint64, u64p]
        L.or_get_differences.restype = C.c_uint64
        L.or_consensus.argtypes = [u8p, C.c_uint64, C.c_uint64, u8p]
        L.or_snp.argtypes = [u8p, u8p, C.c_uint64]
        L.or_snp.restype = C.c_int64
        L.or_snp_consensus.argtypes = [u8p, u8p, u64p, C.c_uint64, u64p, C.c_uint64]
        L.or_snp_consensus.restype = C.c_int64
        for name in ("or_raw", "or_jc69", "or_k80"):
            f = getattr(L, name)
            f.argtypes = [u8p, u8p, C.c_uint64]
            f.restype = C.c_double
        L.or_tn93.argtypes = [u8p, u8p, C.c_uint64, u64p, u64p]
        L.or_tn93.restype = C.c_double
        L.or_raw_counts.argtypes = [u8p, u8p, C.c_uint64, u64p, u64p]
        L.or_k80_counts.argtypes = [u8p, u8p, C.c_uint64, u64p, u64p, u64p]
        L.or_tn93_counts.argtypes = [u8p, u8p, C.c_uint64, u64p, u64p, u64p, u64p]
        L.or_format_float12.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
        L.or_format_float12.restype = C.c_int
        L.or_run.argtypes = [C.c_int, C.c_int, C.POINTER(_Aln), C.POINTER(_Aln), C.c_void_p,
                             C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.or_run.restype = C.c_uint64
        L.or_bench.argtypes = [C.c_int, C.c_int, C.POINTER(_Aln), C.POINTER(_Aln), C.c_uint64,
                               C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.or_bench.restype = C.c_uint64
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


# ---- per-record functions -------------------------------------------------

def encoding_array() -> np.ndarray:
    a = np.zeros(256, dtype=np.uint8)
    lib().or_encoding_array(_p(a))
    return a


def encode(ascii_seq) -> np.ndarray:
    """fastaio.rs:101-118. Raises ValueError(index) on an invalid character."""
    s = _u8(ascii_seq)
    out = np.zeros(s.shape[0], dtype=np.uint8)
    bad = lib().or_encode(_p(s), s.shape[0], _p(out))
    if bad >= 0:
        raise ValueError(int(bad))
    return out


def count_bases(seq) -> np.ndarray:
    s = _u8(seq)
    out = np.zeros(4, dtype=np.uint64)
    lib().or_count_bases(_p(s), s.shape[0], _p(out))
    return out


def encode_count_bases(ascii_seq):
    s = _u8(ascii_seq)
    out = np.zeros(s.shape[0], dtype=np.uint8)
    cnt = np.zeros(4, dtype=np.uint64)
    bad = lib().or_encode_count_bases(_p(s), s.shape[0], _p(out), _p(cnt))
    if bad >= 0:
        raise ValueError(int(bad))
    return out, cnt


def get_differences(seq, other) -> np.ndarray:
    s, o = _u8(seq), _u8(other)
    out = np.zeros(s.shape[0], dtype=np.uint64)
    n = lib().or_get_differences(_p(s), _p(o), s.shape[0], _p(out))
    return out[:n].copy()


def consensus(seqs: np.ndarray) -> np.ndarray:
    s = _u8(seqs)
    assert s.ndim == 2
    out = np.zeros(s.shape[1], dtype=np.uint8)
    lib().or_consensus(_p(s), s.shape[0], s.shape[1], _p(out))
    return out


# ---- per-pair measures ----------------------------------------------------

def snp(q, t) -> int:
    q, t = _u8(q), _u8(t)
    return int(lib().or_snp(_p(q), _p(t), q.shape[0]))


def snp_consensus(q, t, qd, td) -> int:
    q, t = _u8(q), _u8(t)
    qd = np.ascontiguousarray(qd, dtype=np.uint64)
    td = np.ascontiguousarray(td, dtype=np.uint64)
    return int(lib().or_snp_consensus(_p(q), _p(t), _p(qd), qd.shape[0], _p(td), td.shape[0]))


def raw(q, t) -> float:
    q, t = _u8(q), _u8(t)
    return float(lib().or_raw(_p(q), _p(t), q.shape[0]))


def jc69(q, t) -> float:
    q, t = _u8(q), _u8(t)
    return float(lib().or_jc69(_p(q), _p(t), q.shape[0]))


def k80(q, t) -> float:
    q, t = _u8(q), _u8(t)
    return float(lib().or_k80(_p(q), _p(t), q.shape[0]))


def tn93(q, t, qc, tc) -> float:
    q, t = _u8(q), _u8(t)
    qc = np.ascontiguousarray(qc, dtype=np.uint64)
    tc = np.ascontiguousarray(tc, dtype=np.uint64)
    return float(lib().or_tn93(_p(q), _p(t), q.shape[0], _p(qc), _p(tc)))


def pair_counts(q, t) -> dict:
    """Raw integer counts of every site loop (measures.rs:16-20, 59-66, 85-107, 156-175)."""
    q, t = _u8(q), _u8(t)
    w = q.shape[0]
    o = np.zeros(9, dtype=np.uint64)
    L = lib()
    base = o.ctypes.data
    ptr = lambda k: C.c_void_p(base + 8 * k)
    L.or_raw_counts(_p(q), _p(t), w, ptr(0), ptr(1))
    L.or_k80_counts(_p(q), _p(t), w, ptr(2), ptr(3), ptr(4))
    L.or_tn93_counts(_p(q), _p(t), w, ptr(5), ptr(6), ptr(7), ptr(8))
    return {
        "snp": snp(q, t),
        "raw_d": int(o[0]), "raw_n": int(o[1]),
        "k80_L": int(o[2]), "k80_ts": int(o[3]), "k80_tv": int(o[4]),
        "tn93_L": int(o[5]), "tn93_d": int(o[6]), "tn93_P1": int(o[7]), "tn93_P2": int(o[8]),
    }


def format_float12(d: float) -> str:
    buf = C.create_string_buffer(64)
    lib().or_format_float12(d, buf, 64)
    return buf.value.decode()


# ---- whole-alignment drivers ------------------------------------------------

class Alignment:
    """Encoded records + the per-record precompute of set_up (lib.rs:219-241)."""

    def __init__(self, codes: np.ndarray, acgt: np.ndarray | None = None):
        self.codes = _u8(codes)
        assert self.codes.ndim == 2
        self.n, self.w = self.codes.shape
        self.acgt = None if acgt is None else np.ascontiguousarray(acgt, dtype=np.uint64)
        self.diff_off = None
        self.diff_idx = None

    def count_bases(self):
        self.acgt = np.stack([count_bases(r) for r in self.codes]) if self.n else np.zeros((0, 4), np.uint64)
        return self

    def differences(self, cons: np.ndarray):
        lists = [get_differences(r, cons) for r in self.codes]
        off = np.zeros(self.n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(x) for x in lists])
        self.diff_off = off
        self.diff_idx = (np.concatenate(lists) if lists else np.zeros(0)).astype(np.uint64)
        if self.diff_idx.shape[0] == 0:
            self.diff_idx = np.zeros(1, dtype=np.uint64)
        return self

    def c(self) -> _Aln:
        a = _Aln()
        a.seqs = self.codes.ctypes.data
        a.n, a.w = self.n, self.w
        a.acgt = None if self.acgt is None else self.acgt.ctypes.data
        a.diff_off = None if self.diff_off is None else self.diff_off.ctypes.data
        a.diff_idx = None if self.diff_idx is None else self.diff_idx.ctypes.data
        return a


def prepare(measure: str, alns: list[Alignment], consensus_from: list[Alignment] | None = None):
    """set_up's per-measure work: lib.rs:219-241 (n: consensus over the LOADED files +
    differences; tn93: count_bases when not already supplied)."""
    if measure == "n":
        src = consensus_from if consensus_from is not None else alns
        cons = consensus(np.concatenate([a.codes for a in src], axis=0))
        for a in alns:
            a.differences(cons)
    elif measure == "tn93":
        for a in alns:
            if a.acgt is None:
                a.count_bases()


def run(measure: str, mode: str, a: Alignment, b: Alignment | None = None, threads: int = 1):
    """Ordered results: 'square' (lib.rs:502-547), 'rect' (lib.rs:551-596), 'stream'
    (lib.rs:322-325; a = loaded, b = streamed; streamed-major).  Returns (array, seconds)."""
    m = MEASURES[measure]
    md = {"square": 0, "rect": 1, "stream": 2}[mode]
    if md == 0:
        total = a.n * (a.n - 1) // 2
    else:
        total = a.n * b.n
    ca = a.c()
    cb = (b if b is not None else a).c()
    secs = C.c_double(0)
    if measure in INT_MEASURES:
        out = np.zeros(max(total, 1), dtype=np.int64)
        lib().or_run(m, md, C.byref(ca), C.byref(cb), None, _p(out), threads, C.byref(secs))
    else:
        out = np.zeros(max(total, 1), dtype=np.float64)
        lib().or_run(m, md, C.byref(ca), C.byref(cb), _p(out), None, threads, C.byref(secs))
    return out[:total], secs.value


def bench(measure: str, mode: str, a: Alignment, b: Alignment | None, rows: int, threads: int):
    """Bounded CPU-baseline sample: first `rows` major rows. Returns (pairs, seconds)."""
    m = MEASURES[measure]
    md = {"square": 0, "rect": 1, "stream": 2}[mode]
    ca = a.c()
    cb = (b if b is not None else a).c()
    secs, cs = C.c_double(0), C.c_double(0)
    pairs = lib().or_bench(m, md, C.byref(ca), C.byref(cb), rows, threads, C.byref(secs), C.byref(cs))
    return int(pairs), secs.value


def tsv(ids1, ids2, mode: str, values, is_int: bool) -> str:
    """gather_write's text (lib.rs:612-644) for ordered results."""
    lines = ["sequence1\tsequence2\tdistance"]
    k = 0
    fmt = (lambda v: str(int(v))) if is_int else (lambda v: format_float12(float(v)))
    if mode == "square":
        n = len(ids1)
        for i in range(n - 1):
            for j in range(i + 1, n):
                lines.append(f"{ids1[i]}\t{ids1[j]}\t{fmt(values[k])}"); k += 1
    elif mode == "rect":
        for i in range(len(ids1)):
            for j in range(len(ids2)):
                lines.append(f"{ids1[i]}\t{ids2[j]}\t{fmt(values[k])}"); k += 1
    else:  # stream: ids1 = loaded, ids2 = streamed; streamed-major
        for j in range(len(ids2)):
            for i in range(len(ids1)):
                lines.append(f"{ids1[i]}\t{ids2[j]}\t{fmt(values[k])}"); k += 1
    return "\n".join(lines) + "\n"
