"""ctypes binding of include/distance_gpu.h (libdistance_gpu.so).

Mirrors the C ABI one to one; `Engine` is a convenience wrapper used by tests and bench.py.
No computation happens in Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "distance_b200", "_lib", "libdistance_gpu.so")

MEASURES = {"n": 0, "n_high": 1, "raw": 2, "jc69": 3, "k80": 4, "tn93": 5}
DG_INPUT_PARADIS, DG_INPUT_ASCII, DG_INPUT_NIBBLE = 0, 1, 2
DG_MODE_SQUARE, DG_MODE_RECT, DG_MODE_STREAM = 0, 1, 2
DG_RUN_DEVICE_ONLY, DG_RUN_REPACK = 1, 2
DG_OPT_PANEL_BYTES, DG_OPT_KEEP_CODES, DG_OPT_TILE_VARIANT, DG_OPT_ENGINE, DG_OPT_RESULT_U16, DG_OPT_PIPE_PANELS = 1, 2, 3, 4, 5, 6
DG_OPT_PIPE_CHUNK_BYTES = 7
DG_OPT_REPACK_OVERLAP = 8
DG_OPT_RESULT_U8 = 9
DG_OPT_RESULT_COUNTS = 10
DG_RESULT_U8 = 3
DG_RESULT_COUNTS16 = 4
DG_SQUARE_LOOKAHEAD = 8
DG_RESULT_U32, DG_RESULT_F64, DG_RESULT_U16 = 0, 1, 2
DG_ERR = {0: "DG_OK", -1: "DG_ERR_INVALID_ARG", -2: "DG_ERR_CUDA", -3: "DG_ERR_STATE",
          -4: "DG_ERR_INVALID_CODE", -5: "DG_ERR_SINK", -6: "DG_ERR_NOMEM"}

# every symbol include/distance_gpu.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "dg_abi_version", "dg_device_count", "dg_create", "dg_destroy", "dg_last_error", "dg_set_option",
    "dg_load_resident", "dg_load_resident_device", "dg_invalid_site", "dg_run_square", "dg_run_rect", "dg_run_part",
    "dg_stream_begin", "dg_stream_push", "dg_stream_buffer", "dg_stream_end", "dg_debug_counts", "dg_debug_planes",
    "dg_get_timings", "dg_reset_timings", "dg_alloc_pinned", "dg_free_pinned", "dg_plan_panels", "dg_plan_ctx",
    "dg_square_begin", "dg_square_next", "dg_square_plan", "dg_square_push", "dg_square_end", "dg_run_square_host",
    "dg_rect_begin", "dg_run_rect_host", "dg_plan_parts",
]


class DistanceGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{DG_ERR.get(code, code)}: {msg}")
        self.code = code
        self.msg = msg


class Panel(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("result_kind", C.c_int32),
        ("row_begin", C.c_uint64),
        ("row_end", C.c_uint64),
        ("n_cols", C.c_uint64),
        ("n_results", C.c_uint64),
        ("data", C.c_void_p),
        ("overflow", C.c_void_p),
        ("n_overflow", C.c_uint64),
    ]


def panel_values(p) -> np.ndarray:
    """The results of one dg_panel as a numpy array (a copy): uint32 / uint16 / float64 per result_kind; DG_RESULT_U8
    panels are widened to uint16 with their overflow list applied."""
    n = int(p.n_results)
    kind = int(p.result_kind)
    if kind == DG_RESULT_U8:
        v = np.frombuffer((C.c_uint8 * max(n, 1)).from_address(p.data), dtype=np.uint8, count=n).astype(np.uint16)
        k = int(p.n_overflow)
        if k:
            e = np.frombuffer((C.c_uint32 * (2 * k)).from_address(p.overflow), dtype=np.uint32, count=2 * k).reshape(k, 2)
            assert np.all(v[e[:, 0]] == 255)
            v[e[:, 0]] = e[:, 1].astype(np.uint16)
        return v
    if kind == DG_RESULT_COUNTS16:   # four uint16 counts per pair
        return np.frombuffer((C.c_uint8 * max(n * 8, 1)).from_address(p.data), dtype=np.uint16, count=4 * n).reshape(n, 4).copy()
    dtype = {DG_RESULT_U32: np.uint32, DG_RESULT_U16: np.uint16, DG_RESULT_F64: np.float64}[kind]
    isz = np.dtype(dtype).itemsize
    return np.frombuffer((C.c_uint8 * max(n * isz, 1)).from_address(p.data), dtype=dtype, count=n).copy()


class Timings(C.Structure):
    _fields_ = [
        ("pack_ms", C.c_double), ("count_ms", C.c_double), ("h2d_ms", C.c_double),
        ("total_ms", C.c_double), ("run_ms", C.c_double), ("pack_launches", C.c_uint64), ("count_launches", C.c_uint64),
        ("pairs", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("engine", C.c_uint64), ("sm_mhz", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Panel))

_lib = None


def library_path() -> str:
    return _LIB_PATH


def build_library(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a build of the in-tree library (Makefile `lib`)."""
    srcs = [os.path.join(_ROOT, "distance_b200", "csrc", f) for f in ("dg_api.cu", "kernels.cuh")]
    srcs.append(os.path.join(_ROOT, "include", "distance_gpu.h"))
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _ROOT, "lib"] + (["-B"] if force else []))
    return _LIB_PATH


def load_library():
    """Load libdistance_gpu.so.  Fails loudly if it is missing: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise DistanceGpuError(-2, f"{_LIB_PATH} is missing: run `make lib` (or __graft_entry__.build())")
    L = C.CDLL(_LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.dg_abi_version.restype = i32
    L.dg_device_count.restype = i32
    L.dg_create.argtypes = [C.POINTER(C.c_int), i32, i32, u64, C.POINTER(vp)]
    L.dg_destroy.argtypes = [vp]
    L.dg_destroy.restype = None
    L.dg_last_error.argtypes = [vp]
    L.dg_last_error.restype = C.c_char_p
    L.dg_set_option.argtypes = [vp, i32, C.c_int64]
    L.dg_load_resident.argtypes = [vp, i32, vp, u64, i32, vp]
    L.dg_load_resident_device.argtypes = [vp, i32, vp, i32, u64, i32, vp]
    L.dg_invalid_site.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_uint8)]
    L.dg_run_square.argtypes = [vp, SINK_FN, vp, C.c_uint32]
    L.dg_run_rect.argtypes = [vp, SINK_FN, vp, C.c_uint32]
    L.dg_run_part.argtypes = [vp, i32, C.c_uint32, C.c_uint32, SINK_FN, vp, C.c_uint32]
    L.dg_square_begin.argtypes = [vp, u64, i32, vp, C.c_uint32, C.c_uint32, SINK_FN, vp]
    L.dg_square_next.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.dg_square_push.argtypes = [vp, vp, i32, u64, u64, vp]
    L.dg_square_plan.argtypes = [vp, vp, vp, u64]
    L.dg_square_plan.restype = C.c_int64
    L.dg_square_end.argtypes = [vp]
    L.dg_run_square_host.argtypes = [vp, vp, u64, i32, vp, C.c_uint32, C.c_uint32, SINK_FN, vp]
    L.dg_rect_begin.argtypes = [vp, u64, i32, vp, C.c_uint32, C.c_uint32, SINK_FN, vp]
    L.dg_run_rect_host.argtypes = [vp, vp, u64, i32, vp, C.c_uint32, C.c_uint32, SINK_FN, vp]
    L.dg_stream_begin.argtypes = [vp, SINK_FN, vp, u64]
    L.dg_stream_push.argtypes = [vp, vp, u64, i32, vp]
    L.dg_stream_end.argtypes = [vp]
    L.dg_stream_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.dg_debug_counts.argtypes = [vp, i32, i32, vp]
    L.dg_debug_planes.argtypes = [vp, i32, vp, vp, vp, C.POINTER(u64)]
    L.dg_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.dg_reset_timings.argtypes = [vp]
    L.dg_alloc_pinned.argtypes = [C.c_size_t]
    L.dg_alloc_pinned.restype = vp
    L.dg_free_pinned.argtypes = [vp]
    L.dg_free_pinned.restype = None
    L.dg_plan_panels.argtypes = [i32, i32, u64, u64, u64, i32, vp, vp, vp, u64]
    L.dg_plan_panels.restype = C.c_int64
    L.dg_plan_ctx.argtypes = [vp, i32, vp, vp, vp, u64]
    L.dg_plan_ctx.restype = C.c_int64
    L.dg_plan_parts.argtypes = [vp, u64, C.c_uint32, vp]
    L.dg_plan_parts.restype = i32
    for name in ("dg_create", "dg_set_option", "dg_load_resident", "dg_load_resident_device", "dg_invalid_site", "dg_run_square",
                 "dg_run_rect", "dg_run_part", "dg_stream_begin", "dg_stream_push", "dg_stream_buffer", "dg_stream_end",
                 "dg_debug_counts", "dg_debug_planes", "dg_get_timings", "dg_reset_timings", "dg_square_begin",
                 "dg_square_next", "dg_square_push", "dg_square_end", "dg_run_square_host", "dg_rect_begin",
                 "dg_run_rect_host"):
        getattr(L, name).restype = i32
    _lib = L
    return L


def device_count() -> int:
    return int(load_library().dg_device_count())


_NULL_SINK = C.cast(None, SINK_FN)
DEFAULT_PANEL_BYTES = 256 << 20


def plan_panels(measure: str, mode: int, n_rows: int, n_cols: int, panel_bytes: int = DEFAULT_PANEL_BYTES,
                tile_variant: int = 0):
    """dg_plan_panels: [(row_begin, row_end, n_results)] -- host arithmetic only, needs no GPU."""
    L = load_library()
    n = L.dg_plan_panels(MEASURES[measure], mode, n_rows, n_cols, panel_bytes, tile_variant, None, None, None, 0)
    if n < 0:
        raise DistanceGpuError(int(n), "dg_plan_panels")
    rb, re_, nr = (np.zeros(max(n, 1), dtype=np.uint64) for _ in range(3))
    L.dg_plan_panels(MEASURES[measure], mode, n_rows, n_cols, panel_bytes, tile_variant,
                     rb.ctypes.data_as(C.c_void_p), re_.ctypes.data_as(C.c_void_p),
                     nr.ctypes.data_as(C.c_void_p), n)
    return [(int(rb[k]), int(re_[k]), int(nr[k])) for k in range(n)]


def pack_nibbles(codes: np.ndarray) -> np.ndarray:
    """Paradis bytes (n x width) -> DG_INPUT_NIBBLE rows (n x (width + 1) // 2): site 2k in the low nibble of byte k, site
    2k + 1 in the high one; a nibble is the possibility half of the code (N, '-' and '?' all become 15).  This is what a
    host parser would emit directly; tests and bench.py derive it from the byte codes."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n, w = codes.shape
    hi = codes >> 4
    if w % 2:
        hi = np.concatenate([hi, np.full((n, 1), 15, np.uint8)], axis=1)
    return np.ascontiguousarray(hi[:, 0::2] | (hi[:, 1::2] << 4))


def input_stride(width: int, input_kind: int) -> int:
    return (width + 1) // 2 if input_kind == DG_INPUT_NIBBLE else width


def plan_parts(plan, n_parts: int):
    """dg_plan_parts: the part that owns each panel of a plan [(row_begin, row_end, n_results)] (host arithmetic only)."""
    L = load_library()
    sizes = np.array([p[2] for p in plan], dtype=np.uint64)
    out = np.zeros(max(len(plan), 1), dtype=np.uint32)
    rc = L.dg_plan_parts(sizes.ctypes.data_as(C.c_void_p), len(plan), n_parts, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise DistanceGpuError(rc, "dg_plan_parts")
    return [int(x) for x in out[:len(plan)]]


def pinned_array(shape, dtype) -> np.ndarray:
    """A numpy array over page-locked memory from dg_alloc_pinned (freed with the process)."""
    L = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.dg_alloc_pinned(max(n, 1))
    if not p:
        raise DistanceGpuError(-6, "dg_alloc_pinned failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Engine:
    """One dg_ctx.  measure: 'n' | 'n_high' | 'raw' | 'jc69' | 'k80' | 'tn93'."""

    def __init__(self, measure: str, width: int, gpus=None):
        self.L = load_library()
        self.measure = measure
        self.width = int(width)
        ids = list(gpus) if gpus is not None else [0]
        arr = (C.c_int * len(ids))(*ids)
        h = C.c_void_p()
        rc = self.L.dg_create(arr, len(ids), MEASURES[measure], self.width, C.byref(h))
        if rc != 0:
            raise DistanceGpuError(rc, self.L.dg_last_error(None).decode())
        self.h = h
        self.is_int = measure in ("n", "n_high")
        self.u16 = False
        self.u8 = False
        self._n = [0, 0]

    # -- plumbing --------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise DistanceGpuError(rc, self.L.dg_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.dg_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: int, value: int):
        self._check(self.L.dg_set_option(self.h, key, value))
        if key == DG_OPT_RESULT_U16:
            self.u16 = bool(value) and self.is_int
        if key == DG_OPT_RESULT_U8:
            self.u8 = bool(value) and self.is_int

    def _dtype(self):
        """element type of the result panels (dg_panel.result_kind)"""
        return (np.uint16 if (self.u16 or self.u8) else np.uint32) if self.is_int else np.float64

    # -- inputs ----------------------------------------------------------------------------------
    def load(self, which: int, codes: np.ndarray, input_kind: int = DG_INPUT_PARADIS, acgt=None):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        assert codes.ndim == 2 and codes.shape[1] == input_stride(self.width, input_kind)
        cnt = None if acgt is None else np.ascontiguousarray(acgt, dtype=np.uint64)
        self._check(self.L.dg_load_resident(
            self.h, which, codes.ctypes.data_as(C.c_void_p), codes.shape[0], input_kind,
            None if cnt is None else cnt.ctypes.data_as(C.c_void_p)))
        self._n[which] = codes.shape[0]

    def load_device(self, which: int, device_ptr: int, src_device: int, n: int, input_kind: int = DG_INPUT_PARADIS):
        """dg_load_resident_device: n x width code bytes already in device memory (e.g. a torch tensor's data_ptr())."""
        self._check(self.L.dg_load_resident_device(self.h, which, C.c_void_p(device_ptr), src_device, n, input_kind, None))
        self._n[which] = n

    def invalid_site(self):
        r, s, b = C.c_uint64(), C.c_uint64(), C.c_uint8()
        self._check(self.L.dg_invalid_site(self.h, C.byref(r), C.byref(s), C.byref(b)))
        return int(r.value), int(s.value), int(b.value)

    def plan(self, mode: int = DG_MODE_SQUARE):
        """dg_plan_ctx: [(row_begin, row_end, n_results)] of the panels this engine will produce."""
        n = self.L.dg_plan_ctx(self.h, mode, None, None, None, 0)
        if n < 0:
            self._check(int(n))
        rb, re_, nr = (np.zeros(max(n, 1), dtype=np.uint64) for _ in range(3))
        self.L.dg_plan_ctx(self.h, mode, rb.ctypes.data_as(C.c_void_p), re_.ctypes.data_as(C.c_void_p),
                           nr.ctypes.data_as(C.c_void_p), n)
        return [(int(rb[k]), int(re_[k]), int(nr[k])) for k in range(n)]

    # -- runs ------------------------------------------------------------------------------------
    def _collect(self, total: int):
        dtype = self._dtype()
        out = np.empty(total, dtype=dtype)
        state = {"pos": 0, "panels": []}

        def sink(user, pp):
            p = pp.contents
            n = int(p.n_results)
            out[state["pos"]:state["pos"] + n] = panel_values(p)
            state["pos"] += n
            state["panels"].append((int(p.mode), int(p.row_begin), int(p.row_end), int(p.n_cols), n))
            return 0

        return out, state, SINK_FN(sink)

    def run_square(self, flags: int = 0):
        n = self._n[0]
        out, st, cb = self._collect(n * (n - 1) // 2)
        self._check(self.L.dg_run_square(self.h, cb, None, flags))
        assert st["pos"] == out.shape[0], (st["pos"], out.shape)
        self.last_panels = st["panels"]
        return out

    def run_square_counts(self):
        """dg_run_square with DG_OPT_RESULT_COUNTS set: the (pairs, 4) uint16 count tuples of every panel, in output order
        (panels that arrive as f64 -- LOP3 engine, wide alignments -- raise)."""
        chunks = []

        def sink(user, pp):
            p = pp.contents
            assert int(p.result_kind) == DG_RESULT_COUNTS16, int(p.result_kind)
            chunks.append(panel_values(p))
            return 0

        self._check(self.L.dg_run_square(self.h, SINK_FN(sink), None, 0))
        return np.concatenate(chunks) if chunks else np.zeros((0, 4), np.uint16)

    def run_rect(self, flags: int = 0):
        out, st, cb = self._collect(self._n[0] * self._n[1])
        self._check(self.L.dg_run_rect(self.h, cb, None, flags))
        assert st["pos"] == out.shape[0]
        self.last_panels = st["panels"]
        return out

    def run_part(self, mode: int, part: int, n_parts: int, flags: int = 0):
        """Returns [(row_begin, row_end, values)] for this part's panels."""
        dtype = self._dtype()
        got = []

        def sink(user, pp):
            p = pp.contents
            got.append((int(p.row_begin), int(p.row_end), panel_values(p).astype(dtype)))
            return 0

        self._check(self.L.dg_run_part(self.h, mode, part, n_parts, SINK_FN(sink), None, flags))
        return got

    def run_device_only(self, mode: int = DG_MODE_SQUARE, part: int = 0, n_parts: int = 1, repack: bool = False):
        flags = DG_RUN_DEVICE_ONLY | (DG_RUN_REPACK if repack else 0)
        self._check(self.L.dg_run_part(self.h, mode, part, n_parts, _NULL_SINK, None, flags))

    def run_discard(self, mode: int = DG_MODE_SQUARE, part: int = 0, n_parts: int = 1, touch: bool = True):
        """Full e2e run (D2H into pinned panels + sink) whose sink only reads one word per panel."""
        state = {"n": 0, "acc": 0}

        def sink(user, pp):
            p = pp.contents
            state["n"] += int(p.n_results)
            if touch and p.n_results:
                state["acc"] ^= C.c_uint32.from_address(p.data).value
            return 0

        self._check(self.L.dg_run_part(self.h, mode, part, n_parts, SINK_FN(sink), None, 0))
        return state["n"]

    # -- pipelined all-vs-all (dg_square_*) --------------------------------------------------------
    def square_pipelined(self, codes: np.ndarray, input_kind: int = DG_INPUT_PARADIS, acgt=None, part: int = 0,
                         n_parts: int = 1, one_call: bool = False, push=None):
        """dg_square_begin / next / push / end over host codes.  Returns (values, panels): the packed upper
        triangle with this part's panels placed by row (other parts' positions stay 0) and the panels in the
        order the sink saw them.  `push(lo, hi)` may replace the default host-pointer push (multi-rank tests)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        n = codes.shape[0]
        stride = input_stride(self.width, input_kind)
        assert codes.ndim == 2 and codes.shape[1] == stride
        dtype = self._dtype()
        out = np.zeros(n * (n - 1) // 2, dtype=dtype)
        panels = []

        def sink(user, pp):
            p = pp.contents
            cnt = int(p.n_results)
            r0 = int(p.row_begin)
            base = r0 * (2 * n - r0 - 1) // 2
            out[base:base + cnt] = panel_values(p)
            panels.append((int(p.mode), r0, int(p.row_end), int(p.n_cols), cnt))
            return 0

        cb = SINK_FN(sink)
        cnt = None if acgt is None else np.ascontiguousarray(acgt, dtype=np.uint64)
        cptr = None if cnt is None else cnt.ctypes.data_as(C.c_void_p)
        if one_call:
            self._check(self.L.dg_run_square_host(self.h, codes.ctypes.data_as(C.c_void_p), n, input_kind, cptr,
                                                  part, n_parts, cb, None))
        else:
            self._check(self.L.dg_square_begin(self.h, n, input_kind, cptr, part, n_parts, cb, None))
            lo, hi = C.c_uint64(), C.c_uint64()
            while True:
                self._check(self.L.dg_square_next(self.h, C.byref(lo), C.byref(hi)))
                if hi.value == lo.value:
                    break
                if push is not None:
                    push(int(lo.value), int(hi.value))
                else:
                    self._check(self.L.dg_square_push(self.h, C.c_void_p(codes.ctypes.data + lo.value * stride), -1,
                                                      lo.value, hi.value, None))
            self._check(self.L.dg_square_end(self.h))
        self._n[0] = n
        self.last_panels = panels
        return out, panels

    def rect_pipelined(self, codes_a: np.ndarray, input_kind: int = DG_INPUT_PARADIS, acgt=None, part: int = 0,
                       n_parts: int = 1, one_call: bool = True):
        """dg_run_rect_host (or dg_rect_begin + next / push / end): alignment 1 must be loaded.  Returns (values of this
        part's panels concatenated in the order the sink saw them, panels)."""
        codes_a = np.ascontiguousarray(codes_a, dtype=np.uint8)
        n = codes_a.shape[0]
        dtype = self._dtype()
        chunks, panels = [], []

        def sink(user, pp):
            p = pp.contents
            cnt = int(p.n_results)
            chunks.append(panel_values(p).astype(dtype))
            panels.append((int(p.mode), int(p.row_begin), int(p.row_end), int(p.n_cols), cnt))
            return 0

        cb = SINK_FN(sink)
        cnt = None if acgt is None else np.ascontiguousarray(acgt, dtype=np.uint64)
        cptr = None if cnt is None else cnt.ctypes.data_as(C.c_void_p)
        if one_call:
            self._check(self.L.dg_run_rect_host(self.h, codes_a.ctypes.data_as(C.c_void_p), n, input_kind, cptr,
                                                part, n_parts, cb, None))
        else:
            self._check(self.L.dg_rect_begin(self.h, n, input_kind, cptr, part, n_parts, cb, None))
            for lo, hi in self.square_plan():
                self.square_push(codes_a.ctypes.data + lo * input_stride(self.width, input_kind), -1, lo, hi)
            self.square_end()
        self._n[0] = n
        self.last_panels = panels
        return (np.concatenate(chunks) if chunks else np.zeros(0, dtype=dtype)), panels

    def square_pipelined_discard(self, pinned_codes: np.ndarray, part: int = 0, n_parts: int = 1,
                                 input_kind: int = DG_INPUT_PARADIS):
        """dg_run_square_host with a sink that reads one word per panel (bench.py's e2e step)."""
        state = {"n": 0, "acc": 0}

        def sink(user, pp):
            p = pp.contents
            state["n"] += int(p.n_results)
            if p.n_results:
                state["acc"] ^= C.c_uint32.from_address(p.data).value
            return 0

        n = pinned_codes.shape[0]
        assert pinned_codes.shape[1] == input_stride(self.width, input_kind)
        self._check(self.L.dg_run_square_host(self.h, C.c_void_p(pinned_codes.ctypes.data), n, input_kind, None,
                                              part, n_parts, SINK_FN(sink), None))
        self._n[0] = n
        return state["n"]

    def square_begin(self, n: int, sink, part: int = 0, n_parts: int = 1, input_kind: int = DG_INPUT_PARADIS):
        """dg_square_begin with a raw SINK_FN; returns the chunk plan [(lo, hi)] by walking dg_square_next is not
        possible without pushing, so callers use square_chunks() for the plan."""
        self._check(self.L.dg_square_begin(self.h, n, input_kind, None, part, n_parts, sink, None))
        self._n[0] = n

    def square_plan(self):
        """dg_square_plan: [(lo, hi)] of every chunk of the open session, in push order."""
        n = self.L.dg_square_plan(self.h, None, None, 0)
        if n < 0:
            self._check(int(n))
        lo, hi = (np.zeros(max(n, 1), dtype=np.uint64) for _ in range(2))
        self.L.dg_square_plan(self.h, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), n)
        return [(int(lo[k]), int(hi[k])) for k in range(n)]

    def square_next(self):
        lo, hi = C.c_uint64(), C.c_uint64()
        self._check(self.L.dg_square_next(self.h, C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def square_push(self, ptr: int, src_device: int, lo: int, hi: int, ready_event: int = 0):
        self._check(self.L.dg_square_push(self.h, C.c_void_p(ptr), src_device, lo, hi,
                                          C.c_void_p(ready_event) if ready_event else None))

    def square_end(self):
        self._check(self.L.dg_square_end(self.h))

    def stream(self, batches, input_kind: int = DG_INPUT_PARADIS, max_batch: int = 1 << 20, acgt_batches=None):
        """Stream an iterable of (n_b x width) uint8 arrays against alignment 0."""
        dtype = self._dtype()
        chunks, panels = [], []

        def sink(user, pp):
            p = pp.contents
            n = int(p.n_results)
            chunks.append(panel_values(p).astype(dtype))
            panels.append((int(p.mode), int(p.row_begin), int(p.row_end), int(p.n_cols), n))
            return 0

        cb = SINK_FN(sink)
        self._check(self.L.dg_stream_begin(self.h, cb, None, max_batch))
        for k, b in enumerate(batches):
            b = np.ascontiguousarray(b, dtype=np.uint8)
            cnt = None
            if acgt_batches is not None:
                cnt = np.ascontiguousarray(acgt_batches[k], dtype=np.uint64)
            self._check(self.L.dg_stream_push(
                self.h, b.ctypes.data_as(C.c_void_p), b.shape[0], input_kind,
                None if cnt is None else cnt.ctypes.data_as(C.c_void_p)))
        self._check(self.L.dg_stream_end(self.h))
        self.last_panels = panels
        return np.concatenate(chunks) if chunks else np.zeros(0, dtype=dtype)

    # -- debug / parity --------------------------------------------------------------------------
    def debug_counts(self, which_a: int = 0, which_b: int = 0) -> np.ndarray:
        out = np.zeros((self._n[which_a], self._n[which_b], 4), dtype=np.uint32)
        self._check(self.L.dg_debug_counts(self.h, which_a, which_b, out.ctypes.data_as(C.c_void_p)))
        return out

    def debug_planes(self, which: int = 0, want_aux: bool = True):
        words = C.c_uint64()
        self._check(self.L.dg_debug_planes(self.h, which, None, None, None, C.byref(words)))
        n, w = self._n[which], int(words.value)
        core = np.zeros((n, w, 4), dtype=np.uint32)
        aux = np.zeros((n, w, 4), dtype=np.uint32) if want_aux else None
        acgt = np.zeros((n, 4), dtype=np.uint64)
        self._check(self.L.dg_debug_planes(
            self.h, which, core.ctypes.data_as(C.c_void_p),
            None if aux is None else aux.ctypes.data_as(C.c_void_p),
            acgt.ctypes.data_as(C.c_void_p), C.byref(words)))
        return core, aux, acgt

    def timings(self) -> dict:
        t = Timings()
        self._check(self.L.dg_get_timings(self.h, C.byref(t)))
        return t.as_dict()

    def reset_timings(self):
        self._check(self.L.dg_reset_timings(self.h))
