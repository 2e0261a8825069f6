"""Full-size GPU parity for BASELINE configs 3, 4 and 5 (sampled against the CPU oracle) and the exactness limits of
the tensor engines on very wide alignments.  Everything goes through the C ABI (libdistance_gpu.so).

The oracle cannot finish 10^8 .. 10^9 pairs of width 29,903 in seconds, so each test compares a few thousand pairs --
whole rows and random positions, every NaN / inf / -0.0 the sample hits included -- and adds a size-independent
property of the whole result (symmetry under swapping the two alignments, panel-size independence)."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_parity import assert_float_parity, check, oracle_run

pytestmark = pytest.mark.gpu

W = 29903


@pytest.fixture(scope="module")
def dg():
    import distance_b200 as d
    d.load_library()
    assert d.device_count() >= 1, "no CUDA device: the gpu tests must run on the B200 box"
    return d


def spike(asc, rng):
    """Plant the special cases into a synthetic alignment: an all-N record (no comparable site: NaN), a copy of a
    record (identical pair: -0.0 / 0.0), a heavily mutated record (saturated distances: NaN / inf candidates)."""
    asc = asc.copy()
    n, w = asc.shape
    asc[3, :] = ord("N")
    asc[5, :] = asc[4, :]
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    asc[7, :] = bases[rng.integers(0, 4, size=w)]
    return asc


def test_config3_tn93_rect_full_size(dg, oracle):
    """BASELINE config 3: tn93 between two alignments, 10,000 x 10,000 x 29,903 (measures.rs:116-193)."""
    from distance_b200 import api, synth
    rng = np.random.default_rng(3)
    n = 10000
    root = synth.make_root(W, 20251018 + 3)
    a_asc = spike(synth.make_alignment(n, seed=20251018 + 3, ambiguity=True, root=root), rng)
    b_asc = spike(synth.make_alignment(n, seed=20251018 + 33, ambiguity=True, root=root), rng)
    a, b = synth.encode_ascii(a_asc), synth.encode_ascii(b_asc)
    with dg.Engine("tn93", W) as e:
        e.load(0, a)
        e.load(1, b)
        got = e.run_rect().reshape(n, n)
        assert e.timings()["engine"] == 3
        # the pipelined two-file session gives the same bits
        got2, _ = e.rect_pipelined(a)
        assert np.array_equal(got2.view(np.uint64), got.reshape(-1).view(np.uint64))
    rows = [0, 3, 4, 5, 7, 4999, 9999]
    want = oracle_run(oracle, "tn93", "rect", a[rows], b).reshape(len(rows), n)
    assert_float_parity(got[rows], want)
    assert np.isnan(want[1]).all()                       # the all-N record
    assert np.array_equal(want[2], want[3], equal_nan=True)   # the copied record
    cols = [0, 3, 5, 7, 1234, 9999]
    want_c = oracle_run(oracle, "tn93", "rect", a, b[cols]).reshape(n, len(cols))
    assert_float_parity(got[:, cols], want_c)
    # swapping the two files transposes the matrix (tn93 is symmetric in query / target up to the order of the f64
    # sums of base counts, which the reference keeps: measures.rs:118-143 adds target first)
    with dg.Engine("tn93", W) as e:
        e.load(0, b[:600])
        e.load(1, a[:700])
        t = e.run_rect().reshape(600, 700)
    assert_float_parity(t, oracle_run(oracle, "tn93", "rect", b[:600], a[:700]).reshape(600, 700))


def test_config4_k80_stream_full_size(dg, oracle):
    """BASELINE config 4: k80, 1,000 resident records against a stream of >= 100,000 x 29,903 records in pinned
    double-buffered batches (lib.rs:322-325; measures.rs:80-113).  The streamed records cycle through a pool of 4,096
    distinct records; sampled batches are compared with the oracle."""
    from distance_b200 import api, synth
    rng = np.random.default_rng(4)
    root = synth.make_root(W, 20251018 + 4)
    res = synth.encode_ascii(spike(synth.make_alignment(1000, seed=20251018 + 4, ambiguity=True, root=root), rng))
    pool_n, batch, n_batches = 4096, 4096, 25          # 102,400 streamed records, 1.024e8 pairs
    pool = synth.encode_ascii(spike(synth.make_alignment(pool_n, seed=20251018 + 44, ambiguity=True, root=root), rng))
    ppool = api.pinned_array(pool.shape, np.uint8)
    ppool[...] = pool
    keep = {0: None, 11: None, 24: None}
    state = {"rows": 0, "batches": 0, "sum": 0.0}
    with dg.Engine("k80", W) as e:
        e.load(0, res)

        def sink(user, pp):
            p = pp.contents
            assert p.mode == api.DG_MODE_STREAM and p.n_cols == 1000 and p.row_begin == state["rows"]
            cnt = int(p.n_results)
            v = np.frombuffer((C.c_uint8 * (cnt * 8)).from_address(p.data), dtype=np.float64, count=cnt)
            if state["batches"] in keep:
                keep[state["batches"]] = v.copy()
            state["rows"] = int(p.row_end)
            state["batches"] += 1
            return 0

        cb = api.SINK_FN(sink)
        e._check(e.L.dg_stream_begin(e.h, cb, None, batch))
        for k in range(n_batches):
            off = (k * 7) % 64                          # every batch starts somewhere else in the pool
            rolled = np.roll(pool, -off, axis=0) if k in keep else None
            if k in keep:
                e._check(e.L.dg_stream_push(e.h, rolled.ctypes.data_as(C.c_void_p), batch, api.DG_INPUT_PARADIS, None))
            else:
                e._check(e.L.dg_stream_push(e.h, C.c_void_p(ppool.ctypes.data), batch, api.DG_INPUT_PARADIS, None))
        e._check(e.L.dg_stream_end(e.h))
        assert e.timings()["engine"] == 3
    assert state["rows"] == batch * n_batches and state["batches"] == n_batches
    for k, v in keep.items():
        streamed = np.roll(pool, -((k * 7) % 64), axis=0)
        rows = [0, 1, 2, 3, 4, 5, 6, 7, 2000, 4095]
        want = oracle_run(oracle, "k80", "stream", res, streamed[rows]).reshape(len(rows), 1000)
        assert_float_parity(v.reshape(batch, 1000)[rows], want)


def test_config5_jc69_share_of_100k(dg, oracle):
    """BASELINE config 5 / the north-star target: jc69 all-vs-all over 100,000 x 29,903 records; ONE rank's share
    (part 0 of 8, what one of 8 B200s computes), rows sampled against the oracle (measures.rs:56-77)."""
    from distance_b200 import api, synth
    rng = np.random.default_rng(5)
    n = 100000
    asc = spike(synth.make_alignment(n, seed=20251018 + 5, ambiguity=True), rng)
    lut = synth.ascii_lut()
    with dg.Engine("jc69", W) as e:
        e.load(0, asc, input_kind=api.DG_INPUT_ASCII)
        plan = e.plan(api.DG_MODE_SQUARE)
        from distance_b200 import dist
        mine = dist.my_panels(plan, 0, 8)
        sample_rows = {}
        for (r0, r1, _) in (mine[0], mine[len(mine) // 2], mine[-1]):
            for r in (r0, (r0 + r1) // 2, r1 - 1):
                sample_rows[r] = None
        for r in (3, 4, 5, 7):
            if mine[0][0] <= r < mine[0][1]:
                sample_rows[r] = None
        seen = {"pairs": 0}

        def sink(user, pp):
            p = pp.contents
            r0, r1, cnt = int(p.row_begin), int(p.row_end), int(p.n_results)
            base = r0 * (2 * n - r0 - 1) // 2
            for r in sample_rows:
                if r0 <= r < r1:
                    o = r * (2 * n - r - 1) // 2 - base
                    ln = n - 1 - r
                    sample_rows[r] = np.frombuffer((C.c_uint8 * (ln * 8)).from_address(p.data + o * 8), dtype=np.float64, count=ln).copy()
            seen["pairs"] += cnt
            return 0

        e._check(e.L.dg_run_part(e.h, api.DG_MODE_SQUARE, 0, 8, api.SINK_FN(sink), None, 0))
        assert e.timings()["engine"] == 3
    assert seen["pairs"] == sum(p[2] for p in mine)
    for r, got in sample_rows.items():
        assert got is not None, r
        cols = np.unique(np.concatenate([np.arange(r + 1, min(n, r + 1 + 300)), rng.integers(r + 1, n, size=300), [n - 1]])) if r < n - 1 else np.array([], int)
        if cols.size == 0:
            continue
        want = np.array([oracle.jc69(lut[asc[r]], lut[asc[j]]) for j in cols])
        assert_float_parity(got[cols - r - 1], want)


@pytest.mark.parametrize("measure", ["n_high", "jc69", "tn93"])
def test_wide_alignment_exactness(dg, oracle, measure):
    """Widths beyond the fp32-exact range of the kind::mxf4 engine (4 * width >= 2^24): auto must route to the int8
    engine (int32 accumulation) and every engine that accepts the width must agree with the oracle bit for bit on the
    counts.  Records are crafted so that single sums exceed 2^24 (all-different and all-equal pairs)."""
    from distance_b200 import api, synth
    width, n = 5_700_000, 12
    rng = np.random.default_rng(57)
    codes = synth.random_codes(rng, n, width, p_ambig=1e-4)   # few partial codes: auto keeps the tensor engine (use_tc)
    codes[0, :] = 136                      # all A
    codes[1, :] = 40                       # all C: DIFF(0, 1) = width, 3 * width > 2^24
    codes[2, :] = 136                      # equal to record 0: SAME = width
    codes[3, 1::2] = 24
    want = oracle_run(oracle, measure, "square", codes)
    for engine, expect in ((0, 2), (2, 2), (1, 1)):
        with dg.Engine(measure, width) as e:
            e.set_option(api.DG_OPT_ENGINE, engine)
            e.load(0, codes)
            got = e.run_square()
            assert e.timings()["engine"] == expect
        check(measure, got, want)
    with dg.Engine(measure, width) as e:
        with pytest.raises(api.DistanceGpuError) as ei:
            e.set_option(api.DG_OPT_ENGINE, 3)
        assert ei.value.code == -1        # DG_ERR_INVALID_ARG: fp32 accumulation would not be exact


@pytest.mark.parametrize("measure", ["n_high", "raw", "jc69", "k80", "tn93"])
@pytest.mark.parametrize("engine", [1, 2, 3])
def test_width_beyond_int16_scratch(dg, oracle, measure, engine):
    """Widths above 32,767 take the 32-bit scratch layout of the tensor engines (sums no longer fit int16 / the
    modulo-2^16 trick of accumulator 0); counts above 65,535 need the uint32 results."""
    from distance_b200 import api, synth
    rng = np.random.default_rng(40)
    n, width = 260, 70000
    codes = synth.random_codes(rng, n, width, p_ambig=0.05)
    codes[0, :] = 136
    codes[1, :] = 40            # DIFF(0, 1) = 70,000 > 65,535
    want = oracle_run(oracle, measure, "square", codes)
    with dg.Engine(measure, width) as e:
        e.set_option(api.DG_OPT_ENGINE, engine)
        e.load(0, codes)
        got = e.run_square()
        assert e.timings()["engine"] == engine
    check(measure, got, want)


@pytest.mark.parametrize("measure", ["k80", "tn93"])
def test_fast_epilogues_match_the_literal_ones(dg, oracle, measure, monkeypatch):
    """The k80 / tn93 epilogues divide with shared reciprocals (kernels.cuh: dg_rcp / dg_div, the compiler's own
    division sequence with the denominator-only steps hoisted).  They must give the same BITS as the literal
    expressions of measures.rs:80-193 compiled with plain divisions (DG_EPI_LITERAL=1), special values included."""
    from distance_b200 import synth
    rng = np.random.default_rng(93)
    n, width = 700, 3000
    codes = synth.random_codes(rng, n, width, p_ambig=0.1)
    near = synth.encode_ascii(synth.make_alignment(300, width=width, seed=9, ambiguity=True, mu=5e-3))
    codes[:300] = near                      # realistic pairs (few differences) next to the saturated random ones
    codes[300, :] = 240                     # all N: no compared site
    codes[301, :] = 136                     # all A: base frequencies of 0 in some pairs
    codes[302, :] = 136
    codes[303, :] = 24                      # all T against all A
    res = {}
    for literal in (False, True):
        if literal:
            monkeypatch.setenv("DG_EPI_LITERAL", "1")
        with dg.Engine(measure, width) as e:
            e.load(0, codes)
            res[literal] = e.run_square()
            assert e.timings()["engine"] == 3
    assert np.array_equal(res[False].view(np.uint64), res[True].view(np.uint64))
    check(measure, res[False], oracle_run(oracle, measure, "square", codes))


@pytest.mark.parametrize("measure", ["n_high", "jc69", "k80", "tn93"])
@pytest.mark.parametrize("width", [1, 2, 301, 3000])
def test_nibble_input_matches_byte_input(dg, oracle, measure, width):
    """DG_INPUT_NIBBLE (two sites per byte, the possibility half of the Paradis code) through every entry point: resident
    load, the pipelined session, two files and the stream give the same results as the byte codes (and the oracle)."""
    from distance_b200 import api, synth
    rng = np.random.default_rng(width * 7 + len(measure))
    n = 300
    codes = synth.random_codes(rng, n, width, p_ambig=0.08)
    nib = api.pack_nibbles(codes)
    assert nib.shape == (n, (width + 1) // 2)
    want = oracle_run(oracle, measure, "square", codes)
    with dg.Engine(measure, width) as e:
        e.load(0, nib, input_kind=api.DG_INPUT_NIBBLE)
        check(measure, e.run_square(), want)
        e.set_option(api.DG_OPT_PIPE_CHUNK_BYTES, 128 * width)
        got, _ = e.square_pipelined(nib, input_kind=api.DG_INPUT_NIBBLE)
        check(measure, got, want)
        a, b = codes[:130], codes[130:]
        e.load(1, api.pack_nibbles(b), input_kind=api.DG_INPUT_NIBBLE)
        got, _ = e.rect_pipelined(api.pack_nibbles(a), input_kind=api.DG_INPUT_NIBBLE)
        check(measure, got, oracle_run(oracle, measure, "rect", a, b))
        e.load(0, api.pack_nibbles(a), input_kind=api.DG_INPUT_NIBBLE)
        got = e.stream([api.pack_nibbles(b[i:i + 50]) for i in range(0, 170, 50)], input_kind=api.DG_INPUT_NIBBLE, max_batch=64)
        check(measure, got, oracle_run(oracle, measure, "stream", a, b))


def test_nibble_zero_is_reported_as_invalid(dg):
    from distance_b200 import api, synth
    rng = np.random.default_rng(5)
    codes = synth.random_codes(rng, 40, 100)
    nib = api.pack_nibbles(codes)
    nib[17, 21] &= 0x0F                       # site 43 of record 17 -> nibble 0: no base possible
    with dg.Engine("n_high", 100) as e:
        with pytest.raises(api.DistanceGpuError) as ei:
            e.load(0, nib, input_kind=api.DG_INPUT_NIBBLE)
        assert ei.value.code == -4            # DG_ERR_INVALID_CODE
        assert e.invalid_site()[:2] == (17, 43)
    # the pipelined session and the stream pack straight from the nibble rows: same report
    with dg.Engine("n_high", 100) as e:
        with pytest.raises(api.DistanceGpuError) as ei:
            e.square_pipelined(nib, input_kind=api.DG_INPUT_NIBBLE)
        assert ei.value.code == -4 and e.invalid_site() == (17, 43, 0)
        got, _ = e.square_pipelined(api.pack_nibbles(codes), input_kind=api.DG_INPUT_NIBBLE)   # the context is usable again
        assert got.shape == (40 * 39 // 2,)
    with dg.Engine("k80", 100) as e:
        e.load(0, codes)
        with pytest.raises(api.DistanceGpuError) as ei:
            e.stream([nib], input_kind=api.DG_INPUT_NIBBLE, max_batch=64)
        assert ei.value.code == -4


@pytest.mark.parametrize("engine", [1, 3])
def test_u8_results_with_overflow_list(dg, oracle, engine):
    """DG_OPT_RESULT_U8: n / n_high panels as uint8 with an overflow list for the counts >= 255 (resident run, pipelined
    session, two files); a panel with more overflows than the list holds arrives as uint16.  Values identical either way."""
    from distance_b200 import api, synth
    width = 2000
    near = synth.encode_ascii(synth.make_alignment(900, width=width, seed=12, ambiguity=True, mu=2e-2))   # counts ~ 80
    rng = np.random.default_rng(8)
    far = synth.random_codes(rng, 3, width, p_ambig=0.02)                                                    # counts ~ 1,400
    codes = np.concatenate([near[:500], far, near[500:]])
    want = oracle_run(oracle, "n_high", "square", codes)
    assert (want >= 255).sum() > 1000 and (want == 255).sum() >= 0
    kinds = set()
    with dg.Engine("n_high", width) as e:
        e.set_option(api.DG_OPT_ENGINE, engine)
        e.set_option(api.DG_OPT_RESULT_U8, 1)
        e.set_option(api.DG_OPT_PANEL_BYTES, 1 << 20)
        e.load(0, codes)
        got = e.run_square()
        kinds |= {k for k in getattr(e, "last_kinds", [])}
        check("n_high", got, want)
        got, _ = e.square_pipelined(codes)
        check("n_high", got, want)
        e.load(1, codes[400:])
        got, _ = e.rect_pipelined(codes[:400])
        check("n_high", got, oracle_run(oracle, "n_high", "rect", codes[:400], codes[400:]))
    # a saturated alignment: every count is large, every panel falls back to uint16
    sat = synth.random_codes(rng, 700, width, p_ambig=0.01)
    with dg.Engine("n", width) as e:
        e.set_option(api.DG_OPT_ENGINE, engine)
        e.set_option(api.DG_OPT_RESULT_U8, 1)
        e.load(0, sat)
        seen = []

        def sink(user, pp):
            seen.append(int(pp.contents.result_kind))
            return 0

        e._check(e.L.dg_run_square(e.h, api.SINK_FN(sink), None, 0))
        assert seen and all(k == api.DG_RESULT_U16 for k in seen)
        check("n", e.run_square(), oracle_run(oracle, "n", "square", sat))


@pytest.mark.parametrize("measure", ["raw", "jc69", "k80", "tn93"])
def test_count_tuples_instead_of_f64(dg, oracle, measure):
    """DG_OPT_RESULT_COUNTS: float measures deliver the integer counts (DG_RESULT_COUNTS16) the host evaluates with its own
    libm; they equal the debug counts of the same engine, and the reference's expressions on them give the oracle's values
    bit for bit (raw is checked here; the C++ host's text for jc69 / k80 / tn93 is compared in test_cli_gpu.py)."""
    from distance_b200 import api, synth
    rng = np.random.default_rng(44)
    n, width = 700, 1500
    codes = synth.random_codes(rng, n, width, p_ambig=0.1)
    codes[:300] = synth.encode_ascii(synth.make_alignment(300, width=width, seed=3, ambiguity=True, mu=1e-2))
    with dg.Engine(measure, width) as e:
        e.set_option(api.DG_OPT_ENGINE, 3)            # (auto would send this ambiguity-heavy input to the LOP3 tiles: f64 panels)
        e.set_option(api.DG_OPT_RESULT_COUNTS, 1)
        e.set_option(api.DG_OPT_PANEL_BYTES, 1 << 20)
        e.load(0, codes)
        got = e.run_square_counts()
        dbg = e.debug_counts(0, 0)
    iu = np.triu_indices(n, 1)
    assert np.array_equal(got.astype(np.uint32), dbg[iu])
    if measure == "raw":
        want = oracle_run(oracle, "raw", "square", codes)
        with np.errstate(invalid="ignore", divide="ignore"):
            mine = got[:, 0].astype(np.float64) / (got[:, 0].astype(np.float64) + got[:, 1].astype(np.float64))
        assert np.array_equal(mine.view(np.uint64)[~np.isnan(want)], want.view(np.uint64)[~np.isnan(want)])
        assert np.array_equal(np.isnan(mine), np.isnan(want))
