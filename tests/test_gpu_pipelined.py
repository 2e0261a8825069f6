"""GPU parity tests of the pipelined all-vs-all session (dg_square_begin / next / push / end and
dg_run_square_host): chunks pushed highest-records-first, panels delivered in completion order, results
identical to the oracle (and to dg_load_resident + dg_run_square) on the same inputs.  Tiny chunk / panel
sizes force many chunks and panels on small alignments."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_parity import ALL, check, oracle_run

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["lop3", "tc", "fp4", "auto"])
def engine(request, monkeypatch):
    monkeypatch.setenv("DG_ENGINE", {"lop3": "1", "tc": "2", "fp4": "3", "auto": "0"}[request.param])
    return request.param


@pytest.fixture(scope="module")
def dg():
    import distance_b200 as d
    d.load_library()
    assert d.device_count() >= 1, "no CUDA device: the gpu tests must run on the B200 box"
    return d


def small_pieces(e, width, chunk_records=128, panel_bytes=4096):
    from distance_b200 import api
    e.set_option(api.DG_OPT_PIPE_CHUNK_BYTES, chunk_records * width)
    e.set_option(api.DG_OPT_PANEL_BYTES, panel_bytes)


@pytest.mark.parametrize("measure", ALL)
@pytest.mark.parametrize("n,width,amb", [(2, 7, 0.3), (33, 100, 0.3), (700, 300, 0.05), (1300, 131, 0.4)])
def test_pipelined_square_matches_oracle(dg, oracle, measure, n, width, amb):
    from distance_b200 import synth
    rng = np.random.default_rng(n * 7 + width)
    codes = synth.random_codes(rng, n, width, p_ambig=amb)
    with dg.Engine(measure, width) as e:
        small_pieces(e, width, panel_bytes=1 << 20)
        got, panels = e.square_pipelined(codes)
        assert sum(p[4] for p in panels) == n * (n - 1) // 2
        check(measure, got, oracle_run(oracle, measure, "square", codes))
        # the alignment is resident afterwards exactly as after dg_load_resident
        again = e.run_square()
        check(measure, again, oracle_run(oracle, measure, "square", codes))


@pytest.mark.parametrize("measure", ["n_high", "jc69", "tn93"])
def test_pipelined_many_panels_descending_and_parts(dg, oracle, measure):
    from distance_b200 import synth
    n, width = 2100, 257
    rng = np.random.default_rng(99)
    codes = synth.random_codes(rng, n, width, p_ambig=0.1)
    want = oracle_run(oracle, measure, "square", codes)
    total = np.zeros_like(want)
    seen = n_panels = 0
    for part in range(3):
        with dg.Engine(measure, width) as e:
            small_pieces(e, width, chunk_records=256, panel_bytes=600 * 1024)
            got, panels = e.square_pipelined(codes, part=part, n_parts=3)
        rows = [p[1] for p in panels]
        n_panels += len(panels)
        # tensor engines deliver in completion order = descending rows; the LOP3 engine ascending
        assert rows == sorted(rows, reverse=True) or rows == sorted(rows)
        for _, r0, r1, ncols, cnt in panels:
            base = r0 * (2 * n - r0 - 1) // 2
            total[base:base + cnt] = got[base:base + cnt]
            seen += cnt
            assert ncols == n
    assert seen == n * (n - 1) // 2          # the parts tile the triangle exactly once
    assert n_panels >= 4
    check(measure, total, want)


def test_pipelined_one_call_ascii_u16_and_host_counts(dg, oracle):
    from distance_b200 import api, synth
    n, width = 900, 411
    asc = synth.make_alignment(n, width=width, seed=5, ambiguity=True, mu=0.02)
    codes = synth.encode_ascii(asc)
    with dg.Engine("n_high", width) as e:
        small_pieces(e, width, chunk_records=128, panel_bytes=1 << 18)
        e.set_option(api.DG_OPT_RESULT_U16, 1)
        got, _ = e.square_pipelined(asc, input_kind=dg.DG_INPUT_ASCII, one_call=True)
        assert got.dtype == np.uint16
        check("n_high", got, oracle_run(oracle, "n_high", "square", codes))
    acgt = np.stack([oracle.count_bases(r) for r in codes]).astype(np.uint64)
    acgt[:, 0] += 3    # host-supplied counts are taken as given (EncodedFastaRecord.count_*)
    with dg.Engine("tn93", width) as e:
        small_pieces(e, width, chunk_records=128, panel_bytes=1 << 20)
        got, _ = e.square_pipelined(codes, acgt=acgt)
        check("tn93", got, oracle_run(oracle, "tn93", "square", codes, a_acgt=acgt))


def test_pipelined_heavy_ambiguity_falls_back(dg, oracle, engine):
    """An alignment full of partial ambiguity codes: the auto engine finishes on the LOP3 tiles; every engine
    still returns the oracle's counts."""
    from distance_b200 import synth
    n, width = 1500, 96
    rng = np.random.default_rng(3)
    codes = synth.random_codes(rng, n, width, p_ambig=0.9)
    with dg.Engine("raw", width) as e:
        small_pieces(e, width, chunk_records=128, panel_bytes=1 << 20)
        got, _ = e.square_pipelined(codes)
        check("raw", got, oracle_run(oracle, "raw", "square", codes))
        if engine == "auto":
            assert e.timings()["engine"] == 1


def test_pipelined_invalid_byte_and_protocol_errors(dg):
    from distance_b200 import api
    n, width = 600, 100
    good = np.full((n, width), 136, dtype=np.uint8)
    bad = good.copy()
    bad[130, 57] = 7
    bad[400, 3] = 0
    with dg.Engine("raw", width) as e:
        small_pieces(e, width, chunk_records=128, panel_bytes=1 << 16)
        with pytest.raises(dg.DistanceGpuError) as ei:
            e.square_pipelined(bad)
        assert ei.value.code == -4
        rec, site, byte = e.invalid_site()
        # chunks arrive highest records first, so the first offending record SEEN is the one in the later chunk;
        # within what has been seen the report is the lowest (record, site)
        assert (rec, site, byte) in ((400, 3, 0), (130, 57, 7))
        # the session is closed and the context is usable again
        got, _ = e.square_pipelined(good)
        assert not got.any()
        # a push that is not the chunk dg_square_next announced is refused and closes the session
        sink = api.SINK_FN(lambda u, p: 0)
        assert e.L.dg_square_begin(e.h, n, 0, None, 0, 1, sink, None) == 0
        rc = e.L.dg_square_push(e.h, C.c_void_p(good.ctypes.data), -1, 0, 128, None)
        assert rc == -1
        assert e.L.dg_square_end(e.h) == -3
        # no session open: push / next / end report DG_ERR_STATE
        lo, hi = C.c_uint64(), C.c_uint64()
        assert e.L.dg_square_next(e.h, C.byref(lo), C.byref(hi)) == -3
        # other entry points are refused while a session is open
        assert e.L.dg_square_begin(e.h, n, 0, None, 0, 1, sink, None) == 0
        with pytest.raises(dg.DistanceGpuError):
            e.load(0, good)
        assert e.L.dg_square_begin(e.h, n, 0, None, 0, 1, sink, None) == -3
        e.L.dg_square_next(e.h, C.byref(lo), C.byref(hi))
        assert hi.value == n


def test_pipelined_sars_cov_2_width(dg, oracle):
    """29,903-nt records with the bench's ambiguity mix, default chunking, against the oracle (sampled rows)."""
    from distance_b200 import api, synth
    n = 3000
    codes = synth.encode_ascii(synth.make_alignment(n, seed=20251018 + 2, ambiguity=True))
    with dg.Engine("n_high", synth.SC2_WIDTH) as e:
        e.set_option(api.DG_OPT_RESULT_U16, 1)
        e.set_option(api.DG_OPT_PIPE_CHUNK_BYTES, 8 << 20)
        got, panels = e.square_pipelined(codes)
        e.load(0, codes)
        classic = e.run_square()
    assert np.array_equal(got, classic)
    rng = np.random.default_rng(1)
    for i in rng.choice(n - 1, 12, replace=False):
        base = i * (2 * n - i - 1) // 2
        for j in rng.choice(np.arange(i + 1, n), min(8, n - 1 - i), replace=False):
            assert int(got[base + j - i - 1]) == oracle.pair_counts(codes[i], codes[j])["snp"]


@pytest.mark.parametrize("measure", ["n_high", "raw", "tn93"])
def test_overlapped_repack_keeps_the_operands(dg, oracle, measure):
    """Kernel-only runs with DG_RUN_REPACK re-pack the planes chunk by chunk in descending panel order: afterwards
    the resident operands must be what a plain load left (the next ordered run still matches the oracle)."""
    from distance_b200 import api, synth
    n, width = 2600, 200
    rng = np.random.default_rng(17)
    codes = synth.random_codes(rng, n, width, p_ambig=0.05)
    want = oracle_run(oracle, measure, "square", codes)
    with dg.Engine(measure, width) as e:
        e.set_option(api.DG_OPT_PANEL_BYTES, 1 << 20)
        e.set_option(api.DG_OPT_KEEP_CODES, 1)
        e.load(0, codes)
        for overlap in (1, 0, 1):
            e.set_option(api.DG_OPT_REPACK_OVERLAP, overlap)
            e.run_device_only(repack=True)
            e.run_device_only(api.DG_MODE_SQUARE, 1, 3, repack=True)
            check(measure, e.run_square(), want)


# ---------------------------------------------------------------------------------------------
# the two-file session: alignment 1 resident, alignment 0 pushed ascending, panels in the reference's order
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("measure", ALL)
@pytest.mark.parametrize("na,nb,width,amb", [(3, 2, 9, 0.3), (700, 150, 300, 0.1), (1100, 900, 131, 0.4)])
def test_pipelined_rect_matches_oracle_in_order(dg, oracle, measure, na, nb, width, amb):
    from distance_b200 import synth
    rng = np.random.default_rng(na * 13 + nb)
    a, b = synth.random_codes(rng, na, width, p_ambig=amb), synth.random_codes(rng, nb, width, p_ambig=amb)
    want = oracle_run(oracle, measure, "rect", a, b)
    with dg.Engine(measure, width) as e:
        small_pieces(e, width, chunk_records=128, panel_bytes=1 << 19)
        e.load(1, b)
        got, panels = e.rect_pipelined(a, one_call=(na % 2 == 1))
        rows = [p[1] for p in panels]
        assert rows == sorted(rows) and all(p[0] == 1 and p[3] == nb for p in panels)   # ascending = reference order
        check(measure, got, want)
        # alignment 0 is resident afterwards: the classic two-file run gives the same
        check(measure, e.run_rect(), want)


def test_pipelined_rect_parts_errors_and_fallback(dg, oracle, engine):
    from distance_b200 import synth
    na, nb, width = 1500, 700, 120
    rng = np.random.default_rng(5)
    a, b = synth.random_codes(rng, na, width, p_ambig=0.2), synth.random_codes(rng, nb, width, p_ambig=0.2)
    want = oracle_run(oracle, "k80", "rect", a, b)
    got = np.zeros_like(want)
    for part in range(2):
        with dg.Engine("k80", width) as e:
            small_pieces(e, width, chunk_records=256, panel_bytes=1 << 20)
            e.load(1, b)
            vals, panels = e.rect_pipelined(a, part=part, n_parts=2, one_call=False)
            pos = 0
            for _, r0, r1, ncols, cnt in panels:
                got[r0 * nb:r0 * nb + cnt] = vals[pos:pos + cnt]
                pos += cnt
    check("k80", got, want)
    with dg.Engine("n_high", width) as e:
        with pytest.raises(dg.DistanceGpuError) as ei:       # alignment 1 missing
            e.rect_pipelined(a)
        assert ei.value.code == -3
        heavy_a, heavy_b = synth.random_codes(rng, na, width, p_ambig=0.9), synth.random_codes(rng, nb, width, p_ambig=0.9)
        small_pieces(e, width, chunk_records=128, panel_bytes=1 << 20)
        e.load(1, heavy_b)
        vals, _ = e.rect_pipelined(heavy_a)                  # full of partial codes: auto finishes on the LOP3 tiles
        check("n_high", vals, oracle_run(oracle, "n_high", "rect", heavy_a, heavy_b))
        if engine == "auto":
            assert e.timings()["engine"] == 1
        bad = a.copy()
        bad[777, 5] = 3
        e.load(1, b)
        with pytest.raises(dg.DistanceGpuError) as ei:
            e.rect_pipelined(bad)
        assert ei.value.code == -4 and e.invalid_site()[:2] == (777, 5)
