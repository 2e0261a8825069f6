"""GPU parity tests: the CUDA path, called through the C ABI (libdistance_gpu.so), against the CPU
oracle on the same seeded inputs.  Integer counts and n / n_high are bit-exact; raw / jc69 / k80 /
tn93 agree within 1e-12 relative (north_star's tolerance) with NaN / +-inf / -0.0 in exactly the
same places."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12  # BASELINE.json north_star: "raw/jc69/k80/tn93 must agree within 1e-12 relative"
ALL = ["n", "n_high", "raw", "jc69", "k80", "tn93"]


@pytest.fixture(autouse=True, params=["lop3", "tc", "fp4", "auto", "fp4_nosplit"])
def engine(request, monkeypatch):
    """Every test runs on every count engine and on the automatic choice: DG_ENGINE overrides DG_OPT_ENGINE at
    dg_create (1 = LOP3+POPC bit-plane tiles, 2 = tcgen05 kind::i8 GEMM, 3 = tcgen05 kind::mxf4 GEMM with unit
    scales, 0 = auto).  Small launches of the tensor engines split the K range over several CTA pairs (partial sums
    added into a zeroed scratch); "fp4_nosplit" (DG_KSPLIT=1) keeps these small cases on the one-pass epilogues that the
    large configs use (16-bit / modulo-2^16 scratch, direct counts)."""
    monkeypatch.setenv("DG_ENGINE", {"lop3": "1", "tc": "2", "fp4": "3", "auto": "0", "fp4_nosplit": "3"}[request.param])
    if request.param == "fp4_nosplit":
        monkeypatch.setenv("DG_KSPLIT", "1")
        return "fp4"
    return request.param


@pytest.fixture(scope="module")
def dg():
    import distance_b200 as d
    d.load_library()
    assert d.device_count() >= 1, "no CUDA device: the gpu tests must run on the B200 box"
    return d


def assert_float_parity(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"NaN positions differ at {np.flatnonzero(nan_g != nan_w)[:10]}"
    inf_g, inf_w = np.isinf(got), np.isinf(want)
    assert np.array_equal(inf_g, inf_w), f"inf positions differ at {np.flatnonzero(inf_g != inf_w)[:10]}"
    assert np.array_equal(got[inf_g], want[inf_w])  # same sign of infinity
    fin = ~(nan_w | inf_w)
    g, w = got[fin], want[fin]
    zero = w == 0.0
    assert np.array_equal(g[zero], w[zero]) and np.array_equal(np.signbit(g[zero]), np.signbit(w[zero])), \
        "zeros (incl. the sign of -0.0) differ"
    nz = ~zero
    rel = np.abs(g[nz] - w[nz]) / np.abs(w[nz])
    assert rel.size == 0 or rel.max() <= REL_TOL, f"max rel err {rel.max():.3e}"


def check(measure, got, want):
    if measure in ("n", "n_high"):
        assert np.array_equal(got.astype(np.int64), want), \
            f"first mismatch at {np.flatnonzero(got.astype(np.int64) != want)[:10]}"
    else:
        assert_float_parity(got, want)


def oracle_run(oracle, measure, mode, a_codes, b_codes=None, a_acgt=None, b_acgt=None):
    a = oracle.Alignment(a_codes, a_acgt)
    b = None if b_codes is None else oracle.Alignment(b_codes, b_acgt)
    alns = [a] + ([b] if b is not None else [])
    # -s mode: the consensus comes from the loaded file only (lib.rs:224, fastaio.rs:232-240)
    oracle.prepare(measure, alns, consensus_from=[a] if mode == "stream" else None)
    out, _ = oracle.run(measure, mode, a, b, threads=8)
    return out


def ref_planes(codes, wp):
    """numpy restatement of the plane layout documented in kernels.cuh."""
    n, w = codes.shape
    pad = np.full((n, wp * 32), 240, dtype=np.uint8)
    pad[:, :w] = codes

    def pack(bits):
        b = bits.reshape(n, wp, 32).astype(np.uint64)
        return (b << np.arange(32, dtype=np.uint64)).sum(axis=2).astype(np.uint32)

    core = np.stack([pack((pad & m) != 0) for m in (128, 64, 32, 16)], axis=2)
    pur, pyr = (pad & 55) == 0, (pad & 199) == 0
    aux = np.stack([pack((pad & 8) != 0), pack(pyr), pack(pur | pyr), np.zeros((n, wp), np.uint32)], axis=2)
    return core, aux


# ---------------------------------------------------------------------------------------------
# pack_planes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,width", [(1, 1), (3, 31), (5, 32), (7, 33), (130, 257), (40, 1000)])
def test_pack_planes_and_base_counts(dg, oracle, n, width):
    from distance_b200 import synth
    rng = np.random.default_rng(n * 1000 + width)
    codes = synth.random_codes(rng, n, width, p_ambig=0.3)
    with dg.Engine("tn93", width) as e:
        e.load(0, codes)
        core, aux, acgt = e.debug_planes(0)
    rcore, raux = ref_planes(codes, core.shape[1])
    assert np.array_equal(core, rcore)
    assert np.array_equal(aux, raux)
    want = np.stack([oracle.count_bases(r) for r in codes])  # fastaio.rs:53-66
    assert np.array_equal(acgt, want)


def test_pack_ascii_matches_paradis_and_case(dg, oracle):
    from distance_b200 import synth
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACGTacgtRYMWSKVHDBNrymwskvhdbn-?", dtype=np.uint8)
    ascii_codes = letters[rng.integers(0, letters.size, size=(9, 211))]
    codes = np.stack([oracle.encode(r.tobytes()) for r in ascii_codes])
    assert np.array_equal(codes, synth.encode_ascii(ascii_codes))
    with dg.Engine("tn93", 211) as e:
        e.load(0, ascii_codes, input_kind=dg.DG_INPUT_ASCII)
        core_a, aux_a, acgt_a = e.debug_planes(0)
        e.load(0, codes)
        core_p, aux_p, acgt_p = e.debug_planes(0)
    assert np.array_equal(core_a, core_p) and np.array_equal(aux_a, aux_p)
    assert np.array_equal(acgt_a, acgt_p)  # loaded files: case-insensitive count_bases


@pytest.mark.parametrize("kind", ["ascii", "paradis"])
def test_invalid_byte_is_reported(dg, kind):
    width = 100
    good = np.full((4, width), ord("A") if kind == "ascii" else 136, dtype=np.uint8)
    bad = good.copy()
    bad[2, 57] = ord("X") if kind == "ascii" else 7
    bad[3, 3] = ord("U") if kind == "ascii" else 0
    with dg.Engine("raw", width) as e:
        with pytest.raises(dg.DistanceGpuError) as ei:
            e.load(0, bad, input_kind=dg.DG_INPUT_ASCII if kind == "ascii" else dg.DG_INPUT_PARADIS)
        assert ei.value.code == -4
        rec, site, byte = e.invalid_site()
        assert (rec, site) == (2, 57)  # first in record order, like the reference's sequential scan
        assert byte == bad[2, 57]
        e.load(0, good, input_kind=dg.DG_INPUT_ASCII if kind == "ascii" else dg.DG_INPUT_PARADIS)


# ---------------------------------------------------------------------------------------------
# integer counts of every site loop
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("measure", ["n_high", "raw", "k80", "tn93"])
@pytest.mark.parametrize("n,width", [(2, 1), (9, 63), (70, 300), (150, 1025)])
def test_counts_bit_exact(dg, oracle, measure, n, width):
    from distance_b200 import synth
    rng = np.random.default_rng(sum(map(ord, measure)) * 7919 + n * 31 + width)
    codes = synth.random_codes(rng, n, width, p_ambig=0.35)
    with dg.Engine(measure, width) as e:
        e.load(0, codes)
        got = e.debug_counts(0, 0)
    sel = [(i, j) for i in range(n) for j in range(n)]
    if len(sel) > 600:
        sel = [sel[k] for k in rng.choice(len(sel), 600, replace=False)]
    for i, j in sel:
        c = oracle.pair_counts(codes[i], codes[j])
        if measure == "n_high":
            want = [c["snp"], 0, 0, 0]
        elif measure == "raw":
            want = [c["raw_n"], c["raw_d"] - c["raw_n"], 0, 0]
        elif measure == "k80":
            want = [c["k80_L"] - c["k80_ts"] - c["k80_tv"], c["k80_ts"] + c["k80_tv"], c["k80_tv"], 0]
        else:
            want = [c["tn93_L"], c["tn93_d"], c["tn93_P1"], c["tn93_P2"]]
        assert got[i, j].tolist() == want, (i, j)


# ---------------------------------------------------------------------------------------------
# results in reference order: square / rect / stream, all six measures
# ---------------------------------------------------------------------------------------------
def test_reference_golden_pair(dg, oracle):
    # measures.rs:202-208 fixture; SURVEY 8c(5)
    t, q = oracle.encode(b"ATGATGATGATGCCC"), oracle.encode(b"ATTATTATGATGCCC")
    codes = np.stack([q, t])
    want = {"n": 2, "n_high": 2, "raw": 2.0 / 15.0, "jc69": 0.1468084328445715,
            "k80": 0.14908915389629654, "tn93": 0.1494325473614665}
    for m, w in want.items():
        with dg.Engine(m, 15) as e:
            e.load(0, codes)
            got = e.run_square()
        assert got.shape == (1,)
        if m in ("n", "n_high"):
            assert int(got[0]) == w
        elif m == "raw":
            assert got[0] == w  # one IEEE division: exact
        else:
            assert abs(got[0] - w) <= REL_TOL * w


@pytest.mark.parametrize("measure", ALL)
@pytest.mark.parametrize("n,width,amb", [(2, 7, 0.3), (33, 100, 0.3), (200, 1500, 0.02), (260, 333, 0.5)])
def test_square_matches_oracle(dg, oracle, measure, n, width, amb):
    from distance_b200 import synth
    rng = np.random.default_rng(n + width)
    codes = synth.random_codes(rng, n, width, p_ambig=amb)
    with dg.Engine(measure, width) as e:
        e.load(0, codes)
        got = e.run_square()
    check(measure, got, oracle_run(oracle, measure, "square", codes))


@pytest.mark.parametrize("measure", ALL)
def test_rect_matches_oracle_both_orders(dg, oracle, measure):
    from distance_b200 import synth
    rng = np.random.default_rng(99)
    a = synth.random_codes(rng, 70, 400, p_ambig=0.2)
    b = synth.random_codes(rng, 131, 400, p_ambig=0.2)
    with dg.Engine(measure, 400) as e:
        e.load(0, a)
        e.load(1, b)
        got = e.run_rect()
        check(measure, got, oracle_run(oracle, measure, "rect", a, b))
        e.load(0, b)  # reversed inputs: lib.rs:1134-1153
        e.load(1, a)
        got = e.run_rect()
        check(measure, got, oracle_run(oracle, measure, "rect", b, a))


@pytest.mark.parametrize("measure", ALL)
def test_stream_matches_oracle(dg, oracle, measure, engine):
    from distance_b200 import synth
    rng = np.random.default_rng(1234)
    loaded = synth.random_codes(rng, 45, 500, p_ambig=0.2)
    streamed = synth.random_codes(rng, 210, 500, p_ambig=0.2)
    want = oracle_run(oracle, measure, "stream", loaded, streamed)
    for batch in (1, 7, 64, 1000):
        with dg.Engine(measure, 500) as e:
            e.load(0, loaded)
            chunks = [streamed[i:i + batch] for i in range(0, streamed.shape[0], batch)]
            got = e.stream(chunks, max_batch=64)
            rows = [(p[1], p[2]) for p in e.last_panels]
        assert rows[0][0] == 0 and rows[-1][1] == 210
        assert all(rows[k][1] == rows[k + 1][0] for k in range(len(rows) - 1))  # streamed order
        check(measure, got, want)


def test_stream_tn93_uppercase_count_quirk(dg, oracle):
    """-s + tn93: streamed records count raw 'A','T','G','C' only (fastaio.rs:139-142) while the loaded
    file counts encoded bytes (fastaio.rs:62-65).  ASCII input reproduces it on device."""
    rng = np.random.default_rng(8)
    letters = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    loaded_ascii = letters[rng.integers(0, 9, size=(6, 300))]
    streamed_ascii = letters[rng.integers(0, 9, size=(20, 300))]
    loaded = np.stack([oracle.encode(r.tobytes()) for r in loaded_ascii])
    enc = [oracle.encode_count_bases(r.tobytes()) for r in streamed_ascii]
    streamed = np.stack([c for c, _ in enc])
    s_acgt = np.stack([k for _, k in enc])
    want = oracle_run(oracle, "tn93", "stream", loaded, streamed, b_acgt=s_acgt)
    plain = oracle_run(oracle, "tn93", "stream", loaded, streamed)
    assert not np.array_equal(want, plain)  # the quirk is visible on this input
    with dg.Engine("tn93", 300) as e:
        e.load(0, loaded_ascii, input_kind=dg.DG_INPUT_ASCII)
        got = e.stream([streamed_ascii[:9], streamed_ascii[9:]], input_kind=dg.DG_INPUT_ASCII)
        check("tn93", got, want)
        # Paradis input + host-supplied counts gives the same
        got2 = e.stream([streamed[:9], streamed[9:]], acgt_batches=[s_acgt[:9], s_acgt[9:]])
        check("tn93", got2, want)


@pytest.mark.parametrize("measure", ["n_high", "raw"])
def test_no_partial_codes_direct_path(dg, oracle, measure):
    """Alignments without partial ambiguity codes (only A,C,G,T,N,-,?) need no both-partial repair: the
    tensor engine's n / n_high epilogue stores DIFF straight from TMEM.  Square, rect and stream."""
    rng = np.random.default_rng(17)
    pool = np.array([136, 72, 40, 24, 240, 244, 242], np.uint8)
    a = pool[rng.choice(7, size=(300, 700), p=[.22, .22, .22, .22, .06, .03, .03])]
    b = pool[rng.choice(7, size=(77, 700), p=[.22, .22, .22, .22, .06, .03, .03])]
    with dg.Engine(measure, 700) as e:
        e.load(0, a)
        check(measure, e.run_square(), oracle_run(oracle, measure, "square", a))
        e.load(1, b)
        check(measure, e.run_rect(), oracle_run(oracle, measure, "rect", a, b))
        got = e.stream([b[:50], b[50:]], max_batch=64)
        check(measure, got, oracle_run(oracle, measure, "stream", a, b))


@pytest.mark.parametrize("amb", [0.0, 0.3])
def test_uint16_result_panels(dg, oracle, amb):
    """DG_OPT_RESULT_U16: n / n_high panels as uint16 (lossless: a count never exceeds the width)."""
    from distance_b200 import api, synth
    rng = np.random.default_rng(23)
    if amb:
        a, b = synth.random_codes(rng, 1100, 900, p_ambig=amb), synth.random_codes(rng, 64, 900, p_ambig=amb)
    else:
        pool = np.array([136, 72, 40, 24, 240], np.uint8)
        a, b = pool[rng.integers(0, 5, size=(1100, 900))], pool[rng.integers(0, 5, size=(64, 900))]
    for measure in ("n", "n_high"):
        with dg.Engine(measure, 900) as e:
            e.set_option(api.DG_OPT_RESULT_U16, 1)
            e.set_option(api.DG_OPT_PANEL_BYTES, 512 * 1100 * 2)
            e.load(0, a)
            got = e.run_square()
            assert got.dtype == np.uint16 and len(e.last_panels) >= 2
            check(measure, got, oracle_run(oracle, measure, "square", a))
            e.load(1, b)
            check(measure, e.run_rect(), oracle_run(oracle, measure, "rect", a, b))
            check(measure, e.stream([b]), oracle_run(oracle, measure, "stream", a, b))
    with dg.Engine("n_high", 70000) as e:
        with pytest.raises(dg.DistanceGpuError):
            e.set_option(api.DG_OPT_RESULT_U16, 1)   # a count could exceed 65535


def test_special_values(dg, oracle):
    """-0.0 for identical pairs (jc69/k80), +0.0 (tn93), NaN without comparable sites, +inf at p == 3/4."""
    e = oracle.encode
    seqs = [b"ACGTACGT", b"ACGTACGT", b"NNNNNNNN", b"AAAANNNN", b"CCCANNNN", b"CCCCNNNN", b"RYRYRYRY", b"----????"]
    codes = np.stack([e(s) for s in seqs])
    for m in ALL:
        with dg.Engine(m, 8) as eng:
            eng.load(0, codes)
            got = eng.run_square()
        want = oracle_run(oracle, m, "square", codes)
        check(m, got, want)
        if m == "jc69":
            assert got[0] == 0.0 and math.copysign(1, got[0]) == -1.0
            assert np.isinf(want).any() and np.isnan(want).any()
        if m == "tn93":
            assert got[0] == 0.0 and math.copysign(1, got[0]) == 1.0


@pytest.mark.parametrize("measure", ["n_high", "jc69", "tn93"])
def test_panels_and_parts_cover_the_triangle(dg, oracle, measure):
    """Small panels: many launches, ordered sink; dg_run_part shards are disjoint and complete."""
    from distance_b200 import synth
    from distance_b200 import api
    rng = np.random.default_rng(3)
    n, width = 2800, 100
    codes = synth.random_codes(rng, n, width, p_ambig=0.1)
    want = oracle_run(oracle, measure, "square", codes)
    with dg.Engine(measure, width) as e:
        e.set_option(api.DG_OPT_PANEL_BYTES, 512 * n * (4 if measure == "n_high" else 8))
        e.load(0, codes)
        got = e.run_square()
        assert len(e.last_panels) >= 4
        check(measure, got, want)
        # two ranks, no collective: interleave their panels back into global order
        parts = e.run_part(api.DG_MODE_SQUARE, 0, 2) + e.run_part(api.DG_MODE_SQUARE, 1, 2)
        parts.sort(key=lambda t: t[0])
        assert parts[0][0] == 0 and parts[-1][1] == n - 1
        check(measure, np.concatenate([p[2] for p in parts]), want)


def test_load_from_device_memory(dg, oracle):
    """dg_load_resident_device: the codes already sit in HBM (a torch tensor here; a NCCL all-gather in bench.py)."""
    import torch
    from distance_b200 import synth
    rng = np.random.default_rng(41)
    a = synth.random_codes(rng, 300, 777, p_ambig=0.2)
    t = torch.from_numpy(a).cuda()
    torch.cuda.synchronize()
    for measure in ("n_high", "tn93"):
        with dg.Engine(measure, 777) as e:
            e.load_device(0, t.data_ptr(), 0, 300)
            check(measure, e.run_square(), oracle_run(oracle, measure, "square", a))
    bad = a.copy()
    bad[7, 5] = 3
    tb = torch.from_numpy(bad).cuda()
    torch.cuda.synchronize()
    with dg.Engine("raw", 777) as e:
        with pytest.raises(dg.DistanceGpuError):
            e.load_device(0, tb.data_ptr(), 0, 300)
        assert e.invalid_site() == (7, 5, 3)


def test_single_record_and_errors(dg):
    with dg.Engine("raw", 10) as e:
        e.load(0, np.full((1, 10), 136, np.uint8))
        assert e.run_square().shape == (0,)  # generate_pairs_square(1): no pairs (lib.rs:511)
        with pytest.raises(dg.DistanceGpuError):
            e.run_rect()  # alignment 1 not loaded
    with pytest.raises(dg.DistanceGpuError):
        dg.Engine("raw", 0)


# ---------------------------------------------------------------------------------------------
# full alignment width (29,903) and BASELINE sizes through size-independent properties
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("measure", ALL)
def test_sars_cov_2_width_against_oracle(dg, oracle, measure):
    from distance_b200 import synth
    asc = synth.make_alignment(300, seed=20251018 + 2, ambiguity=True)
    codes = synth.encode_ascii(asc)
    with dg.Engine(measure, synth.SC2_WIDTH) as e:
        e.load(0, codes)
        got = e.run_square()
    check(measure, got, oracle_run(oracle, measure, "square", codes))


def test_config2_full_size_properties(dg, oracle, engine):
    """BASELINE config 2 (n_high, 20,000 x 29,903, 1% ambiguity/gaps): rows sampled against the
    oracle, symmetry square-vs-rect on a block, and a checksum that is independent of panel size."""
    from distance_b200 import api, synth
    if engine == "auto":
        pytest.skip("covered by the explicit engines")
    n = 20000
    codes = synth.encode_ascii(synth.make_alignment(n, seed=20251018 + 2, ambiguity=True))
    with dg.Engine("n_high", synth.SC2_WIDTH) as e:
        e.load(0, codes)
        got = e.run_square()
        assert got.shape[0] == n * (n - 1) // 2
        total = int(got.astype(np.uint64).sum())
        e.set_option(api.DG_OPT_PANEL_BYTES, 16 << 20)
        got2 = e.run_square()
        assert int(got2.astype(np.uint64).sum()) == total and np.array_equal(got, got2)

    def off(i):
        return i * (2 * n - i - 1) // 2

    for i in (0, 1, 7777, 19998):
        want = np.array([oracle.snp(codes[i], codes[j]) for j in range(i + 1, min(n, i + 1 + 400))])
        assert np.array_equal(got[off(i):off(i) + want.shape[0]].astype(np.int64), want)
    # symmetry: d(i,j) from the square run == rect run of block B x block A (transposed)
    a, b = codes[100:228], codes[15000:15100]
    with dg.Engine("n_high", synth.SC2_WIDTH) as e:
        e.load(0, b)
        e.load(1, a)
        rect = e.run_rect().reshape(100, 128)
    for ii in (0, 64, 127):
        i = 100 + ii
        assert np.array_equal(got[off(i) + (15000 - i - 1): off(i) + (15100 - i - 1)], rect[:, ii])


def test_auto_engine_choice(dg, oracle, engine):
    """auto: tensor cores for ordinary alignments, LOP3 tiles when nearly every site is a partial
    ambiguity code (the both-partial correction would dominate); results identical either way."""
    from distance_b200 import synth
    if engine != "auto":
        pytest.skip("auto only")
    rng = np.random.default_rng(77)
    normal = synth.encode_ascii(synth.make_alignment(200, width=2000, seed=5, ambiguity=True))
    partial = np.array([192, 160, 144, 96, 80, 48, 224, 176, 208, 112], np.uint8)[rng.integers(0, 10, size=(200, 2000))]
    for codes, want_engine in ((normal, 3), (partial, 1)):
        for measure in ("n_high", "raw"):
            with dg.Engine(measure, 2000) as e:
                e.load(0, codes)
                got = e.run_square()
                assert e.timings()["engine"] == want_engine
            check(measure, got, oracle_run(oracle, measure, "square", codes))
    with dg.Engine("tn93", 2000) as e:  # k80 / tn93 need no correction: always tensor cores
        e.load(0, partial)
        got = e.run_square()
        assert e.timings()["engine"] == 3
    check("tn93", got, oracle_run(oracle, "tn93", "square", partial))


def test_multi_gpu_single_process_matches_oracle(dg, oracle, engine):
    """One process driving every visible device: planes replicated, panels dealt round-robin, the sink
    still sees them in global order.  Skipped on a one-GPU box."""
    from distance_b200 import api, synth
    ndev = dg.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    rng = np.random.default_rng(21)
    n, width = max(2100, 1100 * ndev), 300   # more panels (of >= 512 rows) than devices: some device takes two
    codes = synth.random_codes(rng, n, width, p_ambig=0.1)
    for measure in ("n_high", "tn93"):
        want = oracle_run(oracle, measure, "square", codes)
        with dg.Engine(measure, width, gpus=list(range(ndev))) as e:
            e.set_option(api.DG_OPT_PANEL_BYTES, 512 * n * (4 if measure == "n_high" else 8))
            e.load(0, codes)
            got = e.run_square()
            assert len(e.last_panels) >= ndev       # every device takes a panel (the planner may merge small ones)
            check(measure, got, want)
            loaded, streamed = codes[:40], codes[40:400]
            e.load(0, loaded)
            got = e.stream([streamed[i:i + 50] for i in range(0, 360, 50)], max_batch=64)
        check(measure, got, oracle_run(oracle, measure, "stream", loaded, streamed))
