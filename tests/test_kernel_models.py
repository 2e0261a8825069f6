"""Host-side restatements of two small device algorithms of distance_b200/csrc/tc_engine.cuh whose index / bit arithmetic is
easy to get wrong, checked exhaustively or on random data.  No GPU needed: these pin the arithmetic the kernels implement
(the kernels themselves are checked against the oracle by the `-m gpu` suite).

* `nib_word_to_codes` / `nib4_to_codes`: DG_INPUT_NIBBLE rows (two sites per byte, low nibble = the even site) expanded to
  Paradis codes with byte-SIMD integer operations inside `pack_ops_kernel<..., NIB = true>`;
* `pp_scan_chunk_kernel`: the coalesced three-level exclusive scan of a session chunk's per-site counts (segments of
  16 x 1,024 sites, warp w owns the 32-site groups k * 32 + w)."""
import numpy as np
import pytest

M32 = 0xFFFFFFFF


def nib4_to_codes(m: int) -> int:
    """mirror of tc::nib4_to_codes: four nibbles (one per byte) -> four Paradis codes"""
    t = m & (((m | 0x10101010) - 0x01010101) & M32)
    many = ((t + 0x7F7F7F7F) & M32) & 0x80808080
    some = ((m + 0x7F7F7F7F) & M32) & 0x80808080
    return ((m << 4) & M32) | ((some & ~many & M32) >> 4)


def byte_perm(a: int, b: int, sel: int) -> int:
    """__byte_perm for selectors without the sign-replication bit"""
    src = [(a >> (8 * i)) & 255 for i in range(4)] + [(b >> (8 * i)) & 255 for i in range(4)]
    return sum(src[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))


def nib_word_to_codes(x: int):
    lo, hi = x & 0x0F0F0F0F, (x >> 4) & 0x0F0F0F0F
    return nib4_to_codes(byte_perm(lo, hi, 0x5140)), nib4_to_codes(byte_perm(lo, hi, 0x7362))


def code_of_nibble(m: int) -> int:
    """encoding.rs:7-38 read backwards: possibility nibble -> Paradis byte (known bit = exactly one possibility); 0 is invalid"""
    return (m << 4) | (8 if bin(m).count("1") == 1 else 0)


def test_every_nibble_value_maps_to_its_paradis_code():
    for m in range(16):
        for pos in range(4):
            word = sum((m if b == pos else 15) << (8 * b) for b in range(4))
            got = (nib4_to_codes(word) >> (8 * pos)) & 255
            assert got == code_of_nibble(m), (m, pos, hex(got))
    assert code_of_nibble(0) == 0            # nibble 0 becomes the invalid byte 0, which the pack kernel reports
    assert code_of_nibble(15) == 0xF0        # N-like
    assert [code_of_nibble(m) for m in (8, 4, 2, 1)] == [136, 72, 40, 24]   # A, G, C, T


def test_nibble_word_expansion_keeps_the_site_order():
    rng = np.random.default_rng(11)
    for x in [0, M32, 0x01234567, 0x89ABCDEF] + [int(v) for v in rng.integers(0, 1 << 32, 5000)]:
        c0, c1 = nib_word_to_codes(x)
        got = [(c0 >> (8 * i)) & 255 for i in range(4)] + [(c1 >> (8 * i)) & 255 for i in range(4)]
        want = []
        for b in range(4):
            byte = (x >> (8 * b)) & 255
            want += [code_of_nibble(byte & 15), code_of_nibble(byte >> 4)]   # low nibble = the even site
        assert got == want, hex(x)


def test_pack_nibbles_round_trips_through_the_device_expansion():
    from distance_b200 import api, synth
    rng = np.random.default_rng(3)
    codes = synth.random_codes(rng, 7, 37, p_ambig=0.3)
    nib = api.pack_nibbles(codes)
    for r in range(codes.shape[0]):
        row = bytes(nib[r]) + b"\xff" * 4
        out = []
        for w0 in range(0, len(nib[r]), 4):
            x = int.from_bytes(row[w0:w0 + 4], "little")
            c0, c1 = nib_word_to_codes(x)
            out += [(c0 >> (8 * i)) & 255 for i in range(4)] + [(c1 >> (8 * i)) & 255 for i in range(4)]
        got = np.array(out[:codes.shape[1]], dtype=np.uint8)
        # the nibble form drops what no measure reads: N, '-' and '?' all travel as 15 (-> 0xF0)
        want = np.where((codes[r] & 0xF0) == 0xF0, 0xF0, codes[r])
        assert np.array_equal(got, want)


def scan_chunk_model(site_cnt: np.ndarray, base: int, G: int = 16):
    """mirror of pp_scan_chunk_kernel's index arithmetic: off[0] = base, off[1 + i] = base + sum(site_cnt[:i]); returns
    (off, running total)"""
    width = len(site_cnt)
    off = np.zeros(width + 1, np.uint32)
    off[0] = base
    s_run = base
    warp = np.arange(32)[:, None]
    lane = np.arange(32)[None, :]
    for seg0 in range(0, width, G * 1024):
        gsum = np.zeros(G * 32, np.uint32)
        q = np.zeros((G, 32, 32), np.uint32)
        for k in range(G):
            i = seg0 + (k * 32 + warp) * 32 + lane
            c = np.where(i < width, site_cnt[np.minimum(i, width - 1)], 0).astype(np.uint32)
            incl = np.cumsum(c, axis=1, dtype=np.uint32)          # the warp's shuffle scan
            q[k] = incl - c
            gsum[k * 32 + np.arange(32)] = incl[:, 31]            # lane 31 of warp w -> gsum[k * 32 + w]
        wtot = np.zeros(32, np.uint32)
        for w in range(G):                                        # warp w < G scans group totals [32 w, 32 w + 32)
            v = gsum[w * 32:(w + 1) * 32].copy()
            inc = np.cumsum(v, dtype=np.uint32)
            gsum[w * 32:(w + 1) * 32] = inc - v
            wtot[w] = inc[31]
        v = np.where(np.arange(32) < G, wtot, 0).astype(np.uint32)   # warp 0 scans the G block totals
        inc = np.cumsum(v, dtype=np.uint32)
        wex, seg_total = inc - v, int(inc[31])
        for k in range(G):
            i = seg0 + (k * 32 + warp) * 32 + lane
            ok = i < width
            val = (s_run + wex[k] + gsum[k * 32 + warp] + q[k]).astype(np.uint32)
            off[1 + i[ok]] = val[ok]
        s_run += seg_total
    return off, s_run


@pytest.mark.parametrize("width", [1, 5, 31, 32, 33, 1023, 1024, 1025, 16383, 16384, 16385, 29903, 70000])
def test_chunk_scan_index_arithmetic(width):
    rng = np.random.default_rng(width)
    c = rng.integers(0, 6, width).astype(np.uint32)
    off, total = scan_chunk_model(c, base=1000)
    want = 1000 + np.concatenate([[0], np.cumsum(c)[:-1]])
    assert off[0] == 1000
    assert np.array_equal(off[1:], want.astype(np.uint32))
    assert total == 1000 + int(c.sum())
