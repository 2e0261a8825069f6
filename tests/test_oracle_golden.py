"""Pins the CPU oracle against every golden vector the reference's own tests hold for the
hot path.  Reference citations are /root/reference/src/<file>:<line>."""
import math

import numpy as np

TARGET = b"ATGATGATGATGCCC"  # measures.rs:202-204, fastaio.rs:344-346
QUERY = b"ATTATTATGATGCCC"   # measures.rs:206-208, fastaio.rs:348-350
CODES = [136, 24, 72, 136, 24, 72, 136, 24, 72, 136, 24, 72, 40, 40, 40]  # fastaio.rs:384-386


def test_encoding_array_table(oracle):
    # encoding.rs:7-38: exactly 32 populated entries
    a = oracle.encoding_array()
    want = {"A": 136, "G": 72, "C": 40, "T": 24, "R": 192, "M": 160, "W": 144, "S": 96, "K": 80,
            "Y": 48, "V": 224, "H": 176, "D": 208, "B": 112, "N": 240}
    for ch, v in want.items():
        assert a[ord(ch)] == v and a[ord(ch.lower())] == v
    assert a[ord("-")] == 244 and a[ord("?")] == 242
    assert int((a != 0).sum()) == 32


def test_encode(oracle):  # fastaio.rs:379-389 test_encode, :413-422 test_load_alignment
    assert oracle.encode(TARGET).tolist() == CODES


def test_encode_invalid(oracle):  # fastaio.rs:111-113
    for bad in (b"ATGXATG", b"ATGUATG", b"ATG.ATG", b"ATG ATG"):
        try:
            oracle.encode(bad)
            assert False
        except ValueError as e:
            assert e.args[0] == 3


def test_count_bases(oracle):  # fastaio.rs:358-367 / :391-400; order A,T,G,C
    assert oracle.count_bases(oracle.encode(TARGET)).tolist() == [4, 4, 4, 3]
    codes, cnt = oracle.encode_count_bases(TARGET)
    assert codes.tolist() == CODES and cnt.tolist() == [4, 4, 4, 3]


def test_count_bases_case_quirk(oracle):
    # fastaio.rs:62-65 counts encoded bytes (case-insensitive); :139-142 counts raw upper-case only
    s = b"atgATGatgATGccc"
    assert oracle.count_bases(oracle.encode(s)).tolist() == [4, 4, 4, 3]
    _, cnt = oracle.encode_count_bases(s)
    assert cnt.tolist() == [2, 2, 2, 0]


def test_get_differences(oracle):  # fastaio.rs:369-377 / :402-411
    assert oracle.get_differences(oracle.encode(TARGET), oracle.encode(QUERY)).tolist() == [2, 5]


def test_consensus(oracle):  # fastaio.rs:424-456, incl. the tie -> first of A,G,C,T
    rec, other = oracle.encode(TARGET), oracle.encode(QUERY)
    assert oracle.consensus(np.stack([rec, other])).tolist() == CODES
    assert oracle.consensus(np.stack([rec, rec])).tolist() == CODES
    assert oracle.consensus(np.stack([other, other])).tolist() == [
        136, 24, 24, 136, 24, 24, 136, 24, 72, 136, 24, 72, 40, 40, 40]


def test_snp(oracle):  # measures.rs:219-224
    assert oracle.snp(oracle.encode(TARGET), oracle.encode(QUERY)) == 2


def test_snp_consensus(oracle):  # measures.rs:226-238
    t, q = oracle.encode(TARGET), oracle.encode(QUERY)
    c = oracle.consensus(np.stack([t, q]))
    td, qd = oracle.get_differences(t, c), oracle.get_differences(q, c)
    assert oracle.snp_consensus(t, q, td, qd) == 2


def test_raw(oracle):  # measures.rs:240-245: exact f64 equality
    assert oracle.raw(oracle.encode(QUERY), oracle.encode(TARGET)) == 2.0 / 15.0


def test_jc69(oracle):  # measures.rs:247-255
    want = -0.75 * math.log(1.0 - (4.0 / 3.0) * (2.0 / 15.0))
    assert oracle.jc69(oracle.encode(QUERY), oracle.encode(TARGET)) == want


def test_k80(oracle):  # measures.rs:257-271
    P, Q = 0.0 / 15.0, 2.0 / 15.0
    want = -0.5 * math.log((1.0 - 2.0 * P - Q) * math.sqrt(1.0 - 2.0 * Q))
    assert oracle.k80(oracle.encode(QUERY), oracle.encode(TARGET)) == want


def test_tn93(oracle):  # measures.rs:273-308
    t, q = oracle.encode(TARGET), oracle.encode(QUERY)
    got = oracle.tn93(t, q, oracle.count_bases(t), oracle.count_bases(q))
    g_A, g_T, g_C, g_G = 8.0 / 30.0, 10.0 / 30.0, 6.0 / 30.0, 6.0 / 30.0
    g_R, g_Y = (8.0 + 6.0) / 30.0, (7.0 + 9.0) / 30.0
    k1 = 2.0 * g_A * g_G / g_R
    k2 = 2.0 * g_T * g_C / g_Y
    k3 = 2.0 * (g_R * g_Y - g_A * g_G * g_Y / g_R - g_T * g_C * g_R / g_Y)
    P1, P2, Q = 0.0 / 15.0, 0.0 / 15.0, (2.0 - (0.0 + 0.0)) / 15.0
    w1 = 1.0 - P1 / k1 - Q / (2.0 * g_R)
    w2 = 1.0 - P2 / k2 - Q / (2.0 * g_Y)
    w3 = 1.0 - Q / (2.0 * g_R * g_Y)
    want = -k1 * math.log(w1) - k2 * math.log(w2) - k3 * math.log(w3)
    assert got == want


def test_decimal_values(oracle):  # the same expressions in decimal
    t, q = oracle.encode(TARGET), oracle.encode(QUERY)
    assert oracle.raw(q, t) == 0.13333333333333333
    assert abs(oracle.jc69(q, t) - 0.1468084328445715) < 1e-15
    assert abs(oracle.k80(q, t) - 0.14908915389629654) < 1e-15
    assert abs(oracle.tn93(t, q, oracle.count_bases(t), oracle.count_bases(q)) - 0.1494325473614665) < 1e-15


FASTA_1 = [("seq1", b"ATGATG"), ("seq2", b"ATGATC")]  # lib.rs:906-910
FASTA_2 = [("seqA", b"ATGATG")]                       # lib.rs:912-914


def _aln(oracle, recs):
    return oracle.Alignment(np.stack([oracle.encode(s) for _, s in recs]))


def test_pair_order_square(oracle):
    # lib.rs:667-786: n=4 -> (0,1),(0,2),(0,3),(1,2),(1,3),(2,3); we check the order through
    # distinct n_high values: seq k differs from seq 0 at k sites, pair (i,j) -> j - i
    seqs = [b"AAAAAA", b"CAAAAA", b"CCAAAA", b"CCCAAA"]
    a = oracle.Alignment(np.stack([oracle.encode(s) for s in seqs]))
    out, _ = oracle.run("n_high", "square", a)
    assert out.tolist() == [1, 2, 3, 1, 2, 1]


def test_pair_order_rectangle(oracle):  # lib.rs:817-896: (0,0),(0,1),(1,0),(1,1)
    a = oracle.Alignment(np.stack([oracle.encode(s) for s in (b"AAAA", b"CCAA")]))
    b = oracle.Alignment(np.stack([oracle.encode(s) for s in (b"AAAA", b"CAAA")]))
    out, _ = oracle.run("n_high", "rect", a, b)
    assert out.tolist() == [0, 1, 2, 1]


def test_integration_1_tsv(oracle):  # lib.rs:918-940 (-m n, one file); threads 1 and 2 (:968-999)
    a = _aln(oracle, FASTA_1)
    oracle.prepare("n", [a])
    for threads in (1, 2):
        out, _ = oracle.run("n", "square", a, threads=threads)
        ids = [i for i, _ in FASTA_1]
        assert oracle.tsv(ids, ids, "square", out, True) == "sequence1\tsequence2\tdistance\nseq1\tseq2\t1\n"


def test_integration_2_tsv(oracle):  # lib.rs:1003-1019 (-m n_high, stream)
    a, b = _aln(oracle, FASTA_1), _aln(oracle, FASTA_2)
    out, _ = oracle.run("n_high", "stream", a, b)
    got = oracle.tsv([i for i, _ in FASTA_1], [i for i, _ in FASTA_2], "stream", out, True)
    assert got == "sequence1\tsequence2\tdistance\nseq1\tseqA\t0\nseq2\tseqA\t1\n"


def test_integration_3_tsv(oracle):  # lib.rs:1069-1085 and reversed :1134-1148
    a, b = _aln(oracle, FASTA_1), _aln(oracle, FASTA_2)
    out, _ = oracle.run("n_high", "rect", a, b)
    got = oracle.tsv([i for i, _ in FASTA_1], [i for i, _ in FASTA_2], "rect", out, True)
    assert got == "sequence1\tsequence2\tdistance\nseq1\tseqA\t0\nseq2\tseqA\t1\n"
    out, _ = oracle.run("n_high", "rect", b, a)
    got = oracle.tsv([i for i, _ in FASTA_2], [i for i, _ in FASTA_1], "rect", out, True)
    assert got == "sequence1\tsequence2\tdistance\nseqA\tseq1\t0\nseqA\tseq2\t1\n"


def test_float_text(oracle):  # lib.rs:631 `{:.12}`: Rust Display for f64
    assert oracle.format_float12(2.0 / 15.0) == "0.133333333333"
    assert oracle.format_float12(-0.0) == "-0.000000000000"
    assert oracle.format_float12(0.0) == "0.000000000000"
    assert oracle.format_float12(float("nan")) == "NaN"
    assert oracle.format_float12(float("inf")) == "inf"
    assert oracle.format_float12(float("-inf")) == "-inf"


def test_special_values(oracle):
    e = oracle.encode
    # identical pair: jc69 / k80 give -0.0, tn93 normalises to +0.0 (measures.rs:187-190)
    s = e(b"ACGTACGT")
    assert math.copysign(1.0, oracle.jc69(s, s)) == -1.0 and oracle.jc69(s, s) == 0.0
    assert math.copysign(1.0, oracle.k80(s, s)) == -1.0
    c = oracle.count_bases(s)
    v = oracle.tn93(s, s, c, c)
    assert v == 0.0 and math.copysign(1.0, v) == 1.0
    # no comparable sites -> NaN for every float measure
    n = e(b"NNNN")
    assert math.isnan(oracle.raw(n, n)) and math.isnan(oracle.jc69(n, n)) and math.isnan(oracle.k80(n, n))
    # p == 3/4 exactly -> +inf; p > 3/4 -> NaN (measures.rs:76)
    assert oracle.jc69(e(b"AAAA"), e(b"CCCA")) == math.inf
    assert math.isnan(oracle.jc69(e(b"AAAA"), e(b"CCCC")))


def test_n_equals_n_high_random(oracle):
    # snp_consensus == snp for every input (any ACGT consensus)
    rng = np.random.default_rng(7)
    codes = np.array([136, 72, 40, 24, 192, 160, 144, 96, 80, 48, 224, 176, 208, 112, 240, 244, 242], np.uint8)
    p = np.array([20, 20, 20, 20] + [1] * 10 + [6, 3, 1], float)
    p /= p.sum()
    for trial in range(20):
        n, w = int(rng.integers(2, 9)), int(rng.integers(1, 80))
        a = oracle.Alignment(rng.choice(codes, size=(n, w), p=p))
        oracle.prepare("n", [a])
        got, _ = oracle.run("n", "square", a)
        want, _ = oracle.run("n_high", "square", a)
        assert got.tolist() == want.tolist()


def test_committed_tsv_fixtures_are_what_the_oracle_writes():
    """tests/golden/golden_*.tsv (what the GPU CLI is compared with in tests/test_cli_gpu.py) are regenerated from the
    oracle and must be byte-identical to the committed files."""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    texts = mg.golden_texts()
    assert len(texts) == 2 + 3 * 6
    for name, data in texts.items():
        with open(os.path.join(here, name), "rb") as f:
            assert f.read() == data, name
    sq = texts["golden_square_jc69.tsv"].decode()
    assert "-0.000000000000" in sq and "NaN" in sq and "inf" in sq   # the special values are exercised
    assert "\t0.000000000000\n" in texts["golden_square_tn93.tsv"].decode()
