"""End-to-end: the `distance` binary on a GPU, TSV text compared byte for byte with the oracle's
gather_write text (lib.rs:612-644) for int measures and value-wise (1e-12 rel) for float measures."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "distance_b200", "_bin", "distance")
ALL = ["n", "n_high", "raw", "jc69", "k80", "tn93"]


@pytest.fixture(scope="module", autouse=True)
def build_cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", ROOT, "cli"], stdout=subprocess.DEVNULL)


def run(args, stdin=None):
    p = subprocess.run([CLI] + args, input=stdin, capture_output=True, timeout=600)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def fasta_bytes(names, ascii_rows, line=None, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for name, row in zip(names, ascii_rows):
        s = row.tobytes()
        out.append(b">" + name.encode() + nl)
        if line:
            out += [s[i:i + line] + nl for i in range(0, len(s), line)]
        else:
            out.append(s + nl)
    return b"".join(out)


def make(n, width, seed, lower=False):
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"ACGTACGTACGTACGTRYMWSKVHDBN-?" + (b"acgtn" if lower else b""), dtype=np.uint8)
    return letters[rng.integers(0, letters.size, size=(n, width))]


def compare_tsv(measure, got_text, want_text, streamed=False):
    # Byte-identical text: counts are integers, raw is ONE IEEE division (measures.rs:68), and for loaded files the CLI takes
    # the integer counts from the GPU and evaluates jc69 / k80 / tn93 with the host's libm (host/measures.hpp), which is
    # what Rust's f64::ln calls.  Only STREAMED jc69 / k80 / tn93 come from the device's f64 epilogue, where CUDA's log may
    # differ from glibc's by an ulp: those are compared value-wise.
    if measure in ("n", "n_high", "raw") or not streamed:
        assert got_text == want_text
        return
    g, w = got_text.splitlines(), want_text.splitlines()
    assert len(g) == len(w) and g[0] == w[0]
    for a, b in zip(g[1:], w[1:]):
        a1, a2, av = a.split("\t")
        b1, b2, bv = b.split("\t")
        assert (a1, a2) == (b1, b2)
        if av != bv:  # last printed digit may differ by the 1-ulp difference of log()
            assert av not in ("NaN", "inf", "-inf") and bv not in ("NaN", "inf", "-inf"), (a, b)
            assert abs(float(av) - float(bv)) <= 1.01e-12, (a, b)


@pytest.mark.parametrize("measure", ALL)
def test_all_vs_all_tsv(oracle, tmp_path, measure):
    from distance_b200 import synth
    asc = make(37, 211, 1)
    names = [f"seq{i}" for i in range(37)]
    f = tmp_path / "a.fasta"
    f.write_bytes(fasta_bytes(names, asc))
    codes = synth.encode_ascii(asc)
    a = oracle.Alignment(codes)
    oracle.prepare(measure, [a])
    want, _ = oracle.run(measure, "square", a)
    want_text = oracle.tsv(names, names, "square", want, measure in ("n", "n_high"))
    rc, out, err = run(["-m", measure, str(f)])
    assert rc == 0, err
    compare_tsv(measure, out, want_text)
    # -t / -b never change the output (lib.rs:947-999); -i and stdin are equivalent entry points
    assert run(["-m", measure, "-t", "3", "-b", "1000", "-i", str(f)])[1] == out
    assert run(["-m", measure, "-t", "1"], stdin=f.read_bytes())[1] == out
    o = tmp_path / "out.tsv"
    assert run(["-m", measure, str(f), "-o", str(o)])[0] == 0 and o.read_text() == out


@pytest.mark.parametrize("measure", ["n_high", "raw", "tn93"])
def test_two_files_and_reversed(oracle, tmp_path, measure):
    from distance_b200 import synth
    a_asc, b_asc = make(11, 150, 2, lower=True), make(23, 150, 3, lower=True)
    na, nb = [f"a{i}" for i in range(11)], [f"b{i}" for i in range(23)]
    fa, fb = tmp_path / "a.fa", tmp_path / "b.fa"
    fa.write_bytes(fasta_bytes(na, a_asc, line=60))            # multi-line records
    fb.write_bytes(fasta_bytes(nb, b_asc, line=70, crlf=True))  # CRLF line ends
    A, B = oracle.Alignment(synth.encode_ascii(a_asc)), oracle.Alignment(synth.encode_ascii(b_asc))
    oracle.prepare(measure, [A, B])
    for (x, nx, fx), (y, ny, fy) in (((A, na, fa), (B, nb, fb)), ((B, nb, fb), (A, na, fa))):
        want, _ = oracle.run(measure, "rect", x, y)
        rc, out, err = run(["-m", measure, str(fx), str(fy)])
        assert rc == 0, err
        compare_tsv(measure, out, oracle.tsv(nx, ny, "rect", want, measure == "n_high"))


@pytest.mark.parametrize("measure", ["n", "k80", "tn93"])
def test_stream_mode_tsv(oracle, tmp_path, measure):
    loaded_asc, streamed_asc = make(9, 120, 4, lower=True), make(300, 120, 5, lower=True)
    nl, ns = [f"L{i}" for i in range(9)], [f"S{i}" for i in range(300)]
    fl, fs = tmp_path / "l.fa", tmp_path / "s.fa"
    fl.write_bytes(fasta_bytes(nl, loaded_asc))
    fs.write_bytes(fasta_bytes(ns, streamed_asc))
    L = oracle.Alignment(np.stack([oracle.encode(r.tobytes()) for r in loaded_asc]))
    enc = [oracle.encode_count_bases(r.tobytes()) for r in streamed_asc]  # fastaio.rs:250-254
    S = oracle.Alignment(np.stack([c for c, _ in enc]), np.stack([k for _, k in enc]))
    oracle.prepare(measure, [L, S], consensus_from=[L])
    want, _ = oracle.run(measure, "stream", L, S)
    want_text = oracle.tsv(nl, ns, "stream", want, measure == "n")
    rc, out, err = run(["-m", measure, "-i", str(fl), "-s", str(fs)])
    assert rc == 0, err
    compare_tsv(measure, out, want_text, streamed=True)
    rc, out2, err = run(["-m", measure, str(fl), "-s", "-"], stdin=fs.read_bytes())  # lib.rs:201-203
    assert rc == 0 and out2 == out


def test_reference_integration_fixtures(tmp_path):
    # lib.rs:906-914 FASTA_1 / FASTA_2 and the expected texts of test_integration_1/2/3
    f1, f2 = tmp_path / "f1.fa", tmp_path / "f2.fa"
    f1.write_text(">seq1\nATGATG\n>seq2\nATGATC\n")
    f2.write_text(">seqA\nATGATG\n")
    hdr = "sequence1\tsequence2\tdistance\n"
    assert run(["-m", "n", str(f1)])[1] == hdr + "seq1\tseq2\t1\n"                                   # lib.rs:938-940
    assert run(["-m", "n_high", "-i", str(f1), "-s", str(f2)])[1] == hdr + "seq1\tseqA\t0\nseq2\tseqA\t1\n"  # :1016-1019
    assert run(["-m", "n_high", str(f1), str(f2)])[1] == hdr + "seq1\tseqA\t0\nseq2\tseqA\t1\n"      # :1082-1085
    assert run(["-m", "n_high", str(f2), str(f1)])[1] == hdr + "seqA\tseq1\t0\nseqA\tseq2\t1\n"      # :1145-1148
    assert run([str(f1)])[1] == hdr + "seq1\tseq2\t0.166666666667\n"                                 # default -m raw


def test_special_float_text(tmp_path):
    f = tmp_path / "s.fa"
    f.write_text(">a\nACGTACGT\n>b\nACGTACGT\n>n\nNNNNNNNN\n>c\nAAAANNNN\n>d\nCCCANNNN\n")
    out = run(["-m", "jc69", str(f)])[1].splitlines()
    assert out[1] == "a\tb\t-0.000000000000"   # -0.75 * ln(1) = -0.0 keeps its sign in `{:.12}`
    assert out[2] == "a\tn\tNaN"
    assert "c\td\tinf" in out
    assert run(["-m", "tn93", str(f)])[1].splitlines()[1] == "a\tb\t0.000000000000"  # measures.rs:188-190


def test_stream_length_error_precedes_invalid_char(tmp_path):
    fl, fs = tmp_path / "l.fa", tmp_path / "s.fa"
    fl.write_text(">l\nACGT\n")
    fs.write_text(">s1\nACGT\n>s2\nAXG\n")
    rc, out, err = run(["-i", str(fl), "-s", str(fs)])
    assert rc == 1 and err == 'Error: Message("Different length sequences in alignment(s): 3 vs 4")\n'  # fastaio.rs:246-248
    fs.write_text(">s1\nACGT\n>s2\nAXGT\n")
    assert run(["-i", str(fl), "-s", str(fs)])[2] == "Error: Message(\"Invalid nucleotide character in record 's2': 'X'\")\n"


@pytest.mark.parametrize("measure", ALL)
def test_cli_matches_committed_golden_tsv(measure):
    """The committed fixtures of tests/golden (oracle-generated, tests/golden/make_golden.py): every code of
    encoding.rs, -0.0 / NaN / inf pairs, in all three modes."""
    g = os.path.join(ROOT, "tests", "golden")
    a, b = os.path.join(g, "golden_a.fasta"), os.path.join(g, "golden_b.fasta")
    for mode, args in (("square", [a]), ("rect", [a, b]), ("stream", ["-i", a, "-s", b])):
        rc, out, err = run(["-m", measure] + args)
        assert rc == 0, err
        compare_tsv(measure, out, open(os.path.join(g, f"golden_{mode}_{measure}.tsv")).read(), streamed=mode == "stream")


def test_broken_pipe_exits_zero(tmp_path):
    """`distance a.fasta | head -1`: the writer meets EPIPE and the process leaves with status 0, like the reference
    (lib.rs:598-608: ErrorKind::BrokenPipe -> std::process::exit(0))."""
    f = tmp_path / "a.fasta"
    names = [f"s{i}" for i in range(600)]
    f.write_bytes(fasta_bytes(names, make(600, 400, 8)))          # 179,700 lines: far more than a pipe buffer holds
    p = subprocess.Popen([CLI, "-m", "n_high", str(f)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    first = p.stdout.readline()
    p.stdout.close()                                              # the reader goes away with most of the output unwritten
    rc = p.wait(timeout=120)
    err = p.stderr.read().decode()
    p.stderr.close()
    assert first == b"sequence1\tsequence2\tdistance\n"
    assert rc == 0, err
    assert err == ""


def test_default_device_list_covers_every_gpu(tmp_path):
    """No DISTANCE_GPUS: a streamed run (unknown length) takes every visible device, like the reference takes every core
    (lib.rs:252-264); results do not depend on the device count."""
    import distance_b200 as dg
    fl, fs = tmp_path / "l.fa", tmp_path / "s.fa"
    fl.write_bytes(fasta_bytes([f"L{i}" for i in range(5)], make(5, 300, 11)))
    fs.write_bytes(fasta_bytes([f"S{i}" for i in range(700)], make(700, 300, 12)))
    env = dict(os.environ, DG_TRACE="1")
    env.pop("DISTANCE_GPUS", None)
    p = subprocess.run([CLI, "-m", "k80", "-i", str(fl), "-s", str(fs)], capture_output=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr.decode()
    assert f"{dg.device_count()} GPU(s)" in p.stderr.decode()
    one = subprocess.run([CLI, "-m", "k80", "-i", str(fl), "-s", str(fs)], capture_output=True, timeout=600,
                         env=dict(env, DISTANCE_GPUS="1"))
    assert one.returncode == 0 and one.stdout == p.stdout


@pytest.mark.parametrize("measure", ["jc69", "k80", "tn93"])
def test_float_text_is_byte_identical_at_scale(oracle, tmp_path, measure):
    """500,000 pairs of SARS-CoV-2-like records: with the host-side libm evaluation every line equals the oracle's text
    (glibc log, like Rust's f64::ln); the device epilogue (DISTANCE_DEVICE_F64=1) stays within 1e-12 of it."""
    from distance_b200 import synth
    asc = synth.make_alignment(1001, width=4000, seed=31, ambiguity=True, mu=4e-3)
    names = synth.ids(1001)
    f = tmp_path / "a.fasta"
    synth.write_fasta(str(f), asc, names)
    a = oracle.Alignment(synth.encode_ascii(asc))
    oracle.prepare(measure, [a])
    want, _ = oracle.run(measure, "square", a, threads=8)
    want_text = oracle.tsv(names, names, "square", want, False)
    rc, out, err = run(["-m", measure, str(f)])
    assert rc == 0, err
    assert out == want_text
    p = subprocess.run([CLI, "-m", measure, str(f)], capture_output=True, timeout=600, env=dict(os.environ, DISTANCE_DEVICE_F64="1"))
    assert p.returncode == 0
    compare_tsv(measure, p.stdout.decode(), want_text, streamed=True)
