"""The C++ `distance` host (CLI, FASTA parser, TSV formatter): everything that happens before the
first GPU call is checked here without a GPU -- flags, help text, the reference's error texts and exit
codes, the exact `{:.12}` formatter."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "distance_b200", "_bin", "distance")


@pytest.fixture(scope="module", autouse=True)
def build_cli():
    subprocess.check_call(["make", "-C", ROOT, "cli"], stdout=subprocess.DEVNULL)


def run(args, stdin=None):
    p = subprocess.run([CLI] + args, input=stdin, capture_output=True, timeout=120)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def test_help_matches_reference_readme():
    # README.md:113-139 of the reference is the CLI contract
    want = open(os.path.join(ROOT, "tests", "golden", "help.txt")).read()
    rc, out, _ = run(["-h"])
    assert rc == 0 and out == want
    assert run(["--help"])[1] == want


def test_version():
    assert run(["-V"]) == (0, "distance 0.3.1\n", "")  # Cargo.toml:3 via crate_version!()


def test_exact_float12_formatter():
    rc, out, err = run(["--selftest-format", "400000"])
    assert rc == 0, err
    assert "0 mismatches" in out


def test_invalid_measure_is_a_clap_error():
    rc, _, err = run(["-m", "hamming"])
    assert rc == 2 and "invalid value 'hamming'" in err and "possible values: n, n_high, raw, jc69, k80, tn93" in err


def test_missing_file(tmp_path):
    rc, _, err = run([str(tmp_path / "nope.fasta")])
    assert rc == 1
    assert err == 'Error: IOError(Os { code: 2, kind: NotFound, message: "No such file or directory" })\n'


def test_positional_and_flag_inputs_conflict(tmp_path):  # lib.rs:182-184
    f = tmp_path / "a.fa"
    f.write_text(">a\nACGT\n")
    rc, _, err = run([str(f), "-i", str(f)])
    assert rc == 1
    assert err == ('Error: Message("For loading input files, don\'t use both positional arguments and the '
                   '-i/--input flag")\n')


def test_stream_needs_exactly_one_loaded_file(tmp_path):  # lib.rs:196-199
    f = tmp_path / "a.fa"
    f.write_text(">a\nACGT\n")
    for args in (["-s", str(f)], ["-i", str(f), str(f), "-s", str(f)]):
        rc, _, err = run(args, stdin=b"")
        assert rc == 1
        assert err == ('Error: Message("If you stream one file, you must also provide exactly one other file '
                       'to be loaded")\n')


def test_invalid_nucleotide_message(tmp_path):  # fastaio.rs:89-91, 111-113
    f = tmp_path / "a.fa"
    f.write_text(">seq1 some description\nACGT\n>seq2\nACUT\n")
    rc, _, err = run([str(f)])
    assert rc == 1
    assert err == "Error: Message(\"Invalid nucleotide character in record 'seq2': 'U'\")\n"


def test_invalid_char_wins_over_length_error_when_loading(tmp_path):
    # load_fasta encodes (fastaio.rs:183) before it compares widths (fastaio.rs:188-190)
    f = tmp_path / "a.fa"
    f.write_text(">a\nACGT\n>b\nACX\n")
    assert "Invalid nucleotide character in record 'b': 'X'" in run([str(f)])[2]


def test_different_lengths(tmp_path):  # fastaio.rs:93-95, 188-190
    f = tmp_path / "a.fa"
    f.write_text(">a\nACGT\n>b\nACG\n")
    rc, _, err = run([str(f)])
    assert rc == 1 and err == 'Error: Message("Different length sequences in alignment(s): 3 vs 4")\n'


def test_different_lengths_across_files(tmp_path):  # fastaio.rs:206-208
    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    a.write_text(">a\nACGT\n")
    b.write_text(">b\nACGTA\n")
    rc, _, err = run([str(a), str(b)])
    assert rc == 1 and err == 'Error: Message("Different length sequences in alignment(s): 4 vs 5")\n'


def test_empty_fasta(tmp_path):  # fastaio.rs:97-99, 195-197
    f = tmp_path / "a.fa"
    f.write_text("")
    rc, _, err = run([str(f)])
    assert rc == 1 and err == 'Error: Message("Empty FASTA file")\n'
    assert run([], stdin=b"")[2] == 'Error: Message("Empty FASTA file")\n'  # no input file -> stdin (lib.rs:189-191)


def test_not_fasta(tmp_path):  # rust-bio: "Expected > at record start."
    f = tmp_path / "a.fa"
    f.write_text("ACGT\n")
    rc, _, err = run([str(f)])
    assert rc == 1 and err == 'Error: IOError(Custom { kind: Other, error: "Expected > at record start." })\n'


def test_licences_flag():
    rc, out, _ = run(["-l"])
    assert rc == 0 and "GNU LIBRARY GENERAL PUBLIC LICENSE" in out
