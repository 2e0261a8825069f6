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


# ---------------------------------------------------------------------------------------------
# the parallel FASTA loader against the sequential reader (same ids / bytes / error text)
# ---------------------------------------------------------------------------------------------
def _fasta(records, line=0, eol="\n", final_eol=True):
    out = []
    for rid, seq in records:
        out.append(">" + rid)
        if line:
            out.extend(seq[i:i + line] for i in range(0, len(seq), line))
        else:
            out.append(seq)
    return eol.join(out) + (eol if final_eol else "")


def _selftest_parse(tmp_path, text, threads=4, binary=False):
    f = tmp_path / "x.fa"
    f.write_bytes(text if binary else text.encode())
    rc, out, err = run(["--selftest-parse", str(f), str(threads)])
    assert rc == 0 and "identical" in out, (out, err)
    return out


def test_parallel_loader_matches_sequential_reader(tmp_path):
    import random
    rnd = random.Random(7)
    width = 5000
    recs = [(f"seq{i} some description", "".join(rnd.choice("ACGTNacgtn-?RYKM") for _ in range(width))) for i in range(3000)]
    big = _fasta(recs)                                   # 15 MB: several segments
    assert "parallel pass" in _selftest_parse(tmp_path, big)
    assert "parallel pass" in _selftest_parse(tmp_path, _fasta(recs, line=70))                 # multi-line
    assert "parallel pass" in _selftest_parse(tmp_path, _fasta(recs, line=60, eol="\r\n"))     # CRLF
    assert "parallel pass" in _selftest_parse(tmp_path, _fasta(recs, final_eol=False), threads=16)
    assert "parallel pass" in _selftest_parse(tmp_path, _fasta(recs[:1]))


def test_parallel_loader_falls_back_for_errors_and_odd_input(tmp_path):
    import random
    rnd = random.Random(9)
    width = 4000
    recs = [(f"r{i}", "".join(rnd.choice("ACGT") for _ in range(width))) for i in range(3000)]
    bad = list(recs)
    bad[2500] = ("r2500", recs[2500][1][:100] + "X" + recs[2500][1][101:])      # invalid nucleotide late in the file
    out = _selftest_parse(tmp_path, _fasta(bad))
    assert "fell back" in out and "Invalid nucleotide character in record 'r2500': 'X'" in out
    short = list(recs)
    short[1700] = ("r1700", recs[1700][1][:-1])                                   # width mismatch
    out = _selftest_parse(tmp_path, _fasta(short))
    assert "fell back" in out and "Different length sequences" in out
    both = list(short)
    both[900] = ("r900", "U" + recs[900][1][1:])                                  # the earlier offence wins
    out = _selftest_parse(tmp_path, _fasta(both))
    assert "Invalid nucleotide character in record 'r900': 'U'" in out
    assert "fell back" in _selftest_parse(tmp_path, "junk\n" + _fasta(recs))        # no '>' at the start
    assert "Empty FASTA file" in _selftest_parse(tmp_path, "")
    _selftest_parse(tmp_path, _fasta(recs[:10]) + ">\n\n" + _fasta(recs[10:20]))    # an empty record ends the iteration
    _selftest_parse(tmp_path, _fasta(recs[:50]) + "\n\n")                           # trailing blank lines
    _selftest_parse(tmp_path, _fasta(recs[:5]).encode() + b"\xff\xfe", binary=True)  # non-ASCII tail


def test_pooled_tsv_writer_matches_line_by_line_text():
    rc, out, err = run(["--selftest-tsv"])
    assert rc == 0 and "0 mismatches" in out, (out, err)


def _selftest_stream(tmp_path, text, width, batch, threads=4):
    f = tmp_path / "s.fa"
    f.write_bytes(text.encode())
    rc, out, err = run(["--selftest-stream", str(f), str(width), str(batch), str(threads)])
    assert rc == 0 and "identical" in out, (out, err)
    return out


def test_stream_block_parser_matches_sequential_reader(tmp_path):
    import random
    rnd = random.Random(11)
    width = 3000
    recs = [(f"q{i} desc", "".join(rnd.choice("ACGTNacgt-?RY") for _ in range(width))) for i in range(2500)]
    for batch in (1, 7, 64, 1000, 5000):
        assert "records 2500 (block parser 2500)" in _selftest_stream(tmp_path, _fasta(recs), width, batch)
    _selftest_stream(tmp_path, _fasta(recs, line=70, eol="\r\n"), width, 333)
    _selftest_stream(tmp_path, _fasta(recs, final_eol=False), width, 4096, threads=16)
    # errors: the first offence in file order, width before nucleotides (fastaio.rs:246-254)
    bad = list(recs)
    bad[2000] = ("q2000", recs[2000][1][:50] + "X" + recs[2000][1][51:])
    assert "Invalid nucleotide character in record 'q2000': 'X'" in _selftest_stream(tmp_path, _fasta(bad), width, 256)
    bad[1200] = ("q1200", recs[1200][1][:-3])
    assert "Different length sequences in alignment(s): 2997 vs 3000" in _selftest_stream(tmp_path, _fasta(bad), width, 256)
    assert "Different length" in _selftest_stream(tmp_path, _fasta(recs), width + 1, 256)      # every record is "wrong"
    assert "Expected > at record start." in _selftest_stream(tmp_path, "garbage\n" + _fasta(recs[:10]), width, 4)
    # an empty record ends the iteration; an empty id with a description does not
    out = _selftest_stream(tmp_path, _fasta(recs[:300]) + ">\n\n" + _fasta(recs[300:400]), width, 64)
    assert "records 300 (block parser 300)" in out
    out = _selftest_stream(tmp_path, _fasta(recs[:300]) + "> only a description\n" + recs[0][1] + "\n" + _fasta(recs[300:400]), width, 64)
    assert "records 401 (block parser 401)" in out
    assert "records 0" in _selftest_stream(tmp_path, "", width, 8)


def test_parsers_fuzz_against_sequential_reader(tmp_path):
    """Random FASTA-like inputs (ragged line lengths, CR LF mixes, blank lines, stray whitespace, empty ids, width
    mismatches, invalid bytes, junk) cut into many tiny segments / batches: the parallel loader and the stream block parser
    must agree with the sequential reader on ids, bytes and error text every time."""
    import random
    rnd = random.Random(20251018)
    env = dict(os.environ, DG_PARSE_SEG="97")
    alphabet = "ACGTNacgtn-?RYKMSWBDHV"
    for case in range(int(os.environ.get("DG_FUZZ_CASES", "60"))):
        width = rnd.choice([1, 7, 60, 61, 250])
        n = rnd.randint(1, 40)
        eol = rnd.choice(["\n", "\r\n"])
        parts = []
        for i in range(n):
            seq = "".join(rnd.choice(alphabet) for _ in range(width))
            kind = rnd.random()
            if case % 3 == 1 and kind < 0.04:
                seq = seq[:-1]                                    # width mismatch
            elif case % 3 == 2 and kind < 0.04:
                k = rnd.randrange(width)
                seq = seq[:k] + rnd.choice("XUZ*.1 ") + seq[k + 1:]   # invalid byte (or an inner blank)
            rid = f"id{i}" if rnd.random() > 0.03 else ""
            desc = rnd.choice(["", " d", "\tx y ", "  "])
            parts.append(">" + rid + desc + eol)
            line = rnd.choice([0, 0, 13, 60, 70])
            chunks = [seq] if not line else [seq[j:j + line] for j in range(0, len(seq), line)]
            for ch in chunks:
                parts.append(ch + rnd.choice(["", "", " ", "\t"]) + eol)
                if rnd.random() < 0.05:
                    parts.append(eol)                             # blank line inside a record
        text = "".join(parts)
        if rnd.random() < 0.3:
            text = text.rstrip("\r\n")                            # no final newline
        if case % 10 == 9:
            text = rnd.choice(["junk" + eol, eol, " "]) + text    # something before the first '>'
        f = tmp_path / f"fuzz{case}.fa"
        f.write_bytes(text.encode())
        p = subprocess.run([CLI, "--selftest-parse", str(f), str(rnd.choice([1, 3, 8]))], capture_output=True, env=env, timeout=60)
        assert p.returncode == 0 and b"identical" in p.stdout, (case, p.stdout, p.stderr)
        batch = rnd.choice([1, 2, 5, 17])
        p = subprocess.run([CLI, "--selftest-stream", str(f), str(width), str(batch), str(rnd.choice([1, 4]))],
                           capture_output=True, env=env, timeout=60)
        assert p.returncode == 0 and b"identical" in p.stdout, (case, batch, p.stdout, p.stderr)


def test_fasta_tokenisation_choices(tmp_path):
    """The cases tests/golden/README.md lists as parity-unpinned against rust-bio: each documented choice, checked on both
    readers (`--selftest-parse` runs the parallel loader and the sequential reader and reports records / width / error)."""
    recs = [(f"r{i}", "ACGTACGTAC") for i in range(6)]
    out = _selftest_parse(tmp_path, _fasta(recs))
    assert "records 6 width 10" in out
    # '>' alone with sequence lines behind it is a record with the empty id
    out = _selftest_parse(tmp_path, _fasta(recs[:2]) + ">\nACGTACGTAC\n" + _fasta(recs[2:]))
    assert "records 7 width 10" in out
    # '>' alone with nothing behind it ends the iteration: the records after it are never read
    out = _selftest_parse(tmp_path, _fasta(recs[:2]) + ">\n\n" + _fasta(recs[2:]))
    assert "records 2 width 10" in out
    # header: the id ends at the first blank, the description is dropped; CRLF and trailing blanks are trimmed
    out = _selftest_parse(tmp_path, ">a some text\r\nACGTA \t\r\nCGTAC\r\n>b\r\nACGTACGTAC\r\n")
    assert "records 2 width 10" in out
    # a blank line inside a record adds nothing
    out = _selftest_parse(tmp_path, ">a\nACGTA\n\nCGTAC\n>b\nACGTACGTAC\n\n\n")
    assert "records 2 width 10" in out
    # blank line before the first '>'
    out = _selftest_parse(tmp_path, "\n" + _fasta(recs))
    assert "Expected > at record start." in out
    # an inner blank is a sequence byte: invalid nucleotide
    out = _selftest_parse(tmp_path, ">a\nACGTA CGTA\n")
    assert "Invalid nucleotide character in record 'a': ' '" in out
    # str::trim_end strips Unicode White_Space: U+00A0 at a line end is trimmed like a blank
    out = _selftest_parse(tmp_path, ">a\nACGTACGTAC\u00a0\n")
    assert "records 1 width 10" in out and "error ''" in out


def _dump(tmp_path, data, threads=4):
    f = tmp_path / "u.fa"
    f.write_bytes(data)
    p = subprocess.run([CLI, "--selftest-parse", str(f), str(threads)], capture_output=True, timeout=60,
                       env=dict(os.environ, DG_SELFTEST_DUMP="1"))
    assert p.returncode == 0, (p.stdout, p.stderr)
    lines = p.stdout.decode("utf-8", "replace").split("\n")
    assert "identical" in lines[0], lines[0]
    return lines[0], [tuple(x.split("\t")) for x in lines[1:] if x]


def test_unicode_white_space_and_utf8_like_rust_bio(tmp_path):
    """rust-bio 1.6.0 reads lines into a String (invalid UTF-8 = io::ErrorKind::InvalidData), trims them with str::trim_end and
    splits the header at the first char::is_whitespace: all three use Unicode White_Space, not the ASCII blanks."""
    ws = ["\u0085", "\u00a0", "\u1680", "\u2000", "\u2003", "\u200a", "\u2028", "\u2029", "\u202f", "\u205f", "\u3000", "\x0b", "\x0c"]
    # every White_Space character is trimmed from the end of a sequence line and of a header, and ends an id
    for k, w in enumerate(ws):
        text = f">id{k}{w}a description{w}{w}\nACGTA{w}\nCGTAC{w} \t{w}\n>next\nACGTACGTAC\n"
        head, recs = _dump(tmp_path, text.encode("utf-8"))
        assert "records 2 width 10" in head and "error ''" in head, (w.encode("unicode_escape"), head)
        assert recs == [(f"id{k}", "ACGTACGTAC"), ("next", "ACGTACGTAC")], (w.encode("unicode_escape"), recs)
    # not white space: U+200B (zero width space), U+180E, U+FEFF stay -- in an id as bytes of the id, in a sequence as an invalid byte
    for w in ["\u200b", "\u180e", "\ufeff"]:
        head, recs = _dump(tmp_path, f">a{w}b c\nACGT\n".encode("utf-8"))
        assert recs == [(f"a{w}b", "ACGT")], recs
        head, _ = _dump(tmp_path, f">a\nACGT{w}\n".encode("utf-8"))
        assert "Invalid nucleotide character in record 'a': '" in head
    # a non-ASCII letter in a sequence: the first byte of its UTF-8 form, shown as the Latin-1 character (`*nuc as char`)
    head, _ = _dump(tmp_path, ">a\nAC\u00e9GT\n".encode("utf-8"))
    assert "Invalid nucleotide character in record 'a': '\u00c3'" in head
    # invalid UTF-8 anywhere in a line is BufRead::read_line's error, whatever else is wrong with the record ...
    bad_utf8 = 'IOError(Error { kind: InvalidData, message: "stream did not contain valid UTF-8" })'
    for data in [b">a\nACGT\xff\n", b">a \xc3\x28 desc\nACGT\n", b">a\nACXT\nAC\xe2\x80\n", b">a\nACGT\n>b\xed\xa0\x80\nACGT\n",
                 b">a\nACGT\n>b\nAC\xc0\xafGT\n", b"\xf5>a\nACGT\n", b">a\nACGT\xf4\x90\x80\x80\n"]:
        head, _ = _dump(tmp_path, data)
        assert bad_utf8 in head, (data, head)
    # ... read while its record is read: the header of the NEXT record is read with this record's lines, before encode() runs
    head, _ = _dump(tmp_path, b">a\nACXT\n>b\xff\nACGT\n")
    assert bad_utf8 in head
    # but an invalid nucleotide in an EARLIER record is met first
    head, _ = _dump(tmp_path, b">a\nACXT\n>b\nACGT\n>c\xff\nACGT\n")
    assert "Invalid nucleotide character in record 'a': 'X'" in head
    # the block parser of the streamed file hands such records to the same sequential reader
    recs = [(f"q{i}", "ACGTACGTACGT") for i in range(400)]
    text = _fasta(recs[:100]) + ">odd\u2003id\u00a0\nACGTAC\u3000\nGTACGT\n" + _fasta(recs[100:])
    assert "records 401 (block parser 401)" in _selftest_stream(tmp_path, text, 12, 64)
    f = tmp_path / "s.fa"
    f.write_bytes(_fasta(recs[:100]).encode() + b">x\nACGTAC\xffTACGT\n" + _fasta(recs[100:]).encode())
    rc, out, err = run(["--selftest-stream", str(f), "12", "64", "4"])
    assert rc == 0 and "stream did not contain valid UTF-8" in out, (out, err)
    # through the whole CLI: the reference's `Error: {:?}` line and exit code 1
    f = tmp_path / "bad.fa"
    f.write_bytes(b">a\nACGT\n>b\nAC\xffT\n")
    rc, out, err = run(["-i", str(f)])
    assert rc == 1 and out == "" and err.strip() == "Error: " + bad_utf8
