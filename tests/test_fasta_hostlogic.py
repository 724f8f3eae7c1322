"""FASTA ingestion, host-side checks without a GPU.

1. The reference's own parser (read_fasta_files, src/read_fasta_files.c:412, compiled into
   oracle/_ref over an in-memory file) against the Python restatement oracle/fasta.py on
   adversarial texts: CRLF and lone CR, comments, empty lines and empty records, invalid
   letters, no trailing newline, lower case, lines longer than a chunk.
2. The chunk walk + resolve step the two FASTA kernels execute (bsgpu_fasta.cuh), compiled for
   the CPU by tests/hostcheck/fasta_hostcheck.cpp and run chunk after chunk, against both --
   every chunk boundary position is exercised by sliding the same records through the text.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import fasta as OF
from oracle import ref
from tests.helpers import rng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")

TEXTS = [
    b">sequence1\nAGTACGTAGTCGCTGCTGCTACGGGCGCTAGCTAGTACGTCA\nCGACGTAGATGCTAGCTGACTAAACGATGC\n"
    b">sequence2\nAAACGATCGATCGTACTCGACTGATGTAGTATATACGTCGTACGTAG\nCATCGTCAGTTACTGCATGCGGG\n\n",
    b";comment\n\n>r1 desc\r\nACGTN\r\nacgXt\n\n>r2\n>r3\nAC\rGT",
    b">only header",
    b">a\n>b\n>c\n",
    b">a\nACGT",
    b">a\r\nAC GT\r\n\r\n;in the middle\r\nNNNN\r\n>b\r\n\r\n",
    b"\n\n;x\n>a\nA\n",
    b">a\n-+.acgtmrwsykvhdbn\nACGTMRWSYKVHDBN\n12345\n",
    b"",
    b"\n",
    b">a\n\r",
    b">x\n" + b"ACGT" * 300 + b"\n>y\n" + b"TTGCA" * 250,
]


def rand_fasta(g, nrec, crlf=False, max_len=3000, line=60):
    eol = b"\r\n" if crlf else b"\n"
    parts = []
    letters = np.frombuffer(b"ACGTacgtNRY-X*", dtype=np.uint8)
    for r in range(nrec):
        parts.append(b">rec%d some description %d" % (r, int(g.integers(0, 10 ** 6))) + eol)
        n = int(g.integers(0, max_len))
        s = letters[g.integers(0, len(letters), n)].tobytes()
        w = int(g.integers(1, line + 1))
        for i in range(0, n, w):
            parts.append(s[i:i + w] + eol)
        if g.random() < 0.3:
            parts.append(eol)
        if g.random() < 0.2:
            parts.append(b";a comment line" + eol)
    return b"".join(parts)


def same(a, b):
    (wa, ca, na, ia), (wb, cb, nb, ib) = a, b
    assert wa.tolist() == wb.tolist()
    assert np.array_equal(ca, cb)
    assert na == nb
    assert ia == ib


def ref_parse(text):
    w, c, names, warn = ref.read_fasta(text)
    ninvalid = 0
    if warn:
        ninvalid = int(warn.split("ignored ")[1].split(" ")[0])
    return w, c, names, ninvalid


@needs_ref
@pytest.mark.parametrize("i", range(len(TEXTS)))
def test_restatement_equals_reference_on_fixed_texts(i):
    same(OF.read_fasta(TEXTS[i]), ref_parse(TEXTS[i]))


@needs_ref
def test_restatement_equals_reference_on_random_files():
    g = rng(2024)
    for k in range(12):
        text = rand_fasta(g, int(g.integers(1, 12)), crlf=bool(k % 2))
        same(OF.read_fasta(text), ref_parse(text))


@needs_ref
def test_errors_of_the_reference_and_the_restatement():
    for text, line in [(b"ACGT\n>r\nAC\n", 1), (b"\n;c\n\nAC\n>r\n", 4), (b"X", 1), (b"\r\r\n>a\n", 1)]:
        with pytest.raises(ref.RefError, match=f'">" expected at beginning of line {line}'):
            ref.read_fasta(text)
        with pytest.raises(ValueError, match=f'">" expected at beginning of line {line}'):
            OF.read_fasta(text)


def test_reference_fixture_facts():
    """inst/extdata/fastaEx.fa of the reference (two records, 72 and 70 letters; first text
    of TEXTS) as its man page uses it (man/XStringSet-io.Rd:313)."""
    w, c, names, ninvalid = OF.read_fasta(TEXTS[0])
    assert w.tolist() == [72, 70] and names == ["sequence1", "sequence2"] and ninvalid == 0
    assert c[:4].tolist() == [1, 4, 8, 1]


# ---- the kernels' chunk walk on the CPU --------------------------------------------------------

@pytest.fixture(scope="module")
def hostcheck(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostcheck") / "fasta_hostcheck.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "hostcheck", "fasta_hostcheck.cpp")])
    L = C.CDLL(so)
    L.fasta_hostcheck.argtypes = [C.c_char_p, C.c_int64] + [C.c_void_p] * 5
    return L


def walk(L, text):
    n = len(text)
    lk = OF.dna_lkup16()
    out = np.full(128 + 2 * n + 256, 0x80, dtype=np.uint8)
    seg_start = np.zeros(n + 2, dtype=np.int64)
    desc_off = np.zeros(n + 2, dtype=np.int64)
    totals = np.zeros(5, dtype=np.int64)
    rc = L.fasta_hostcheck(text, n, lk.ctypes.data, out.ctypes.data, seg_start.ctypes.data,
                           desc_off.ctypes.data, totals.ctypes.data)
    kept, invalid, nrec, first_header, first_seqline = totals.tolist()
    if rc == 1:
        raise ValueError("sequence data before the first record")
    starts = seg_start[:nrec]
    ends = np.append(starts[1:] - 1, 128 + kept + max(nrec - 1, 0)) if nrec else starts
    widths = (ends - starts).astype(np.int32) if nrec else np.zeros(0, np.int32)
    concat = np.concatenate([out[s:s + w] for s, w in zip(starts, widths)]) if nrec else np.zeros(0, np.uint8)
    # separators between the records are untouched
    for s, w in zip(starts, widths):
        assert out[s + w] == 0x80
    names = [OF.desc_at(text, int(o)) for o in desc_off[:nrec]]
    return widths, concat, names, invalid


@pytest.mark.parametrize("i", range(len(TEXTS)))
def test_chunk_walk_on_fixed_texts(hostcheck, i):
    same(walk(hostcheck, TEXTS[i]), OF.read_fasta(TEXTS[i]))


def test_chunk_walk_at_every_boundary_position(hostcheck):
    """The same records slid byte by byte across the 512-byte chunk boundaries: description
    lines, CRLF pairs, comments and empty lines all get cut at every possible place."""
    core = b">r1 d\r\nACGTN\r\nacgXt\n\n;c\n>r2\n>r3\nAC\rGT\r\n"
    for pad in range(0, 540, 1):
        text = b">p\n" + b"A" * pad + b"\n" + core + b"C" * (600 - pad) + b"\n" + core
        same(walk(hostcheck, text), OF.read_fasta(text))


def test_chunk_walk_on_random_files(hostcheck):
    g = rng(77)
    for k in range(30):
        text = rand_fasta(g, int(g.integers(1, 40)), crlf=bool(k % 3 == 0), line=int(g.integers(1, 700)))
        same(walk(hostcheck, text), OF.read_fasta(text))


def test_chunk_walk_long_lines(hostcheck):
    """a description and a sequence line much longer than a chunk (classes carried over
    several chunks with no line start: class_out = PASS)."""
    text = b">" + b"d" * 3000 + b"\n" + b"ACGT" * 1000 + b"\n;" + b"c" * 2000 + b"\nTT\n"
    same(walk(hostcheck, text), OF.read_fasta(text))


def test_chunk_walk_rejects_data_before_the_first_record(hostcheck):
    for text in (b"ACGT\n>r\nAC\n", b"\n;c\n\nAC\n>r\n", b"A" * 2000 + b"\n>r\nAC\n"):
        with pytest.raises(ValueError):
            walk(hostcheck, text)


@needs_ref
def test_nrec_skip_seek_first_rec_cut_equals_reference():
    """biostrings_b200.io._cut (host logic of the product: nrec / skip / seek.first.rec applied
    to the text before the upload) against the reference's own handling of these arguments."""
    from biostrings_b200.io import _cut
    g = rng(5150)
    text = b"junk line\n;c\n" + rand_fasta(g, 9)
    clean = text[text.index(b">"):]
    for nrec, skip in [(-1, 0), (3, 0), (-1, 2), (2, 4), (0, 0), (5, 7), (-1, 9), (4, 20)]:
        w, c, names, warn = ref.read_fasta(clean, nrec=nrec, skip=skip)
        got = OF.read_fasta(_cut(clean, nrec, skip, False)[0])
        assert got[0].tolist() == w.tolist() and np.array_equal(got[1], c) and got[2] == names
        w, c, names, warn = ref.read_fasta(text, nrec=nrec, skip=skip, seek_first_rec=True)
        got = OF.read_fasta(_cut(text, nrec, skip, True)[0])
        assert got[0].tolist() == w.tolist() and np.array_equal(got[1], c) and got[2] == names
    with pytest.raises(ref.RefError, match="no FASTA record found"):
        ref.read_fasta(b"junk\nmore junk\n", seek_first_rec=True)
    with pytest.raises(ValueError, match="no FASTA record found"):
        _cut(b"junk\nmore junk\n", -1, 0, True)
    # sequence data in front of the first record is an error whatever `skip` says
    with pytest.raises(ref.RefError, match='">" expected at beginning of line 3'):
        ref.read_fasta(b";c\n\nACGT\n>a\nAC\n>b\nGT\n", skip=1)
    with pytest.raises(Exception, match='">" expected at beginning of line 3'):
        _cut(b";c\n\nACGT\n>a\nAC\n>b\nGT\n", -1, 1, False)


@needs_ref
def test_file_list_shares_one_record_counter():
    """readDNAStringSet(<list of files>, nrec, skip): the per-file (nrec, skip) windows the product
    derives (biostrings_b200.io._file_windows) applied file by file equal read_fasta_files() of
    the reference on the whole list."""
    from biostrings_b200.io import _cut, _file_windows, _record_starts
    g = rng(5160)
    files = [rand_fasta(g, n) for n in (4, 1, 6, 3)]
    files = [f[f.index(b">"):] for f in files]
    counts = [int(_record_starts(np.frombuffer(f, dtype=np.uint8)).size) for f in files]
    for nrec, skip in [(-1, 0), (3, 2), (3, 0), (-1, 5), (6, 4), (2, 11), (0, 3), (50, 1), (1, 4)]:
        w, c, names, warn = ref.read_fasta(files, nrec=nrec, skip=skip)
        gw, gc, gn = [], [], []
        for f, (f_nrec, f_skip) in zip(files, _file_windows(counts, nrec, skip)):
            ww, cc, nn, _ = OF.read_fasta(_cut(f, f_nrec, f_skip, False)[0])
            gw += ww.tolist()
            gc.append(cc)
            gn += nn or []
        assert gw == w.tolist() and gn == (names or []), (nrec, skip)
        assert np.array_equal(np.concatenate(gc) if gc else np.zeros(0, np.uint8), c)


def test_chunk_walk_lines_longer_than_a_group(hostcheck):
    """description, comment and sequence lines longer than a group of 64 chunks (32 KiB): the
    class a group starts in is only known to the host step (both hypotheses are folded on the
    device), including groups in which no line starts at all."""
    g = rng(31337)
    seq = np.frombuffer(b"ACGTNacgtX", dtype=np.uint8)[g.integers(0, 10, 150_000)].tobytes()
    text = (b">" + b"d" * 70_000 + b"\n" + seq[:100_000] + b"\n;" + b"c" * 40_000 + b"\r\n" +
            seq[100_000:] + b"\n>r2\n" + b"A" * 33_000 + b"\r\n>r3 " + b"x" * 66_000)
    same(walk(hostcheck, text), OF.read_fasta(text))
    for cut in (32_768, 32_769, 65_536, 98_304 + 17):
        same(walk(hostcheck, text[:cut]), OF.read_fasta(text[:cut]))
