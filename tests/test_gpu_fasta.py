"""FASTA ingestion on the GPU (fasta_summary_kernel / fasta_emit_kernel
through bsgpu_subject_fasta) against the reference's own parser (read_fasta_files,
src/read_fasta_files.c:412, compiled into oracle/_ref) and its Python restatement: widths,
letters, names, the invalid-letter warning, the error for data before the first record,
nrec / skip / seek.first.rec; then the resident set used as a subject (vcountPDict,
oligonucleotideFrequency) and one record turned into a resident single-sequence subject
(countPDict / matchPDict), equal to the same calls on host sequences."""
import gzip
import warnings

import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import _lib
from oracle import fasta as OF
from oracle import ref
from tests.helpers import random_dna, rng
from tests.test_fasta_hostlogic import TEXTS, rand_fasta

pytestmark = pytest.mark.gpu


def gpu_parse(text, **kw):
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        dset, widths, names = bs.read_fasta_to_device(text, **kw)
    with dset:
        sset = bs.download_set(dset, names)
    ninvalid = 0
    for w in wlist:
        ninvalid = int(str(w.message).split("ignored ")[1].split(" ")[0])
    assert sset.widths.tolist() == widths.tolist()
    return sset.widths, sset.concat, names, ninvalid


def same(a, b):
    assert a[0].tolist() == b[0].tolist()
    assert np.array_equal(a[1], b[1])
    assert a[2] == b[2] and a[3] == b[3]


@pytest.mark.parametrize("i", range(len(TEXTS)))
def test_fixed_texts(i):
    same(gpu_parse(TEXTS[i]), OF.read_fasta(TEXTS[i]))
    if ref.available():
        w, c, names, warn = ref.read_fasta(TEXTS[i])
        got = gpu_parse(TEXTS[i])
        assert got[0].tolist() == w.tolist() and np.array_equal(got[1], c) and got[2] == names
        assert (got[3] != 0) == bool(warn)


def test_every_chunk_boundary_position():
    core = b">r1 d\r\nACGTN\r\nacgXt\n\n;c\n>r2\n>r3\nAC\rGT\r\n"
    for pad in range(0, 540, 7):
        text = b">p\n" + b"A" * pad + b"\n" + core + b"C" * (600 - pad) + b"\n" + core
        same(gpu_parse(text), OF.read_fasta(text))


def test_random_files_and_a_large_one():
    g = rng(99)
    for k in range(8):
        text = rand_fasta(g, int(g.integers(1, 60)), crlf=bool(k % 3 == 0), line=int(g.integers(1, 700)))
        same(gpu_parse(text), OF.read_fasta(text))
    # 24 "chromosomes", 12 Mbp, 60 letters per line
    dec = np.frombuffer(b"ACGTN", dtype=np.uint8)
    parts, seqs = [], []
    for r in range(24):
        n = int(g.integers(100_000, 900_000))
        s = dec[g.choice(5, size=n, p=[0.29, 0.2, 0.2, 0.29, 0.02])]
        seqs.append(s)
        lines = np.full((n + 59) // 60 * 61, ord("\n"), dtype=np.uint8)
        idx = np.arange(n)
        lines[idx + idx // 60] = s
        parts.append(b">chr%d\n" % (r + 1) + lines[:n + (n + 59) // 60].tobytes())
    text = b"".join(parts)
    w, c, names, ninvalid = gpu_parse(text)
    assert names == [f"chr{r + 1}" for r in range(24)] and ninvalid == 0
    assert w.tolist() == [s.size for s in seqs]
    assert np.array_equal(c, bs.encode_dna(np.concatenate(seqs).tobytes()))
    assert _lib.lib().bsgpu_last_plan().decode().startswith("fasta_ingest bytes=")


def test_errors_and_arguments(tmp_path):
    with pytest.raises(bs.BsgpuError, match='reading FASTA file <memory>: ">" expected at beginning of line 4'):
        bs.read_fasta_to_device(b"\n;c\n\nAC\n>r\n")
    with pytest.raises(bs.BsgpuError, match='">" expected at beginning of line 1'):
        bs.read_fasta_to_device(b"A" * 2000 + b"\n>r\nAC\n")
    with pytest.raises(ValueError, match="no FASTA record found"):
        bs.read_fasta_to_device(b"junk\n", seek_first_rec=True)
    g = rng(3)
    text = b"junk\n" + rand_fasta(g, 7)
    for nrec, skip in [(-1, 0), (3, 0), (-1, 2), (2, 4), (0, 0), (5, 7)]:
        got = gpu_parse(text, nrec=nrec, skip=skip, seek_first_rec=True)
        if ref.available():
            w, c, names, warn = ref.read_fasta(text, nrec=nrec, skip=skip, seek_first_rec=True)
            assert got[0].tolist() == w.tolist() and np.array_equal(got[1], c) and got[2] == names
    # files on disk, plain and gzip-ed, and several at once
    p1, p2 = tmp_path / "a.fa", tmp_path / "b.fa.gz"
    p1.write_bytes(TEXTS[0])
    with gzip.open(p2, "wb") as f:
        f.write(TEXTS[5])
    s = bs.readDNAStringSet([str(p1), str(p2)])
    assert s.names == ["sequence1", "sequence2", "a", "b"] and s.widths.tolist() == [72, 70, 8, 0]
    assert str(s[0]).startswith("AGTACGTAGTCG")
    assert bs.readDNAStringSet(str(p1), use_names=False).names is None


def test_resident_genome_serves_the_matching_calls():
    """read once, query many times: the set as a vcountPDict / oligonucleotideFrequency subject,
    one record as a countPDict / matchPDict subject -- equal to the calls on host sequences."""
    g = rng(11)
    seqs = [random_dna(g, n) for n in (50_000, 0, 120_001, 33)]
    seqs[2][1000:1100] = 15
    text = b"".join(b">s%d\n" % i + bs.decode_dna(s).encode() + b"\n" for i, s in enumerate(seqs))
    probes = [seqs[0][100:125], seqs[2][5000:5025], seqs[2][119_976:120_001], random_dna(g, 25)]
    pdict = bs.PDict(probes)
    host_set = bs.DNAStringSet(concat=np.concatenate(seqs), widths=np.array([s.size for s in seqs], np.int32))
    dset, widths, names = bs.read_fasta_to_device(text)
    with dset:
        assert names == ["s0", "s1", "s2", "s3"] and widths.tolist() == [s.size for s in seqs]
        assert np.array_equal(bs.vcountPDict(pdict, dset), bs.vcountPDict(pdict, host_set))
        a = bs.oligonucleotideFrequency(dset, 4)
        b = bs.oligonucleotideFrequency(host_set, 4)
        assert np.array_equal(np.asarray(a), np.asarray(b))
        with bs.device_subject_element(dset, 2) as chrom:
            assert np.array_equal(bs.countPDict(pdict, chrom), bs.countPDict(pdict, bs.DNAString(seqs[2])))
            m1, m2 = bs.matchPDict(pdict, chrom), bs.matchPDict(pdict, bs.DNAString(seqs[2]))
            assert [None if e is None else e.tolist() for e in m1.endIndex()] == \
                [None if e is None else e.tolist() for e in m2.endIndex()]
            assert np.array_equal(np.asarray(bs.oligonucleotideFrequency(chrom, 6)),
                                  np.asarray(bs.oligonucleotideFrequency(bs.DNAString(seqs[2]), 6)))
        with pytest.raises(bs.BsgpuError, match="subscript out of bounds"):
            bs.device_subject_element(dset, 4)


def test_lines_longer_than_a_group_of_chunks():
    """classes carried over whole groups of 64 chunks (tests/test_fasta_hostlogic.py has the
    same text on the CPU walk)."""
    g = rng(31337)
    seq = np.frombuffer(b"ACGTNacgtX", dtype=np.uint8)[g.integers(0, 10, 150_000)].tobytes()
    text = (b">" + b"d" * 70_000 + b"\n" + seq[:100_000] + b"\n;" + b"c" * 40_000 + b"\r\n" +
            seq[100_000:] + b"\n>r2\n" + b"A" * 33_000 + b"\r\n>r3 " + b"x" * 66_000)
    same(gpu_parse(text), OF.read_fasta(text))
    for cut in (32_768, 32_769, 65_536, 98_304 + 17):
        same(gpu_parse(text[:cut]), OF.read_fasta(text[:cut]))


def test_letters_encoded_on_the_device():
    """DNAString(<character>) with the translation done by encode_letters_kernel
    (bsgpu_subject_letters): same resident subject as uploading host-encoded codes, the
    reference's error text for a letter outside the alphabet (first such letter)."""
    g = rng(21)
    for n in (0, 1, 15, 16, 17, 4095, 100_003):
        letters = np.frombuffer(b"ACGTacgtNnMRWSYKVHDB-+.", dtype=np.uint8)[g.integers(0, 23, n)]
        with bs.device_subject_from_letters(letters) as d:
            got = bs.download_set(d)
        assert got.widths.tolist() == [n]
        assert np.array_equal(got.concat, bs.encode_dna(letters.tobytes()))
    s = random_dna(g, 60_000)
    text = bs.decode_dna(s)
    pdict = bs.PDict([s[100:125], s[59_975:60_000], random_dna(g, 25)])
    with bs.device_subject_from_letters(text) as d:
        assert np.array_equal(bs.countPDict(pdict, d), bs.countPDict(pdict, bs.DNAString(s)))
        assert np.array_equal(np.asarray(bs.oligonucleotideFrequency(d, 5)),
                              np.asarray(bs.oligonucleotideFrequency(bs.DNAString(s), 5)))
    bad = bytearray(text[:5000].encode())
    bad[4000], bad[1234] = ord("!"), ord("Z")
    with pytest.raises(bs.BsgpuError, match=r"key 90 \(char 'Z'\) not in lookup table"):
        bs.device_subject_from_letters(bytes(bad))
    with pytest.raises(ValueError, match=r"key 90 \(char 'Z'\) not in lookup table"):
        bs.encode_dna(bytes(bad))


def test_file_list_nrec_skip_share_one_record_counter(tmp_path):
    """readDNAStringSet(c(f1, f2, f3), nrec, skip): nrec / skip select records of the concatenated
    files (src/read_fasta_files.c:447-458), checked against the reference on the same list."""
    g = rng(8800)
    texts = []
    for k, n in enumerate((3, 1, 5)):
        recs = [f">f{k}r{i} d\n{bs.decode_dna(random_dna(g, int(g.integers(0, 90))))}\n" for i in range(n)]
        texts.append("".join(recs).encode())
    paths = []
    for k, t in enumerate(texts):
        p = tmp_path / f"f{k}.fa"
        p.write_bytes(t)
        paths.append(str(p))
    for nrec, skip in [(-1, 0), (3, 2), (-1, 4), (2, 3), (5, 1), (0, 0), (1, 8), (9, 9)]:
        got = bs.readDNAStringSet(paths, nrec=nrec, skip=skip)
        w, c, names, _ = ref.read_fasta(texts, nrec=nrec, skip=skip)
        assert got.widths.tolist() == w.tolist() and np.array_equal(got.concat, c), (nrec, skip)
        assert (got.names or []) == (names or [])
