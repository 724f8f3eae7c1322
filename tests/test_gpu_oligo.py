"""oligonucleotideFrequency on the GPU (oligo_freq_kernel through the C ABI)
against the reference's own C (src/letter_frequency.c:634,667 compiled in oracle/_ref) and the
numpy restatement: every step regime, both two-bit orders, as.prob, matrix / list / collapsed,
views and masked subjects, labels, sharded chunks, the reference's
countPDict(PDict(all k-mers)) identity (R/letterFrequency.R:541-553) and its error texts."""
import ctypes as C

import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import _lib
from biostrings_b200.letterfreq import oligo_frequency_device
from oracle import letterfreq as LF
from oracle import ref
from tests.helpers import random_dna, rng
from tests.test_oligo_oracle import all_strings, noisy_dna

pytestmark = pytest.mark.gpu
HAVE_REF = ref.available()


def want_xstring(x, width, step=1, as_prob=False, invert=False):
    want = LF.xstring_oligo_frequency(x, width, step, as_prob=as_prob, invert=invert)
    if HAVE_REF and width <= 12:
        r = ref.oligo_frequency_xstring(x, width, step, as_prob=as_prob, with_labels=False,
                                        fast_moving_side="left" if invert else "right")[0]
        assert np.array_equal(r, want)
    return want


@pytest.mark.parametrize("width,step", [(1, 1), (1, 4), (2, 1), (2, 2), (3, 1), (3, 2), (3, 3), (4, 9),
                                        (5, 1), (6, 1), (6, 5), (7, 1), (7, 3), (8, 1), (9, 1), (8, 3), (10, 7),
                                        (12, 1), (13, 2)])
@pytest.mark.parametrize("invert", [False, True])
def test_xstring_counts_and_probabilities(width, step, invert):
    g = rng(500 + 17 * width + step)
    x = noisy_dna(g, 100_003)
    side = "left" if invert else "right"
    got = bs.oligonucleotideFrequency(bs.DNAString(x), width, step=step, fast_moving_side=side,
                                      with_labels=width <= 6)
    assert got.dtype == np.int32
    assert np.array_equal(np.asarray(got), want_xstring(x, width, step, invert=invert))
    assert _lib.lib().bsgpu_last_plan().decode().startswith(f"oligo_freq w={width} step={step} rows=1")
    if width <= 6:
        assert got.names == all_strings(width, invert)
    p = bs.oligonucleotideFrequency(bs.DNAString(x), width, step=step, as_prob=True,
                                    fast_moving_side=side, with_labels=False)
    assert p.dtype == np.float64 and p.names is None
    assert np.array_equal(np.asarray(p), want_xstring(x, width, step, as_prob=True, invert=invert))


def test_edge_sequences():
    for s in ["", "A", "ACG", "NNNNNNNN", "ACGTN", "NACGT", "AC-GT+AC.GT"]:
        x = bs.DNAString(s)
        for width in (1, 2, 3, 5):
            for step in (1, 2, 5):
                got = bs.oligonucleotideFrequency(x, width, step=step, with_labels=False)
                assert np.array_equal(np.asarray(got), LF.xstring_oligo_frequency(x.codes, width, step)), \
                    (s, width, step)
    # all-zero table with as.prob stays all zero (normalize_oligo_freqs skips the row)
    p = bs.oligonucleotideFrequency("NNNN", 2, as_prob=True)
    assert not np.asarray(p).any()


def test_word_boundaries_and_long_runs():
    """windows straddling the 32-letter words of the planes; runs of N longer than a word."""
    g = rng(5)
    x = random_dna(g, 4096)
    x[100:100 + 77] = 15
    x[1000] = 16
    x[2047:2049] = 15
    for width in (1, 2, 7, 9, 11):
        got = oligo_counts_product(x, width)
        assert np.array_equal(got, LF.oligo_counts(x, width).astype(np.int32))


def oligo_counts_product(x, width, step=1, invert=False):
    with bs.DeviceSubject.from_xstring(x) as d:
        return oligo_frequency_device(d, width, step, invert, collapsed=True)


@pytest.mark.parametrize("simplify_as", ["matrix", "list", "collapsed"])
@pytest.mark.parametrize("as_prob", [False, True])
def test_sets(simplify_as, as_prob):
    g = rng(78)
    widths = g.integers(0, 300, 2000).astype(np.int32)
    widths[[0, 7, 1999]] = 0
    concat = noisy_dna(g, int(widths.sum()))
    sset = bs.DNAStringSet(concat=concat, widths=widths)
    for width, step in [(1, 1), (2, 1), (3, 3), (4, 2), (6, 1), (7, 5)]:
        names = all_strings(width)
        got = bs.oligonucleotideFrequency(sset, width, step=step, as_prob=as_prob, simplify_as=simplify_as)
        want = LF.xstringset_oligo_frequency(concat, widths, width, step, as_prob=as_prob,
                                             simplify_as=simplify_as)
        if HAVE_REF:
            r, dim, _ = ref.oligo_frequency_set(concat, widths, width, step, as_prob=as_prob,
                                                simplify_as=simplify_as, with_labels=False)
            if simplify_as == "matrix":
                assert np.array_equal(r.reshape(dim, order="F"), want)
            elif simplify_as == "collapsed":
                assert np.array_equal(r, want)
        if simplify_as == "matrix":
            assert got.shape == (len(widths), 4 ** width)
            assert got.dimnames == (None, names)
            assert np.array_equal(np.asarray(got), want)
        elif simplify_as == "list":
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert np.array_equal(np.asarray(a), b) and a.names == names
        else:
            assert np.array_equal(np.asarray(got), want) and got.names == names


def test_reference_expectations_on_the_gpu():
    """tests/testthat/test-letterFrequency.R:229-292 re-expressed (see tests/test_oligo_oracle.py)."""
    from collections import Counter
    g = rng(999)
    ln = 102
    d = random_dna(g, ln)
    ds = bs.decode_dna(d)
    x = bs.DNAString(d)

    def table(s, width, step=1):
        c = Counter(s[i:i + width] for i in range(0, len(s) - width + 1, step))
        return {w: c.get(w, 0) for w in all_strings(width)}

    assert bs.oligonucleotideFrequency(x, 1).as_dict() == {c: ds.count(c) for c in "ACGT"}
    assert bs.oligonucleotideFrequency(x, 2).as_dict() == table(ds, 2)
    assert bs.dinucleotideFrequency(x).as_dict() == table(ds, 2)
    assert bs.oligonucleotideFrequency(x, 3).as_dict() == table(ds, 3)
    assert bs.trinucleotideFrequency(x).as_dict() == table(ds, 3)
    assert bs.oligonucleotideFrequency(x, 3, step=3).as_dict() == table(ds, 3, step=3)
    # Views (codons): one row per view
    codons = bs.XStringViews(x, np.arange(1, ln, 3), 3)
    mat = bs.oligonucleotideFrequency(codons, 3)
    names = all_strings(3)
    exp = np.zeros((ln // 3, 64), dtype=np.int32)
    for i in range(ln // 3):
        exp[i, names.index(ds[3 * i:3 * i + 3])] = 1
    assert np.array_equal(np.asarray(mat), exp) and mat.dimnames == (None, names)
    # MaskedXString: collapsed over the unmasked gaps
    masked = ds[:9] + "#" * 10 + ds[19:49] + "#" * 8 + ds[57:]
    want = Counter(masked[i:i + 2] for i in range(ln - 1))
    got = bs.oligonucleotideFrequency(bs.MaskedDNAString(x, [10, 50], [19, 57]), 2)
    assert got.as_dict() == {w: want.get(w, 0) for w in all_strings(2)}
    # as.array: first letter = first index, labels on every dimension
    arr = bs.oligonucleotideFrequency(x, 3, as_array=True)
    assert arr.shape == (4, 4, 4) and arr.dimnames == (list("ACGT"),) * 3
    t3 = table(ds, 3)
    for w, n in t3.items():
        assert arr["ACGT".index(w[0]), "ACGT".index(w[1]), "ACGT".index(w[2])] == n
    # oligonucleotideTransitions: rows = left word
    tr = bs.oligonucleotideTransitions(x, 1, 2)
    assert tr.shape == (4, 16)
    assert tr[1, 4 * 2 + 3] == t3["CGT"]


def test_count_pdict_identity():
    """countPDict(PDict(all k-mers), x) == oligonucleotideFrequency(x, k, with.labels=FALSE)
    (R/letterFrequency.R:541-553)."""
    g = rng(31)
    x = bs.DNAString(noisy_dna(g, 300_000))
    for k in (3, 6, 8):
        pdict = bs.PDict(bs.mkAllStrings(bs.DNA_BASES, k))
        c1 = bs.countPDict(pdict, x)
        c2 = bs.oligonucleotideFrequency(x, k, with_labels=False)
        assert np.array_equal(np.asarray(c1), np.asarray(c2))


def test_chunks_add_up_to_the_whole():
    """A sharded sequence: every chunk counts the windows ending in its report range, step
    judged in whole-subject coordinates (include/bsgpu.h)."""
    g = rng(64)
    x = noisy_dna(g, 200_001)
    L = x.size
    for width, step in [(3, 1), (5, 3), (8, 7), (12, 1)]:
        total = np.zeros(4 ** width, dtype=np.int64)
        bounds = [0, 50_000, 50_007, 131_072, L]
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            a = max(lo - (width - 1), 0)
            with bs.DeviceSubject.from_chunk(x[a:hi], a, L, lo, hi) as d:
                total += oligo_frequency_device(d, width, step, False, collapsed=True)
        assert np.array_equal(total, LF.oligo_counts(x, width, step))


def test_device_side_counts_accumulate():
    import torch
    g = rng(8)
    x = noisy_dna(g, 1_000_000)
    acc = torch.zeros(4 ** 6, dtype=torch.int64, device="cuda")
    with bs.DeviceSubject.from_xstring(x) as d:
        for _ in range(3):
            _lib.check(_lib.lib().bsgpu_oligo_count_device(d.handle, 6, 1, 0, C.c_void_p(acc.data_ptr()),
                                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert np.array_equal(acc.cpu().numpy(), 3 * LF.oligo_counts(x, 6))


def test_host_buffer_entry_points():
    g = rng(9)
    x = noisy_dna(g, 50_000)
    out = np.zeros(4 ** 4, dtype=np.int32)
    L = _lib.lib()
    _lib.check(L.bsgpu_XString_oligo_frequency(x.ctypes.data_as(C.c_void_p), x.size, 4, 2, 0, 0,
                                               out.ctypes.data_as(C.c_void_p), None))
    assert np.array_equal(out, LF.xstring_oligo_frequency(x, 4, 2))
    widths = np.array([10_000, 0, 25_000, 15_000], dtype=np.int32)
    p = np.zeros(4 * 4 ** 3, dtype=np.float64)
    _lib.check(L.bsgpu_XStringSet_oligo_frequency(x.ctypes.data_as(C.c_void_p), widths.ctypes.data_as(C.c_void_p),
                                                  4, 3, 1, 1, 0, 1, None, p.ctypes.data_as(C.c_void_p)))
    want = LF.xstringset_oligo_frequency(x, widths, 3, as_prob=True, invert=True)
    assert np.array_equal(p.reshape((4, 64), order="F"), want)


def test_errors():
    x = bs.DNAString("ACGTACGT")
    with pytest.raises(bs.BsgpuError, match="'buflength' must be >= 1 and <= 15"):
        bs.oligonucleotideFrequency(x, 16)
    with pytest.raises(ValueError, match="'width' must be an integer >= 1"):
        bs.oligonucleotideFrequency(x, 0)
    with pytest.raises(ValueError, match="'step' must be a positive integer"):
        bs.oligonucleotideFrequency(x, 2, step=0)
    with pytest.raises(ValueError, match="'as.array' cannot be TRUE when 'simplify.as' is \"matrix\""):
        bs.oligonucleotideFrequency(bs.DNAStringSet(["ACGT"]), 2, as_array=True)
    with pytest.raises(NotImplementedError, match="not ready yet"):
        bs.vmatchPDict(bs.PDict(["ACGT"]), bs.DNAStringSet(["ACGT"]))
