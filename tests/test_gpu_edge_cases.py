"""Edge cases of the path on the GPU against the reference: empty and tiny inputs, subjects
shorter than the Trusted Band, single patterns, all-N subjects, patterns longer than the
subject, widths at the 32/64-letter word boundaries, many duplicates, long runs."""
import warnings

import numpy as np
import pytest

import biostrings_b200 as bs
from oracle import ref, rlevel as R
from tests.helpers import assert_same_ends, random_dna, rng

pytestmark = pytest.mark.gpu
BACKEND = "ref" if ref.available() else "port"


def both(pats, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return bs.PDict(bs.DNAStringSet(pats), **kw), R.PDict(R.SeqSet(pats), backend=BACKEND, **kw)


def check(prod, orac, subj_p, subj_o=None, **kw):
    subj_o = subj_p if subj_o is None else subj_o
    assert bs.countPDict(prod, subj_p, **kw).tolist() == R.countPDict(orac, subj_o, **kw).tolist()
    assert bs.whichPDict(prod, subj_p, **kw).tolist() == R.whichPDict(orac, subj_o, **kw).tolist()
    assert_same_ends(bs.matchPDict(prod, subj_p, **kw).endIndex(), R.matchPDict(orac, subj_o, **kw).endIndex())


@pytest.mark.parametrize("n", [0, 1, 3, 11, 12, 13, 31, 32, 33, 63, 64, 65, 127, 128, 129])
def test_tiny_subjects(n):
    g = rng(n)
    s = random_dna(g, 200)
    pats = [s[i:i + 12].copy() for i in range(0, 60, 3)] + [s[5:5 + 12].copy()]
    prod, orac = both(pats)
    check(prod, orac, s[:n])
    prod, orac = both([np.concatenate([p, random_dna(g, 6)]) for p in pats], tb_end=12)
    for mm in (0, 3, 6):
        check(prod, orac, s[:n], max_mismatch=mm)


@pytest.mark.parametrize("w", [20, 25, 31, 32, 33, 40, 63, 64])
def test_word_boundary_widths(w):
    g = rng(1000 + w)
    s = random_dna(g, 4000)
    s[1000:1003] = 15
    pats = [s[i:i + w].copy() for i in g.integers(0, 4000 - w, 150)]
    for p in pats:
        p[p == 15] = 1
    pats += [s[4000 - w:].copy(), s[:w].copy()]          # touching both ends of the subject
    prod, orac = both(pats)
    check(prod, orac, s)
    if w >= 24:
        for p in pats[:80]:
            p[int(g.integers(0, w))] = 2
        prod, orac = both(pats, max_mismatch=1)
        check(prod, orac, s, max_mismatch=1)


def test_single_pattern_and_all_n_subject():
    prod, orac = both(["ACGTACGTACGTACGTACGTACGT"])
    check(prod, orac, "ACGTACGTACGTACGTACGTACGTACGTACGT")
    check(prod, orac, "N" * 300)
    check(prod, orac, "ACGTACGTACGTNCGTACGTACGTACGTACGT" * 3)
    prod, orac = both(["ACGTACGTACGTACGTACGTACGT"], max_mismatch=2)
    check(prod, orac, "ACGTACGTACGTNCGTACGTACGTACGTACGT" * 3, max_mismatch=2)
    check(prod, orac, "ACGTACG", max_mismatch=2)            # subject shorter than the pattern


def test_many_duplicates_and_low_complexity():
    g = rng(77)
    s = np.concatenate([np.full(500, 1, np.uint8), random_dna(g, 1000), np.tile([1, 2], 300).astype(np.uint8)])
    pats = [np.full(20, 1, np.uint8)] * 5 + [np.tile([1, 2], 10).astype(np.uint8)] * 3 + \
           [np.tile([2, 1], 10).astype(np.uint8)] + [s[600:620].copy()] * 2
    prod, orac = both(pats)
    assert prod.duplicated().tolist() == orac.duplicated().tolist()
    check(prod, orac, s)
    prod, orac = both(pats, tb_start=5, tb_width=8)
    check(prod, orac, s, max_mismatch=1)
    prod, orac = both(pats, max_mismatch=2)                 # poly-A: hundreds of hits per pattern
    check(prod, orac, s, max_mismatch=2)
    # a dictionary so repetitive that the q-gram index refuses it (> 254 entries per key)
    rep = [np.full(24, 1, np.uint8).copy() for _ in range(300)]
    for i, p in enumerate(rep):
        p[23] = [1, 2, 4, 8][i % 4]
        p[22] = [1, 2, 4, 8][(i // 4) % 4]
        p[21] = [1, 2, 4, 8][(i // 16) % 4]
        p[20] = [1, 2, 4, 8][(i // 64) % 4]
        p[19] = [1, 2, 4, 8][(i // 256) % 4]
    prod, orac = both(rep, max_mismatch=1)
    check(prod, orac, s, max_mismatch=1)


def test_empty_and_degenerate_sets_and_views():
    g = rng(5)
    s = random_dna(g, 500)
    pats = [s[i:i + 15].copy() for i in range(0, 300, 20)]
    prod, orac = both(pats, tb_end=8)
    empty_p, empty_o = bs.DNAStringSet(["", "", ""]), R.SeqSet(["", "", ""])
    assert bs.vcountPDict(prod, empty_p).tolist() == R.vcountPDict(orac, empty_o).tolist()
    assert [v.tolist() for v in bs.vwhichPDict(prod, empty_p)] == [[], [], []]
    v_p, v_o = bs.XStringViews(s, [1, 1, 500], [0, 500, 1]), R.Views(s, [1, 1, 500], [0, 500, 1])
    check(prod, orac, v_p, v_o, max_mismatch=1)
    one_p, one_o = bs.DNAStringSet([s]), R.SeqSet([s])
    assert (bs.vcountPDict(prod, one_p, max_mismatch=1) == R.vcountPDict(orac, one_o, max_mismatch=1)).all()


def test_large_set_layout_matches_small_path():
    """The device-side layout of sets (scatter kernel) against the oracle on ragged reads."""
    g = rng(9)
    genome = random_dna(g, 30_000)
    reads = [genome[i:i + int(g.integers(0, 300))].copy() for i in g.integers(0, 29_700, 3000)]
    pats = [genome[i:i + 22].copy() for i in g.integers(0, 29_900, 400)]
    prod, orac = both(pats)
    a = bs.vcountPDict(prod, bs.DNAStringSet(reads), collapse=1)
    b = R.vcountPDict(orac, R.SeqSet(reads), collapse=1)
    assert (a == b).all() and a.sum() > 100
    wa = bs.vwhichPDict(prod, bs.DNAStringSet(reads))
    wb = R.vwhichPDict(orac, R.SeqSet(reads))
    assert [v.tolist() for v in wa] == [v.tolist() for v in wb]
