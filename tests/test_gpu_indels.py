"""`with.indels=TRUE` on the GPU (bsgpu_count_indels) against the reference's own
_match_pattern_indels (src/match_pattern_indels.c:58-107).  The engine's per-position core is
also checked on the CPU (tests/test_indels_oracle.py)."""
import os

import numpy as np
import pytest

import biostrings_b200 as bs
from oracle import indels as OI
from oracle import ref
from tests.helpers import random_dna, rng

pytestmark = [pytest.mark.gpu]


def want_counts(pats, S, k, fixed=(True, True)):
    if ref.available():
        concat = np.concatenate(pats)
        widths = np.array([len(p) for p in pats], dtype=np.int32)
        return ref.match_set_xstring(concat, widths, S, k, 0, True, fixed, "indels",
                                     ref.MATCHES_AS_COUNTS).tolist()
    return OI.count_indels(pats, S, k, *fixed)


def planted(g, S, n, k):
    pats = []
    for _ in range(n):
        w = int(g.integers(k + 1, 30))
        a = int(g.integers(0, len(S) - 40))
        p = S[a:a + w].copy()
        for _ in range(int(g.integers(0, k + 2))):
            kind, at = int(g.integers(3)), int(g.integers(0, len(p)))
            if kind == 0:
                p[at] = [1, 2, 4, 8][int(g.integers(4))]
            elif kind == 1 and len(p) > k + 1:
                p = np.delete(p, at)
            else:
                p = np.insert(p, at, [1, 2, 4, 8][int(g.integers(4))])
        pats.append(p.astype(np.uint8))
    return pats


@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_counts_and_which(k):
    g = rng(600 + k)
    S = random_dna(g, 20_000)
    S[g.integers(0, len(S), 100)] = 15
    pats = planted(g, S, 80, k) + [S[:k + 1].copy(), S[-(k + 3):].copy()]
    want = want_counts(pats, S, k)
    pset = bs.DNAStringSet(pats)
    got = bs.countPDict(pset, bs.DNAString(S), max_mismatch=k, with_indels=True)
    assert got.tolist() == want
    which = bs.whichPDict(pset, bs.DNAString(S), max_mismatch=k, with_indels=True)
    assert which.tolist() == [i + 1 for i, c in enumerate(want) if c]


@pytest.mark.parametrize("fixed", [True, False, "pattern", "subject"])
def test_fixed_modes(fixed):
    g = rng(10)
    codes = np.array([1, 2, 4, 8, 1, 2, 4, 8, 15, 3, 5, 10], dtype=np.uint8)
    S = codes[g.integers(0, len(codes), 5000)]
    pats = [codes[g.integers(0, len(codes), int(g.integers(4, 16)))] for _ in range(30)]
    pats += [S[a:a + 12].copy() for a in g.integers(0, 4900, 10)]
    pair = {True: (True, True), False: (False, False), "pattern": (True, False), "subject": (False, True)}[fixed]
    got = bs.countPDict(bs.DNAStringSet(pats), bs.DNAString(S), max_mismatch=3, with_indels=True, fixed=fixed)
    assert got.tolist() == want_counts(pats, S, 3, pair)


def test_views_and_sets_are_independent_subjects():
    g = rng(11)
    S = random_dna(g, 6000)
    pats = planted(g, S, 40, 2)
    starts, widths = [1, 2500, 2400, 5990], [2000, 1000, 0, 11]
    want = np.zeros(len(pats), dtype=np.int64)
    per_view = []
    for st, w in zip(starts, widths):
        c = np.array(want_counts(pats, S[st - 1:st - 1 + w], 2))
        per_view.append(c)
        want += c
    pset = bs.DNAStringSet(pats)
    got = bs.countPDict(pset, bs.XStringViews(bs.DNAString(S), starts, widths), max_mismatch=2, with_indels=True)
    assert got.tolist() == want.tolist()
    sset = bs.DNAStringSet([S[st - 1:st - 1 + w] for st, w in zip(starts, widths)])
    mat = bs.vcountPDict(pset, sset, max_mismatch=2, with_indels=True)
    assert np.array_equal(mat, np.stack(per_view, axis=1))
    assert bs.vcountPDict(pset, sset, max_mismatch=2, with_indels=True, collapse=1).tolist() == want.tolist()
    vw = bs.vwhichPDict(pset, sset, max_mismatch=2, with_indels=True)
    assert [v.tolist() for v in vw] == [(np.flatnonzero(c) + 1).tolist() for c in per_view]


def test_limits():
    # a pattern not longer than max.mismatch is counted by the mismatch-only loop
    # (src/match_pattern.c:95-96), out-of-limits matches included
    S = bs.encode_dna("ACGTACGTTTACGGA")
    pats = [bs.encode_dna(p) for p in ("AC", "ACGTACGT", "G", "TTT")]
    got = bs.countPDict(bs.DNAStringSet(pats), S, max_mismatch=2, with_indels=True)
    assert got.tolist() == want_counts(pats, S, 2)
    sset = bs.DNAStringSet([S[:7], S[5:], S[:0]])
    mat = bs.vcountPDict(bs.DNAStringSet(pats), sset, max_mismatch=2, with_indels=True)
    assert mat.tolist() == np.stack([want_counts(pats, S[:7], 2), want_counts(pats, S[5:], 2),
                                     want_counts(pats, S[:0], 2)], axis=1).tolist()
    with pytest.raises(ValueError, match="only countPDict"):
        bs.matchPDict(bs.DNAStringSet(["ACGT"]), "ACGT", max_mismatch=1, with_indels=True)
