"""Non-preprocessed dictionaries (`pdict` = plain DNAStringSet; SURVEY 8f rank 1:
match_XStringSet_XString / ...XStringViews / vmatch_XStringSet_XStringSet,
src/match_pdict.c:174,267,519) on the GPU against the reference's own C -- any letters, any
widths, mismatches anywhere, out-of-limits matches, the four `fixed` modes -- through the
Python mirror, and through the `.Call`-level dispatch."""
import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import _lib
from oracle import ref, rlevel as R
from tests.helpers import assert_same_ends, random_dna, rng

pytestmark = pytest.mark.gpu
BACKEND = "ref" if ref.available() else "port"
CODES = np.array([1, 2, 4, 8, 1, 2, 4, 8, 1, 2, 4, 8, 15, 3, 5, 10, 16], dtype=np.uint8)
FIXED = [True, False, "pattern", "subject"]


def letters(g, n):
    return CODES[g.integers(0, len(CODES), n)]


def check(pats, subj_p, subj_o, **kw):
    pset, oset = bs.DNAStringSet(pats), R.SeqSet(pats)
    okw = dict(kw, backend=BACKEND)
    assert bs.countPDict(pset, subj_p, **kw).tolist() == R.countPDict(oset, subj_o, **okw).tolist()
    assert _lib.lib().bsgpu_last_plan().decode().startswith("plain_bruteforce")
    assert bs.whichPDict(pset, subj_p, **kw).tolist() == R.whichPDict(oset, subj_o, **okw).tolist()
    mp, mo = bs.matchPDict(pset, subj_p, **kw), R.matchPDict(oset, subj_o, **okw)
    assert_same_ends(mp.endIndex(), mo.endIndex())
    assert_same_ends(mp.startIndex(), mo.startIndex())


def test_reference_golden_vectors_on_the_gpu():
    """tests/testthat/test-matchPattern.R:55-125 and test-matchPDict.R:15-17, product only."""
    dna = "ATGCATGCATGC"

    def se(pat, subject, **kw):
        m = bs.matchPDict(bs.DNAStringSet([pat]), subject, **kw)
        s, e = m.startIndex()[0], m.endIndex()[0]
        return ([] if s is None else s.tolist(), [] if e is None else e.tolist())

    for algo in ("auto", "naive-exact", "naive-inexact", "boyer-moore", "shift-or"):
        assert se("ATG", dna, algorithm=algo) == ([1, 5, 9], [3, 7, 11])
        assert se("ATG", bs.MaskedDNAString(dna, [2, 9], [4, 11]), algorithm=algo) == ([5], [7])
        assert se("ATG", bs.XStringViews(dna, [2, 9], [3, 3]), algorithm=algo) == ([9], [11])
    assert se("ATC", dna, max_mismatch=2) == ([1, 2, 5, 6, 9, 10], [3, 4, 7, 8, 11, 12])
    assert se("ATC", dna, max_mismatch=2, min_mismatch=2) == ([2, 6, 10], [4, 8, 12])
    dna2 = "ACGTNMRWSYKVHDB"
    assert se("GTN", dna2, fixed=True) == ([3], [5])
    assert se("GTN", dna2, fixed=False) == ([3, 7, 9, 12], [5, 9, 11, 14])
    assert se("---", "ACGTGCA", max_mismatch=3) == (list(range(-1, 8)), list(range(1, 10)))
    assert se("---", "A") == ([], [])
    assert se("ATA", "ATGGATACACAA") == ([5], [7])
    with pytest.raises(ValueError, match="valid algos for your string matching problem"):
        se("ATG", dna, algorithm="indels")
    assert bs.vcountPDict(bs.DNAStringSet(["TG"]), bs.XStringViews(dna, [2, 9], [3, 3])).tolist() == [[1, 1]]


def test_reference_examples():
    # matchPattern edge cases quoted in R/matchPattern.R:183-185, through a one-element set
    check(["TG"], "GTGACGTGCAT", "GTGACGTGCAT")
    check(["TG"], "GTGACGTGCAT", "GTGACGTGCAT", max_mismatch=1)
    check(["---"], "ACGTGCA", "ACGTGCA", max_mismatch=3)
    check(["---"], "A", "A")
    # the codon dictionary of tests/testthat/test-matchPDict.R, not preprocessed
    check(["AAA", "ACC", "NNN", "ACN"], "AAACCNACCAAA", "AAACCNACCAAA", fixed=False)
    check(["AAA", "ACC", "NNN", "ACN"], "AAACCNACCAAA", "AAACCNACCAAA", fixed=True)


@pytest.mark.parametrize("seed", range(6))
def test_random_small_against_the_reference(seed):
    g = rng(9100 + seed)
    for _ in range(25):
        s = letters(g, int(g.integers(0, 70)))
        pats = [letters(g, int(g.integers(1, 10))) for _ in range(int(g.integers(1, 6)))]
        mx = int(g.integers(0, 4))
        mn = int(g.integers(0, mx + 1))
        check(pats, s, s, max_mismatch=mx, min_mismatch=mn, fixed=FIXED[int(g.integers(0, 4))])


def test_longer_subject_variable_widths_and_iupac_patterns():
    g = rng(9200)
    s = random_dna(g, 30_000)
    s[g.integers(0, s.size, 60)] = 15
    pats = []
    for _ in range(120):
        w = int(g.integers(8, 45))
        i = int(g.integers(0, s.size - w))
        p = s[i:i + w].copy()
        for k in g.integers(0, w, int(g.integers(0, 4))):
            p[k] = [15, 3, 5, 1, 2][int(g.integers(0, 5))]
        pats.append(p)
    pats += [s[:20].copy(), s[-20:].copy(), random_dna(g, 70)]
    for fixed in FIXED:
        check(pats, s, s, max_mismatch=2, fixed=fixed)
    check(pats, s, s, max_mismatch=3, min_mismatch=2, fixed=False)
    check(pats, s, s)
    with pytest.raises(ValueError, match="valid algos"):
        bs.countPDict(bs.DNAStringSet(pats), s, max_mismatch=1, algorithm="boyer-moore")
    check(pats, s, s, max_mismatch=1, algorithm="naive-inexact")


def test_views_and_masked_subjects():
    g = rng(9300)
    s = random_dna(g, 5000)
    pats = [s[i:i + int(g.integers(4, 20))].copy() for i in g.integers(0, 4900, 40)]
    starts = np.array([1, 100, 90, 4000, 4990, 2500], dtype=np.int32)
    widths = np.array([120, 50, 400, 1000, 11, 0], dtype=np.int32)
    for mm in (0, 2):
        check(pats, bs.XStringViews(s, starts, widths), R.Views(s, starts, widths), max_mismatch=mm)
    check(pats, bs.MaskedDNAString(s, [50, 3000], [2000, 3500]), R.Masked(s, [50, 3000], [2000, 3500]),
          max_mismatch=1)
    with pytest.raises(bs.BsgpuError, match="out of limits"):
        bs.countPDict(bs.DNAStringSet(pats), bs.XStringViews(s, [4990], [20]))


def test_vcount_vwhich_on_sets():
    g = rng(9400)
    genome = random_dna(g, 8000)
    reads = [genome[i:i + int(g.integers(0, 90))].copy() for i in g.integers(0, 7900, 150)]
    reads[5] = reads[5][:0]
    pats = [genome[i:i + int(g.integers(5, 25))].copy() for i in g.integers(0, 7900, 30)]
    pats[3][2] = 15
    pset, oset = bs.DNAStringSet(pats), R.SeqSet(pats)
    rp, ro = bs.DNAStringSet(reads), R.SeqSet(reads)
    for kw in ({}, {"max_mismatch": 2}, {"max_mismatch": 2, "min_mismatch": 1, "fixed": False}):
        okw = dict(kw, backend=BACKEND)
        a, b = bs.vcountPDict(pset, rp, **kw), R.vcountPDict(oset, ro, **okw)
        assert a.shape == b.shape and (a == b).all()
        wa, wb = bs.vwhichPDict(pset, rp, **kw), R.vwhichPDict(oset, ro, **okw)
        assert [v.tolist() for v in wa] == [v.tolist() for v in wb]
        w1 = g.integers(1, 5, len(reads)).astype(np.int32)
        w2 = g.integers(1, 5, len(pats)).astype(np.int32)
        assert (bs.vcountPDict(pset, rp, collapse=1, weight=w1, **kw) ==
                R.vcountPDict(oset, ro, collapse=1, weight=w1, **okw)).all()
        assert (bs.vcountPDict(pset, rp, collapse=2, weight=w2, **kw) ==
                R.vcountPDict(oset, ro, collapse=2, weight=w2, **okw)).all()
        d1 = g.random(len(reads)) * 3
        x, y = bs.vcountPDict(pset, rp, collapse=1, weight=d1, **kw), \
            R.vcountPDict(oset, ro, collapse=1, weight=d1, **okw)
        assert x.dtype == np.float64 and (x == y).all()


def test_errors():
    with pytest.raises(ValueError, match="empty patterns are not supported"):
        bs.countPDict(bs.DNAStringSet(["ACGT", ""]), "ACGT")
    with pytest.raises(ValueError, match="only support indels"):
        bs.matchPDict(bs.PDict(["ACGT", "GGGG"]), "ACGT", max_mismatch=1, with_indels=True)
    with pytest.raises(ValueError, match="support indels"):
        bs.matchPDict(bs.DNAStringSet(["ACGT"]), "ACGT", max_mismatch=1, with_indels=True)
    # with.indels=TRUE is served by the indels engine (tests/test_gpu_indels.py)
    assert bs.countPDict(bs.DNAStringSet(["ACGT"]), "ACGT", max_mismatch=1, with_indels=True).tolist() == [1]
    with pytest.raises(ValueError, match="XStringSet object"):
        bs.countPDict(bs.DNAStringSet(["ACGT"]), bs.DNAStringSet(["ACGT"]))


@pytest.mark.skipif(not (ref.available("ref") and ref.available("dispatch")),
                    reason="needs the reference and dispatch libraries")
def test_through_the_dot_call_boundary():
    """match_XStringSet_XString & co of the dispatch library against the reference's, SEXP by
    SEXP (NULL placement of empty rows included)."""
    g = rng(9500)
    s = letters(g, 400)
    pats = [letters(g, int(g.integers(1, 12))) for _ in range(15)]
    pc, pw = np.concatenate(pats), [len(p) for p in pats]
    starts, widths = np.array([1, 50, 300], np.int32), np.array([100, 0, 101], np.int32)
    reads = [letters(g, int(g.integers(0, 40))) for _ in range(12)]
    rc, rw = np.concatenate(reads), [len(r) for r in reads]
    out = {}
    for which in ("ref", "dispatch"):
        ref.use(which)
        try:
            res = []
            for mx, mn, fixed, algo in ((0, 0, (1, 1), "boyer-moore"), (2, 0, (0, 0), "shift-or"),
                                        (3, 1, (1, 0), "naive-inexact"), (1, 0, (0, 1), "naive-inexact")):
                for mode in (ref.MATCHES_AS_WHICH, ref.MATCHES_AS_COUNTS, ref.MATCHES_AS_ENDS):
                    res.append(ref.match_set_xstring(pc, pw, s, mx, mn, 0, fixed, algo, mode))
                    res.append(ref.match_set_views(pc, pw, s, starts, widths, mx, mn, 0, fixed, algo, mode))
                for coll, w in ((0, [1]), (1, np.arange(1, 13)), (2, np.arange(1, 16) / 2.0)):
                    res.append(ref.vmatch_set_set(pc, pw, rc, rw, mx, mn, 0, fixed, algo, coll, w,
                                                  ref.MATCHES_AS_COUNTS))
                res.append(ref.vmatch_set_set(pc, pw, rc, rw, mx, mn, 0, fixed, algo, 0, [1],
                                              ref.MATCHES_AS_WHICH))
            out[which] = res
        finally:
            ref.use("ref")
    assert len(out["ref"]) == len(out["dispatch"])
    for a, b in zip(out["ref"], out["dispatch"]):
        if isinstance(a, list):
            assert_same_ends(a, b)
        else:
            assert np.array_equal(np.asarray(a), np.asarray(b))
    ref.use("dispatch")
    try:
        # with.indels: patterns of 1 letter (<= max.mismatch) take the mismatch-only loop
        got = ref.match_set_xstring(pc, pw, s, 1, 0, 1, (1, 1), "indels", ref.MATCHES_AS_COUNTS)
        ref.use("ref")
        want = ref.match_set_xstring(pc, pw, s, 1, 0, 1, (1, 1), "indels", ref.MATCHES_AS_COUNTS)
        ref.use("dispatch")
        assert np.array_equal(got, want)
        with pytest.raises(ref.RefError, match="empty patterns"):
            ref.match_set_xstring(pc[:3], [3, 0], s, 0, 0, 0, (1, 1), "naive-exact", ref.MATCHES_AS_COUNTS)
    finally:
        ref.use("ref")


def test_chunked_subject_equals_whole():
    """Overlap chunks for a non-preprocessed dictionary: a match belongs to the chunk
    holding its last letter; only the ends of the whole subject are out of limits."""
    from biostrings_b200 import matchpdict as mp
    g = rng(9600)
    s = letters(g, 6000)
    pats = [s[i:i + int(g.integers(3, 30))].copy() for i in g.integers(0, 5900, 60)]
    pats += [s[:12].copy(), s[-12:].copy()]                # overhang both ends with mismatches
    pset = bs.DNAStringSet(pats)
    maxw = max(len(p) for p in pats)
    pat = mp.PlainPatterns(pset)
    try:
        for mx, fixed in ((0, (True, True)), (3, (False, False))):
            with mp.DeviceSubject.from_xstring(s) as whole:
                want = mp.match_parts([pat], whole, mx, 0, fixed, _lib.MATCHES_AS_COUNTS)
                wo, we = mp.match_parts([pat], whole, mx, 0, fixed, _lib.MATCHES_AS_ENDS)
            bounds = [0, 1, 40, 2048, 2049, 5990, s.size]
            got = np.zeros_like(want)
            rows = [[] for _ in pats]
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                a = max(0, lo - (maxw - 1))
                with mp.DeviceSubject.from_chunk(s[a:hi], a, s.size, lo, hi) as ch:
                    got += mp.match_parts([pat], ch, mx, 0, fixed, _lib.MATCHES_AS_COUNTS)
                    o, e = mp.match_parts([pat], ch, mx, 0, fixed, _lib.MATCHES_AS_ENDS)
                    for i in range(len(pats)):
                        rows[i].extend(e[o[i]:o[i + 1]].tolist())
            assert (got == want).all() and want.sum() > 60
            for i in range(len(pats)):
                assert rows[i] == we[wo[i]:wo[i + 1]].tolist()
        with mp.DeviceSubject.from_chunk(s[1000:2000], 1000, s.size, 1000, 2000) as ch:
            with pytest.raises(bs.BsgpuError, match="insufficient halo"):
                mp.match_parts([pat], ch, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS)
    finally:
        pat.close()


@pytest.mark.parametrize("widths", [(12, 40), (25, 25), (40, 64)])
def test_large_base_only_dictionary_gets_its_own_index(widths, monkeypatch):
    """>= 256 base-only patterns matched exactly: the library builds a Trusted-Band index behind the
    non-preprocessed dictionary (plain_auto_index) -- counts, which and ends must equal the
    reference's one-matchPattern-per-pattern loop, whole-pattern duplicates included, on a
    sequence, on views, on a set, and equal the brute force the same call runs with
    BSGPU_PLAIN_INDEX=0.  Anything the index does not cover (mismatches, IUPAC subject) stays brute force."""
    g = rng(9700 + widths[0])
    s = random_dna(g, 60_000)
    s[1000:1010] = 15
    pats = []
    for i in g.integers(0, 59_000, 500):
        w = int(g.integers(widths[0], widths[1] + 1))
        p = s[i:i + w].copy()
        if not np.isin(p, [1, 2, 4, 8]).all():
            p = random_dna(g, w)
        pats.append(p)
    pats += [random_dna(g, int(g.integers(widths[0], widths[1] + 1))) for _ in range(100)]
    pats += [pats[int(k)].copy() for k in g.integers(0, 500, 40)]          # whole-pattern duplicates
    pats += [pats[int(k)][:widths[0]].copy() for k in g.integers(0, 500, 20)]  # prefixes of other patterns
    pset, oset = bs.DNAStringSet(pats), R.SeqSet(pats)
    L = _lib.lib()
    subjects = [(s, s), (bs.XStringViews(s, [1, 500, 30_000], [20_000, 3000, 30_000]),
                         R.Views(s, [1, 500, 30_000], [20_000, 3000, 30_000]))]
    for sp, so in subjects:
        want = R.countPDict(oset, so, backend=BACKEND)
        got = bs.countPDict(pset, sp)
        assert not L.bsgpu_last_plan().decode().startswith("plain_bruteforce")
        assert got.tolist() == want.tolist() and got.sum() > 300
        assert bs.whichPDict(pset, sp).tolist() == R.whichPDict(oset, so, backend=BACKEND).tolist()
        mp_, mo = bs.matchPDict(pset, sp), R.matchPDict(oset, so, backend=BACKEND)
        assert_same_ends(mp_.endIndex(), mo.endIndex())
        assert_same_ends(mp_.startIndex(), mo.startIndex())
    # a set of reads
    reads = [s[i:i + int(g.integers(0, 200))].copy() for i in g.integers(0, 59_000, 200)]
    rp, ro = bs.DNAStringSet(reads), R.SeqSet(reads)
    a, b = bs.vcountPDict(pset, rp), R.vcountPDict(oset, ro, backend=BACKEND)
    assert not L.bsgpu_last_plan().decode().startswith("plain_bruteforce")
    assert a.shape == b.shape and (a == b).all()
    # what the index does not cover keeps the brute force
    assert bs.countPDict(pset, s, max_mismatch=1).tolist() == R.countPDict(oset, s, max_mismatch=1, backend=BACKEND).tolist()
    assert L.bsgpu_last_plan().decode().startswith("plain_bruteforce")
    # and the brute force agrees when the index is switched off (a fresh dictionary object)
    monkeypatch.setenv("BSGPU_PLAIN_INDEX", "0")
    pset2 = bs.DNAStringSet(pats)
    got2 = bs.countPDict(pset2, s)
    assert L.bsgpu_last_plan().decode().startswith("plain_bruteforce")
    assert got2.tolist() == R.countPDict(oset, s, backend=BACKEND).tolist()
