"""GPU parity on the reference's own golden vectors (through the C ABI).

Vectors: tests/testthat/test-matchPDict.R:11-61,63-76,78-123,125-145 and the worked
examples of man/matchPDict-inexact.Rd:123-220.
"""
import numpy as np
import pytest

import biostrings_b200 as bs
from tests.helpers import ends_as_lists

pytestmark = pytest.mark.gpu


def L(x):
    return ends_as_lists(x)


def test_inexact_counts_and_ends():
    pd = bs.PDict(["acgt", "gt", "cgt", "ac"], tb_end=2)
    d = "acggaccg"
    assert bs.countPDict(pd, d, max_mismatch=0).tolist() == [0, 0, 0, 2]
    assert bs.countPDict(pd, d, max_mismatch=1).tolist() == [1, 0, 2, 2]
    assert bs.countPDict(pd, d, max_mismatch=2).tolist() == [2, 0, 2, 2]
    assert L(bs.matchPDict(pd, d).endIndex()) == [None, None, None, [2, 6]]
    assert L(bs.matchPDict(pd, d, max_mismatch=1).endIndex()) == [[4], None, [4, 9], [2, 6]]
    assert L(bs.matchPDict(pd, d, max_mismatch=2).endIndex()) == [[4, 8], None, [4, 9], [2, 6]]
    assert bs.whichPDict(pd, d, max_mismatch=1).tolist() == [1, 3, 4]


@pytest.mark.parametrize("algo", ["ACtree2", "Twobit"])
def test_mindex_at(algo):
    pd = bs.PDict({"AA": "AA", "AT": "AT", "TA": "TA", "TT": "TT"}, algorithm=algo)
    m = bs.matchPDict(pd, "ATTATTAATTAA")
    assert len(m) == 4 and m.names() == ["AA", "AT", "TA", "TT"]
    assert L(m.startIndex()) == [[7, 11], [1, 4, 8], [3, 6, 10], [2, 5, 9]]
    assert L(m.endIndex()) == [[8, 12], [2, 5, 9], [4, 7, 11], [3, 6, 10]]
    assert m.elementNROWS().tolist() == [2, 3, 3, 3]
    s, e, w = m[0]
    assert s.tolist() == [7, 11] and e.tolist() == [8, 12] and w.tolist() == [2, 2]


def test_all_codons_xstring_views_masked():
    import itertools
    codons = ["".join(c) for c in itertools.product("ACGT", repeat=3)]
    pd = bs.PDict(codons)
    d1 = "ATGGATACACAA"
    win = [d1[i:i + 3] for i in range(len(d1) - 2)]
    good = [i + 1 for i, c in enumerate(codons) if c in win]
    cnt = bs.countPDict(pd, d1)
    assert bs.whichPDict(pd, d1).tolist() == good
    assert cnt.tolist() == [win.count(c) for c in codons]
    assert bs.matchPDict(pd, d1).elementNROWS().tolist() == cnt.tolist()
    # codons(d1): non-overlapping views of width 3
    v = bs.XStringViews(d1, start=[1, 4, 7, 10], width=3)
    win = [d1[i:i + 3] for i in (0, 3, 6, 9)]
    assert bs.whichPDict(pd, v).tolist() == [i + 1 for i, c in enumerate(codons) if c in win]
    assert bs.countPDict(pd, v).tolist() == [win.count(c) for c in codons]
    # mask 4..8 -> views 1-3 and 9-12
    m = bs.MaskedDNAString(d1, 4, 8)
    win = [d1[0:3], d1[8:11], d1[9:12]]
    assert bs.whichPDict(pd, m).tolist() == [i + 1 for i, c in enumerate(codons) if c in win]
    assert bs.countPDict(pd, m).tolist() == [win.count(c) for c in codons]


def test_rd_inexact_examples():
    pd = bs.PDict(["TACCNG", "TAGT", "CGGNT", "AGTAG", "TAGT"], tb_end=3)
    s = "TAGTACCAGTTTCGGG"
    assert bs.countPDict(pd, s).tolist() == [0, 1, 0, 0, 1]
    assert bs.countPDict(pd, s, fixed=False).tolist() == [1, 1, 0, 0, 1]
    assert bs.countPDict(pd, s, max_mismatch=1).tolist() == [1, 1, 0, 1, 1]
    assert bs.countPDict(pd, s, max_mismatch=1, fixed=False).tolist() == [1, 1, 1, 1, 1]
    assert bs.countPDict(pd, s, max_mismatch=2).tolist() == [1, 1, 1, 2, 1]
    assert L(bs.matchPDict(pd, s, max_mismatch=1, fixed=False).endIndex())[2] == [17]
    assert L(bs.matchPDict(pd, s, max_mismatch=2).endIndex())[3] == [6, 12]


def test_rd_iupac_pattern_and_subject():
    pd = bs.PDict(["AAACAAKS", "GGGAAA", "TNCCGGG"], tb_start=3, tb_width=4)
    s = "AAACAATCCCGGGAAACAAGG"
    assert L(bs.matchPDict(pd, s).endIndex()) == [None, [16], None]
    assert L(bs.matchPDict(pd, s, fixed="subject").endIndex()) == [[8, 21], [16], [13]]
    pd = bs.PDict(["ACAC", "TCCG"])
    want = {"ACNCCGT": [[4], [6]], "ACWCCGT": [[4], [6]], "ACRCCGT": [[4], None],
            "ACKCCGT": [None, [6]]}
    for subj, w in want.items():
        assert L(bs.matchPDict(pd, subj, fixed="pattern").endIndex()) == w
        assert L(bs.matchPDict(pd, subj).endIndex()) == [None, None]
    pd = bs.PDict(["TTC", "CTT"])
    assert L(bs.matchPDict(pd, "CYTCACTTC", fixed="pattern").endIndex()) == [[4, 9], [3, 8]]


def test_vcount_vwhich():
    pd = bs.PDict(["acgt", "gt", "cgt", "ac"], tb_end=2)
    d = "acggaccg"
    dss = bs.DNAStringSet([d, d])
    dvv = bs.XStringViews(d, start=[1, 1], width=[8, 8])
    exp = np.array([[0, 0], [0, 0], [0, 0], [2, 2]])
    for subj in (dss, dvv):
        assert (bs.vcountPDict(pd, subj) == exp).all()
        assert [x.tolist() for x in bs.vwhichPDict(pd, subj)] == [[4], [4]]
    exp1 = exp.copy()
    exp1[0, :] = 1
    exp1[2, :] = 2
    for subj in (dss, dvv):
        assert (bs.vcountPDict(pd, subj, max_mismatch=1) == exp1).all()
        assert [x.tolist() for x in bs.vwhichPDict(pd, subj, max_mismatch=1)] == [[1, 3, 4]] * 2
    exp2 = exp1.copy()
    exp2[0, :] = 2
    assert (bs.vcountPDict(pd, dss, max_mismatch=2) == exp2).all()
    with pytest.raises(ValueError, match="please use vmatchPDict"):
        bs.matchPDict(pd, dss)
    with pytest.raises(ValueError, match="please use countPDict"):
        bs.vcountPDict(pd, d)
    with pytest.raises(ValueError, match="please use whichPDict"):
        bs.vwhichPDict(pd, d)
    with pytest.raises(ValueError, match="please use matchPDict"):
        bs.vmatchPDict(pd, d)
    with pytest.raises(NotImplementedError, match="not ready yet"):
        bs.vmatchPDict(pd, dss)


def test_mtb_pdict_small():
    pats = ["ACGTAC", "TTTTTT", "ACGTAC", "GGGCCC"]
    pd = bs.PDict(pats, max_mismatch=1)
    assert isinstance(pd, bs.MTB_PDict) and len(pd.parts()) == 2
    s = "ACGTACTTTATTGGGCCGACGAAC"
    # brute force: Hamming distance <= 1 at every in-bounds start
    def brute(p, k):
        return [i + len(p) for i in range(len(s) - len(p) + 1)
                if sum(a != b for a, b in zip(p, s[i:i + len(p)])) <= k]
    m = bs.matchPDict(pd, s, max_mismatch=1)
    ends = L(m.endIndex())
    for i, p in enumerate(pats):
        b = brute(p, 1)
        inb = [e for e in (ends[i] or []) if len(p) <= e <= len(s)]
        assert inb == b
    assert bs.countPDict(pd, s, max_mismatch=1).tolist() == [len(e or []) for e in ends]
    assert pd.duplicated().tolist() == [False, False, True, False]


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("variable", [False, True])
def test_seeded_dictionaries_of_the_reference_tests(seed, variable):
    """tests/testthat/test-matchPDict.R:176-234 re-expressed (see tests/test_oracle_golden.py):
    six windows cut from a 150-letter target, shuffled, must come back at exactly their places."""
    from tests.test_oracle_golden import seeded_dictionary
    target, pats, starts, widths = seeded_dictionary(seed, variable)
    kw = dict(tb_start=1, tb_width=min(widths)) if variable else {}
    pd = bs.PDict(bs.DNAStringSet(pats), **kw)
    res = bs.matchPDict(pd, bs.DNAString(target))
    assert [None if e is None else e.tolist() for e in res.startIndex()] == [[s] for s in starts]
    assert [None if e is None else e.tolist() for e in res.endIndex()] == \
        [[s + w - 1] for s, w in zip(starts, widths)]
    assert bs.countPDict(pd, bs.DNAString(target)).tolist() == [1] * 6
