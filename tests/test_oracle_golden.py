"""The CPU oracles against every golden vector the reference holds for this path
(SURVEY.md section 8c): tests/testthat/test-matchPDict.R, tests/testthat/test-PDict.R and
the worked examples of man/matchPDict-inexact.Rd.  Runs on the compiled reference
(oracle/_ref, when built) and on our C restatement; both must reproduce them."""
import itertools

import numpy as np
import pytest

from oracle import rlevel as R
from tests.helpers import ends_as_lists as L


@pytest.fixture(params=["port", "ref"])
def be(request, backends):
    if request.param not in backends:
        pytest.skip("oracle/_ref not built here")
    return request.param


def test_inexact_counts(be):                     # test-matchPDict.R:63-76
    pd = R.PDict(["acgt", "gt", "cgt", "ac"], tb_end=2, backend=be)
    d = "acggaccg"
    assert R.countPDict(pd, d, max_mismatch=0).tolist() == [0, 0, 0, 2]
    assert R.countPDict(pd, d, max_mismatch=1).tolist() == [1, 0, 2, 2]
    assert R.countPDict(pd, d, max_mismatch=2).tolist() == [2, 0, 2, 2]
    assert L(R.matchPDict(pd, d, max_mismatch=1).endIndex()) == [[4], None, [4, 9], [2, 6]]
    assert L(R.matchPDict(pd, d, max_mismatch=2).endIndex()) == [[4, 8], None, [4, 9], [2, 6]]


def test_codons_xstring_views_masked(be):        # test-matchPDict.R:11-47
    codons = ["".join(c) for c in itertools.product("ACGT", repeat=3)]
    pd = R.PDict(codons, backend=be)
    d1 = "ATGGATACACAA"
    win = [d1[i:i + 3] for i in range(len(d1) - 2)]
    assert R.whichPDict(pd, d1).tolist() == [i + 1 for i, c in enumerate(codons) if c in win]
    assert R.countPDict(pd, d1).tolist() == [win.count(c) for c in codons]
    v = R.Views(d1, [1, 4, 7, 10], [3, 3, 3, 3])
    win = [d1[i:i + 3] for i in (0, 3, 6, 9)]
    assert R.countPDict(pd, v).tolist() == [win.count(c) for c in codons]
    m = R.Masked(d1, 4, 8)
    win = [d1[0:3], d1[8:11], d1[9:12]]
    assert R.whichPDict(pd, m).tolist() == [i + 1 for i, c in enumerate(codons) if c in win]


def test_vcount_vwhich(be):                      # test-matchPDict.R:78-123
    pd = R.PDict(["acgt", "gt", "cgt", "ac"], tb_end=2, backend=be)
    d = "acggaccg"
    dss = R.SeqSet([d, d])
    dvv = R.Views(d, [1, 1], [8, 8])
    exp = np.array([[0, 0], [0, 0], [0, 0], [2, 2]])
    for s in (dss, dvv):
        assert (R.vcountPDict(pd, s) == exp).all()
        assert [x.tolist() for x in R.vwhichPDict(pd, s)] == [[4], [4]]
    exp[0, :] = 1
    exp[2, :] = 2
    assert (R.vcountPDict(pd, dss, max_mismatch=1) == exp).all()
    assert [x.tolist() for x in R.vwhichPDict(pd, dvv, max_mismatch=1)] == [[1, 3, 4]] * 2
    exp[0, :] = 2
    assert (R.vcountPDict(pd, dvv, max_mismatch=2) == exp).all()
    with pytest.raises(ValueError, match="please use vmatchPDict"):
        R.matchPDict(pd, dss)
    with pytest.raises(ValueError, match="please use countPDict"):
        R.vcountPDict(pd, d)
    with pytest.raises(NotImplementedError, match="not ready yet"):
        R.vmatchPDict(pd, dss)


def test_mindex(be):                             # test-matchPDict.R:125-145
    pd = R.PDict(["AA", "AT", "TA", "TT"], backend=be)
    m = R.matchPDict(pd, "ATTATTAATTAA")
    assert L(m.startIndex()) == [[7, 11], [1, 4, 8], [3, 6, 10], [2, 5, 9]]
    assert m.elementNROWS().tolist() == [2, 3, 3, 3]


def test_rd_examples(be):                        # man/matchPDict-inexact.Rd:123-220
    pd = R.PDict(["TACCNG", "TAGT", "CGGNT", "AGTAG", "TAGT"], tb_end=3, backend=be)
    s = "TAGTACCAGTTTCGGG"
    assert R.countPDict(pd, s).tolist() == [0, 1, 0, 0, 1]
    assert R.countPDict(pd, s, fixed=False).tolist() == [1, 1, 0, 0, 1]
    assert R.countPDict(pd, s, max_mismatch=1).tolist() == [1, 1, 0, 1, 1]
    assert R.countPDict(pd, s, max_mismatch=1, fixed=False).tolist() == [1, 1, 1, 1, 1]
    assert R.countPDict(pd, s, max_mismatch=2).tolist() == [1, 1, 1, 2, 1]
    assert L(R.matchPDict(pd, s, max_mismatch=1, fixed=False).endIndex())[2] == [17]
    pd = R.PDict(["AAACAAKS", "GGGAAA", "TNCCGGG"], tb_start=3, tb_width=4, backend=be)
    s = "AAACAATCCCGGGAAACAAGG"
    assert L(R.matchPDict(pd, s).endIndex()) == [None, [16], None]
    assert L(R.matchPDict(pd, s, fixed="subject").endIndex()) == [[8, 21], [16], [13]]
    pd = R.PDict(["ACAC", "TCCG"], backend=be)
    for subj, w in {"ACNCCGT": [[4], [6]], "ACWCCGT": [[4], [6]], "ACRCCGT": [[4], None],
                    "ACKCCGT": [None, [6]]}.items():
        assert L(R.matchPDict(pd, subj, fixed="pattern").endIndex()) == w
        assert L(R.matchPDict(pd, subj).endIndex()) == [None, None]
    pd = R.PDict(["TTC", "CTT"], backend=be)
    assert L(R.matchPDict(pd, "CYTCACTTC", fixed="pattern").endIndex()) == [[4, 9], [3, 8]]


@pytest.mark.parametrize("algo", ["ACtree2", "Twobit"])
def test_pdict_constructor(be, algo):            # test-PDict.R:13-96
    pd = R.PDict(["A", "A"], algorithm=algo, backend=be)
    assert pd.duplicated().tolist() == [False, True]
    assert pd.patternFrequency().tolist() == [2, 2]
    pd = R.PDict(["ACGT", "ANNT"], tb_start=2, tb_width=2, algorithm=algo, backend=be) \
        if False else R.PDict(["ACGN", "NCGT"], tb_start=2, tb_width=2, algorithm=algo, backend=be)
    tp = pd.threeparts_list[0]
    assert tp.head.widths.tolist() == [1, 1] and tp.tb.widths.tolist() == [2, 2]
    assert tp.tail.widths.tolist() == [1, 1]
    pd = R.PDict(["ACGTAC", "TTGACA", "CAGTTT"], max_mismatch=1, algorithm=algo, backend=be)
    assert pd.kind == "MTB_PDict" and len(pd.threeparts_list) == 2
    assert pd.threeparts_list[0].tb.widths.tolist() == [3, 3, 3]
    assert pd.threeparts_list[0].tail.widths.tolist() == [3, 3, 3]
    assert pd.threeparts_list[1].head.widths.tolist() == [3, 3, 3]


def test_pdict_errors(be):                       # test-PDict.R:98-142
    with pytest.raises(Exception, match="non base DNA letter found in Trusted Band for pattern 2"):
        R.PDict(["ACGT", "ANGT"], backend=be)
    with pytest.raises(Exception, match="non-base DNA letter found in Trusted Band for pattern 2"):
        R.PDict(["ACGT", "ANGT"], algorithm="Twobit", backend=be)
    with pytest.raises(Exception, match="element 2 in Trusted Band has a different length"):
        R.PDict(["ACGT", "ACG"], backend=be)
    with pytest.raises(Exception, match="all the trusted regions must have the same length"):
        R.PDict(["ACGT", "ACG"], algorithm="Twobit", backend=be)
    with pytest.raises(Exception, match="must be <= 14"):
        R.PDict(["ACGTACGTACGTACGT"], algorithm="Twobit", backend=be)
    with pytest.raises(ValueError, match="is >= 6"):
        R.PDict(["ACGTA"], max_mismatch=1, backend=be)
    with pytest.raises(ValueError, match="must be <= 1 given"):
        R.PDict(["ACGTAC"], max_mismatch=2, backend=be)
    with pytest.raises(ValueError, match="must be NAs"):
        R.PDict(["ACGTAC"], max_mismatch=1, tb_end=3, backend=be)
    with pytest.raises(ValueError, match="duplicated names"):
        R.PDict(R.SeqSet(["ACGT", "ACGA"], names=["a", "a"]), backend=be)
    with pytest.raises(ValueError, match="invalid names"):
        R.PDict(R.SeqSet(["ACGT", "ACGA"], names=["a", ""]), backend=be)
    with pytest.raises(ValueError, match="at least one pattern"):
        R.PDict([], backend=be)
    pd = R.PDict(["ACGT", "CCGT"], backend=be)
    with pytest.raises(ValueError, match="must be 0 when there is no head and no tail"):
        R.countPDict(pd, "ACGTACGT", max_mismatch=1)
    pd = R.PDict(["ACGT", "CCGT"], algorithm="Twobit", tb_end=3, backend=be)
    with pytest.raises(Exception, match="cannot treat IUPAC extended letters in the subject"):
        R.countPDict(pd, "ACGTACGT", fixed="pattern")


def test_known_answer_oligo_frequency(be):
    """countPDict(PDict(all k-mers), x) == oligonucleotideFrequency(x, k)
    (R/letterFrequency.R:541-553)."""
    g = np.random.Generator(np.random.PCG64(9))
    x = np.array([1, 2, 4, 8], dtype=np.uint8)[g.integers(0, 4, 5000)]
    x[100:110] = 15
    k = 4
    kmers = ["".join(c) for c in itertools.product("ACGT", repeat=k)]
    pd = R.PDict(kmers, backend=be)
    want = np.zeros(4 ** k, dtype=np.int64)
    code = {1: 0, 2: 1, 4: 2, 8: 3}
    for i in range(len(x) - k + 1):
        w = x[i:i + k]
        if (w != 15).all():
            idx = 0
            for b in w.tolist():
                idx = idx * 4 + code[b]
            want[idx] += 1
    assert R.countPDict(pd, x).tolist() == want.tolist()


def test_plain_dictionary_known_answers(be):
    """Non-preprocessed dictionaries (one matchPattern per pattern): the golden vectors of
    tests/testthat/test-matchPattern.R:55-125 and test-matchPDict.R:15-17, through
    matchPDict(DNAStringSet, ...)."""
    dna = "ATGCATGCATGC"

    def se(pat, subject, **kw):
        m = R.matchPDict(R.SeqSet([pat]), subject, backend=be, **kw)
        s, e = m.startIndex()[0], m.endIndex()[0]
        return ([] if s is None else s.tolist(), [] if e is None else e.tolist())

    algos = ["naive-exact", "naive-inexact", "boyer-moore", "shift-or"] if be == "ref" else ["auto"]
    for algo in algos:
        assert se("ATG", dna, algorithm=algo) == ([1, 5, 9], [3, 7, 11])
        assert se("ATG", R.Masked(dna, [2, 9], [4, 11]), algorithm=algo) == ([5], [7])
        assert se("ATG", R.Views(dna, [2, 9], [3, 3]), algorithm=algo) == ([9], [11])
    assert se("ATC", dna, max_mismatch=2) == ([1, 2, 5, 6, 9, 10], [3, 4, 7, 8, 11, 12])
    assert se("ATC", dna, max_mismatch=2, min_mismatch=2) == ([2, 6, 10], [4, 8, 12])
    dna2 = "ACGTNMRWSYKVHDB"
    assert se("GTN", dna2, fixed=True) == ([3], [5])
    assert se("GTN", dna2, fixed=False) == ([3, 7, 9, 12], [5, 9, 11, 14])
    assert se("---", "ACGTGCA", max_mismatch=3) == (list(range(-1, 8)), list(range(1, 10)))
    assert se("---", "A") == ([], [])
    assert se("ATA", "ATGGATACACAA") == ([5], [7])                  # test-matchPDict.R:15-17
    with pytest.raises(ValueError, match="valid algos for your string matching problem"):
        se("ATG", dna, algorithm="indels")
    with pytest.raises(ValueError, match="empty patterns are not supported"):
        se("", dna)
    # vcountPattern("TG", Views) = c(1, 1)   (test-matchPattern.R:148)
    v = R.Views(dna, [2, 9], [3, 3])
    assert R.vcountPDict(R.SeqSet(["TG"]), v, backend=be).tolist() == [[1, 1]]


# ---- the reference's seeded tests, re-expressed (test-matchPDict.R:176-234, test-PDict.R:174-221) ----
# They draw a 150-letter target with R's RNG (set.seed(1) + sample), cut L = 6 windows out of it
# with successiveIRanges(), shuffle them and expect every window back at exactly its place.  R's
# RNG cannot be replayed without R, so the same construction runs on our own seeded target: the
# expectations are properties of the construction, not of the particular random draw.

def successive_ranges(widths, gaps):
    """IRanges::successiveIRanges(width, gapwidth): 1-based starts of back-to-back ranges."""
    starts, pos = [], 1
    for i, w in enumerate(widths):
        starts.append(pos)
        pos += w + (gaps[i] if i < len(gaps) else 0)
    return starts


def seeded_dictionary(seed, variable):
    from tests.helpers import random_dna, rng
    g = rng(seed)
    target = random_dna(g, 150, gc=0.5)
    W, n = 20, 6
    n_cut = g.integers(0, 6, n) if variable else np.zeros(n, dtype=np.int64)
    widths = (W - n_cut).tolist()
    starts = successive_ranges(widths, (1 + n_cut[:-1]).tolist())
    order = g.permutation(n)
    pats = [target[starts[i] - 1:starts[i] - 1 + widths[i]].copy() for i in order]
    return target, pats, [starts[i] for i in order], [widths[i] for i in order]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_constant_width_dictionary(be, seed):    # test-matchPDict.R:176-201, test-PDict.R:174-192
    target, pats, starts, widths = seeded_dictionary(seed, variable=False)
    pd = R.PDict(R.SeqSet(pats), backend=be)
    assert len(pd) == 6 and pd.width().tolist() == [20] * 6
    tp = pd.threeparts_list[0]
    assert tp.head_or_none() is None and tp.tail_or_none() is None and tp.tb.widths.tolist() == [20] * 6
    res = R.matchPDict(pd, target)
    assert L(res.startIndex()) == [[s] for s in starts]
    assert L(res.endIndex()) == [[s + 19] for s in starts]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_variable_width_dictionary(be, seed):    # test-matchPDict.R:203-234, test-PDict.R:194-221
    target, pats, starts, widths = seeded_dictionary(seed, variable=True)
    mw = min(widths)
    pd = R.PDict(R.SeqSet(pats), tb_start=1, tb_width=mw, backend=be)
    assert len(pd) == 6 and pd.width().tolist() == widths
    tp = pd.threeparts_list[0]
    assert tp.head_or_none() is None and tp.tb.widths.tolist() == [mw] * 6
    assert tp.tail.widths.tolist() == [w - mw for w in widths]
    for i, p in enumerate(pats):                # tail(pdict) == substring(x, min width + 1)
        a = int(tp.tail.offsets[i]) if hasattr(tp.tail, "offsets") else sum(tp.tail.widths[:i])
        assert tp.tail.concat[a:a + widths[i] - mw].tolist() == p[mw:].tolist()
    res = R.matchPDict(pd, target)
    assert L(res.startIndex()) == [[s] for s in starts]
    assert L(res.endIndex()) == [[s + w - 1] for s, w in zip(starts, widths)]
