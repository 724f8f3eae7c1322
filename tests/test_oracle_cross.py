"""Our C restatement (oracle/pdict_oracle.c) against the compiled reference (oracle/_ref)
on seeded inputs, including the shapes where the reference switches to its BitMatrix flank
code (src/match_pdict_utils.c:835-850: >= 15 TB ends and >= 25 patterns per TB group)."""
import numpy as np
import pytest

from oracle import ref, rlevel as R
from tests.helpers import IUPAC_AMBIG, assert_same_ends, random_dna, rng

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built here")


def both(pats, **kw):
    return (R.PDict(R.SeqSet(pats), backend="ref", **kw),
            R.PDict(R.SeqSet(pats), backend="port", **kw))


def cmp_all(a, b, subj, **kw):
    assert R.countPDict(a, subj, **kw).tolist() == R.countPDict(b, subj, **kw).tolist()
    assert R.whichPDict(a, subj, **kw).tolist() == R.whichPDict(b, subj, **kw).tolist()
    assert_same_ends(R.matchPDict(a, subj, **kw).endIndex(), R.matchPDict(b, subj, **kw).endIndex())


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_tb_flanks(seed):
    g = rng(seed)
    s = random_dna(g, 20_000)
    s[500:520] = 15
    pats = [s[i:i + 30].copy() for i in g.integers(0, 19_000, 300)] + \
           [random_dna(g, 30) for _ in range(300)]
    for p in pats:
        p[p == 15] = 1
    for p in pats[:200]:
        p[int(g.integers(10, 30))] = 4
    a, b = both(pats, tb_start=1, tb_end=10)
    for mm in (0, 1, 3):
        cmp_all(a, b, s, max_mismatch=mm)
    cmp_all(a, b, s, max_mismatch=2, min_mismatch=2)


def test_bitmatrix_regime():
    """TB width 4, ~40 patterns per TB group, ~100 TB ends per group, short base-only
    flanks, max.mismatch 2: the reference's BitMatrix path is active."""
    g = rng(11)
    s = random_dna(g, 30_000)
    s[1000:1004] = 15
    pats = [s[i:i + 20].copy() for i in g.integers(1100, 29_900, 10_000)]
    for p in pats:
        for _ in range(int(g.integers(0, 4))):
            j = int(g.integers(0, 16))
            p[j if j < 8 else j + 4] = [1, 2, 4, 8][int(g.integers(4))]
    a, b = both(pats, tb_start=9, tb_width=4)
    for mm in (1, 2):
        ca = R.countPDict(a, s, max_mismatch=mm)
        cb = R.countPDict(b, s, max_mismatch=mm)
        assert ca.tolist() == cb.tolist() and ca.sum() > 1000
    cmp_all(a, b, s[:6000], max_mismatch=2, min_mismatch=1)


def test_iupac_subject_and_patterns():
    g = rng(21)
    s = random_dna(g, 8000)
    amb = g.random(s.size) < 0.02
    s[amb] = IUPAC_AMBIG[g.integers(0, len(IUPAC_AMBIG), int(amb.sum()))]
    s[300] = 16
    pats = [random_dna(g, 12) for _ in range(3000)]
    for p in pats[:500]:
        p[int(g.integers(8, 12))] = IUPAC_AMBIG[int(g.integers(len(IUPAC_AMBIG)))]
    a, b = both(pats, tb_end=8)
    for fixed in (True, False, "pattern", "subject"):
        cmp_all(a, b, s, max_mismatch=1, fixed=fixed)


def test_sets_and_views():
    g = rng(31)
    s = random_dna(g, 5000)
    pats = [s[i:i + 15].copy() for i in g.integers(0, 4900, 200)]
    a, b = both(pats, tb_end=8)
    reads = [s[i:i + int(g.integers(0, 60))] for i in g.integers(0, 4900, 50)]
    sa = R.SeqSet(reads)
    for coll, w in ((False, 1), (1, g.integers(1, 4, 50)), (2, g.integers(1, 4, 200)),
                    (1, g.random(50)), (2, g.random(200))):
        x = R.vcountPDict(a, sa, max_mismatch=1, collapse=coll, weight=w)
        y = R.vcountPDict(b, sa, max_mismatch=1, collapse=coll, weight=w)
        assert np.array_equal(x, y)
    assert [v.tolist() for v in R.vwhichPDict(a, sa)] == [v.tolist() for v in R.vwhichPDict(b, sa)]
    v = R.Views(s, [1, 2000, 100], [3000, 500, 0])
    cmp_all(a, b, v, max_mismatch=1)
    with pytest.raises(Exception, match="out of limits"):
        R.countPDict(a, R.Views(s, [4990], [100]))
    with pytest.raises(Exception, match="out of limits"):
        R.countPDict(b, R.Views(s, [4990], [100]))


def test_mtb():
    g = rng(41)
    s = random_dna(g, 10_000)
    pats = [s[i:i + 24].copy() for i in g.integers(0, 9900, 400)]
    for p in pats:
        for _ in range(int(g.integers(0, 3))):
            p[int(g.integers(0, 24))] = 2
    pats[5] = pats[2].copy()
    a, b = both(pats, max_mismatch=2)
    cmp_all(a, b, s, max_mismatch=2)
    cmp_all(a, b, s, max_mismatch=1)
    assert a.duplicated().tolist() == b.duplicated().tolist()


YEAST = "/root/reference/data/yeastSEQCHR1.rda"


@pytest.mark.skipif(not __import__("os").path.exists(YEAST), reason="reference data not mounted")
def test_config1_yeast_chr1_twobit():
    """BASELINE config 1: countPDict + matchPDict of 10k random 25-mers (+1k cut from the
    subject) against yeastSEQCHR1 with a Twobit PDict (tb.start=1, tb.end=12), reference C vs
    our restatement.  The sequence is read from the gzip'd RDX2 file (string of 230,208 letters
    at offset 63 of the stream)."""
    import gzip
    raw = gzip.open(YEAST).read()
    seq = raw[63:63 + 230_208].decode("ascii")
    assert set(seq) <= set("ACGT") and seq.count("A") == 69_830 and seq.count("C") == 44_643
    s = R.encode(seq)
    g = rng(1)
    pats = [random_dna(g, 25) for _ in range(10_000)] + \
           [s[i:i + 25].copy() for i in g.integers(0, len(s) - 25, 1000)]
    a, b = both(pats, tb_start=1, tb_end=12, algorithm="Twobit")
    ca, cb = R.countPDict(a, s), R.countPDict(b, s)
    assert ca.tolist() == cb.tolist() and ca[10_000:].min() >= 1
    assert_same_ends(R.matchPDict(a, s).endIndex(), R.matchPDict(b, s).endIndex())
    cmp_all(a, b, s, max_mismatch=2)
