"""GPU parity against the CPU oracle on seeded synthetic inputs (through the C ABI).

The oracle is oracle/rlevel.py over the compiled reference C (oracle/_ref) when it was
built, else over our C restatement.  Bit-exact bar: every count, which-index and end.
"""
import numpy as np
import pytest

import biostrings_b200 as bs
from oracle import ref, rlevel as R
from tests.helpers import IUPAC_AMBIG, assert_same_ends, random_dna, rng

pytestmark = pytest.mark.gpu

BACKEND = "ref" if ref.available() else "port"


@pytest.fixture(autouse=True, params=["qgram", "tb_only", "tb_hash"])
def engine(request, monkeypatch):
    """Every case runs three times: with the seed engines allowed (default), with the Trusted-Band
    engines forced (BSGPU_NO_QINDEX=1: band directory where it applies, else the hash scan), and with
    the hash scan alone (BSGPU_NO_TBDIR=1: scan_tb_kernel + tb_hit / verify_flanks_kernel); all must
    equal the reference."""
    for k in ("BSGPU_NO_QINDEX", "BSGPU_NO_BYTESEED", "BSGPU_NO_TBDIR"):
        monkeypatch.delenv(k, raising=False)
    if request.param in ("tb_only", "tb_hash"):
        monkeypatch.setenv("BSGPU_NO_QINDEX", "1")
        monkeypatch.setenv("BSGPU_NO_BYTESEED", "1")        # plain one-probe-per-window scan
    if request.param == "tb_hash":
        monkeypatch.setenv("BSGPU_NO_TBDIR", "1")
    return request.param


def make_dict(g, subject, M, width, frac_planted=0.3, frac_dup=0.05, nsub=0, sub_lo=0):
    """M patterns of `width` (int or (lo, hi)): random, planted windows of the subject
    (optionally with `nsub` substitutions at positions >= sub_lo), and exact duplicates."""
    pats = []
    for i in range(M):
        w = width if isinstance(width, int) else int(g.integers(width[0], width[1] + 1))
        u = g.random()
        if u < frac_dup and pats:
            pats.append(pats[int(g.integers(len(pats)))].copy())
            continue
        if u < frac_dup + frac_planted and len(subject) > w:
            st = int(g.integers(0, len(subject) - w))
            p = subject[st:st + w].copy()
            if not np.isin(p, [1, 2, 4, 8]).all():
                p = random_dna(g, w)
            for _ in range(int(g.integers(0, nsub + 1))):
                j = int(g.integers(sub_lo, w))
                p[j] = [1, 2, 4, 8][int(g.integers(4))]
            pats.append(p)
        else:
            pats.append(random_dna(g, w))
    return pats


def both(pats, **kw):
    """Same dictionary for the product and the oracle."""
    prod = bs.PDict(bs.DNAStringSet(pats), **kw)
    orac = R.PDict(R.SeqSet(pats), backend=BACKEND, **kw)
    assert (prod.width() == orac.width()).all()
    return prod, orac


def check_all(prod, orac, psub, osub, **kw):
    c1 = bs.countPDict(prod, psub, **kw)
    c2 = R.countPDict(orac, osub, **kw)
    assert (c1 == c2).all(), f"counts differ at {np.flatnonzero(c1 != c2)[:10]}"
    assert bs.whichPDict(prod, psub, **kw).tolist() == R.whichPDict(orac, osub, **kw).tolist()
    m1 = bs.matchPDict(prod, psub, **kw)
    m2 = R.matchPDict(orac, osub, **kw)
    assert_same_ends(m1.endIndex(), m2.endIndex())
    assert_same_ends(m1.startIndex(), m2.startIndex())
    assert m1.elementNROWS().tolist() == m2.elementNROWS().tolist()
    return int(c1.sum())


def subject_with_n_runs(g, n):
    s = random_dna(g, n)
    for _ in range(4):
        a = int(g.integers(0, n - 50))
        s[a:a + int(g.integers(1, 40))] = 15
    s[:7] = 15
    s[-5:] = 15
    return s


@pytest.mark.parametrize("algo,width", [("ACtree2", 25), ("ACtree2", 36), ("ACtree2", 5),
                                        ("Twobit", 12), ("ACtree2", 32), ("ACtree2", 33)])
def test_exact_full_tb(algo, width):
    g = rng(100 + width)
    s = subject_with_n_runs(g, 60_000)
    pats = make_dict(g, s, 3000, width)
    prod, orac = both(pats, algorithm=algo)
    assert (prod.duplicated() == orac.duplicated()).all()
    assert (prod.patternFrequency() == orac.patternFrequency()).all()
    total = check_all(prod, orac, s, s)
    assert total > 100


@pytest.mark.parametrize("algo", ["ACtree2", "Twobit"])
@pytest.mark.parametrize("mm", [0, 1, 2, 3])
def test_tb_with_tail(algo, mm):
    g = rng(7 + mm)
    s = subject_with_n_runs(g, 40_000)
    pats = make_dict(g, s, 2000, 36, nsub=3, sub_lo=12)
    prod, orac = both(pats, tb_start=1, tb_end=12, algorithm=algo)
    total = check_all(prod, orac, s, s, max_mismatch=mm)
    assert total > 50
    if mm >= 1:
        check_all(prod, orac, s, s, max_mismatch=mm, min_mismatch=1)


@pytest.mark.parametrize("fixed", [True, False, "pattern", "subject"])
def test_variable_width_iupac_flanks(fixed):
    g = rng(11)
    s = random_dna(g, 30_000)
    amb = g.random(s.size) < 0.01
    s[amb] = IUPAC_AMBIG[g.integers(0, len(IUPAC_AMBIG), int(amb.sum()))]
    s[g.integers(0, s.size, 20)] = 16                        # gap letters
    pats = make_dict(g, s, 1500, (15, 40), frac_dup=0.0)
    for p in pats:                                           # TB = letters 3..12 must be base
        bad = ~np.isin(p[2:12], [1, 2, 4, 8])
        p[2:12][bad] = 1
        if g.random() < 0.3:                                 # IUPAC letters in head / tail
            j = int(g.integers(12, p.size))
            p[j] = IUPAC_AMBIG[int(g.integers(len(IUPAC_AMBIG)))]
        if g.random() < 0.2:
            p[int(g.integers(0, 2))] = 15
    prod, orac = both(pats, tb_start=3, tb_width=10)
    assert prod.dups() is None and orac.dups() is None
    for mm in (0, 2):
        total = check_all(prod, orac, s, s, max_mismatch=mm, fixed=fixed)
    assert total > 20


def test_out_of_limits_matches():
    g = rng(5)
    s = random_dna(g, 400)
    # patterns hanging off both ends of the subject
    pats = [np.concatenate([random_dna(g, 6), s[0:14]]),
            np.concatenate([s[380:400], random_dna(g, 5)]),
            np.concatenate([random_dna(g, 3), s[0:10], random_dna(g, 3)])]
    for tb in ((7, 14), (1, 8)):
        pats3 = [p.copy() for p in pats]
        prod, orac = both(pats3, tb_start=tb[0], tb_end=tb[1])
        for mm in (0, 3, 6):
            check_all(prod, orac, s, s, max_mismatch=mm)


@pytest.mark.parametrize("k,width", [(1, 25), (2, 25), (2, (20, 30)), (3, 36), (1, 18), (1, (21, 40)),
                                     (2, 30), (2, (36, 50)), (2, 64), (3, 50)])
def test_mtb_pdict(k, width):
    g = rng(20 + k)
    s = subject_with_n_runs(g, 30_000)
    pats = make_dict(g, s, 1500, width, nsub=k + 1, frac_dup=0.05 if isinstance(width, int) else 0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prod, orac = both(pats, max_mismatch=k)
    assert len(prod.parts()) == k + 1
    total = check_all(prod, orac, s, s, max_mismatch=k)
    assert total > 100
    check_all(prod, orac, s, s, max_mismatch=k, min_mismatch=1)
    if k == 2:
        check_all(prod, orac, s, s, max_mismatch=1)          # allowed, "not optimal"
    with pytest.raises(ValueError, match="preprocessed to be used"):
        bs.countPDict(prod, s, max_mismatch=k + 1)


@pytest.mark.parametrize("variant", ["compiled_shape", "generic_shape", "bitmap_directory"])
@pytest.mark.parametrize("k,width", [(1, 25), (2, 25), (1, 18), (2, 24), (2, 28), (2, 29)])
def test_mtb_scan_variants(k, width, variant, monkeypatch, engine):
    """The three forms of the Hamming scan (qscan2_kernel with the shape compiled in, the same with
    the shape read from the index, and the older bitmap + directory qscan_kernel) on a dictionary
    with families of near-duplicates: keys with more than 3 entries go through the chain slots."""
    if engine != "qgram":
        pytest.skip("q-gram engine only")
    if variant == "generic_shape":
        monkeypatch.setenv("BSGPU_Q_GENERIC", "1")
    elif variant == "bitmap_directory":
        monkeypatch.setenv("BSGPU_Q_V1", "1")
    g = rng(900 + 10 * k + width)
    s = subject_with_n_runs(g, 30_000)
    pats = make_dict(g, s, 600, width, nsub=k + 1, frac_dup=0.0)
    fam = []
    for p in pats[:60]:                 # 12 relatives each, differing in 1..2 letters
        for _ in range(12):
            q = p.copy()
            for _ in range(int(g.integers(1, 3))):
                q[int(g.integers(0, width))] = [1, 2, 4, 8][int(g.integers(4))]
            fam.append(q)
    pats = pats + fam
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prod, orac = both(pats, max_mismatch=k)
    total = check_all(prod, orac, s, s, max_mismatch=k)
    assert total > 100
    check_all(prod, orac, s, s, max_mismatch=k, min_mismatch=1)
    from biostrings_b200 import _lib
    bs.countPDict(prod, s, max_mismatch=k)
    plan = _lib.lib().bsgpu_last_plan().decode()
    assert ("rank_directory" in plan) == (variant != "bitmap_directory"), plan


def test_views_subject():
    g = rng(31)
    s = subject_with_n_runs(g, 20_000)
    pats = make_dict(g, s, 1000, 30, nsub=2, sub_lo=10)
    prod, orac = both(pats, tb_start=1, tb_end=10)
    # overlapping, unordered, empty and tiny views
    starts = np.array([5000, 1, 4000, 19_990, 300, 7000, 7000], dtype=np.int32)
    widths = np.array([3000, 4500, 2500, 11, 0, 5, 1000], dtype=np.int32)
    pv = bs.XStringViews(s, starts, widths)
    ov = R.Views(s, starts, widths)
    for mm in (0, 2):
        check_all(prod, orac, pv, ov, max_mismatch=mm)
    with pytest.raises(bs.BsgpuError, match="out of limits"):
        bs.countPDict(prod, bs.XStringViews(s, [19_000], [2000]))
    # masked subject = views on the gaps
    pm = bs.MaskedDNAString(s, [100, 9000], [2000, 12_000])
    om = R.Masked(s, [100, 9000], [2000, 12_000])
    check_all(prod, orac, pm, om, max_mismatch=1)
    # MTB + overlapping views: union dedupes across views
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prod2, orac2 = both(make_dict(g, s, 800, 24, nsub=2), max_mismatch=1)
    check_all(prod2, orac2, pv, ov, max_mismatch=1)


def test_vcount_vwhich_sets():
    g = rng(41)
    genome = random_dna(g, 50_000)
    reads = []
    for _ in range(400):
        n = int(g.integers(0, 200))
        st = int(g.integers(0, genome.size - 200))
        reads.append(genome[st:st + n].copy())
    reads[3] = reads[3][:0]
    pats = make_dict(g, genome, 600, 25, nsub=1, frac_dup=0.1)
    prod, orac = both(pats)
    pset, oset = bs.DNAStringSet(reads), R.SeqSet(reads)
    a, b = bs.vcountPDict(prod, pset), R.vcountPDict(orac, oset)
    assert a.shape == b.shape and (a == b).all() and a.sum() > 50
    w1 = g.integers(1, 5, len(reads)).astype(np.int32)
    w2 = g.integers(1, 5, len(pats)).astype(np.int32)
    assert (bs.vcountPDict(prod, pset, collapse=1, weight=w1) ==
            R.vcountPDict(orac, oset, collapse=1, weight=w1)).all()
    assert (bs.vcountPDict(prod, pset, collapse=2, weight=w2) ==
            R.vcountPDict(orac, oset, collapse=2, weight=w2)).all()
    assert (bs.vcountPDict(prod, pset, collapse=1) == a.sum(axis=1)).all()
    d1 = g.random(len(reads)) * 3
    d2 = g.random(len(pats)) * 3
    x, y = bs.vcountPDict(prod, pset, collapse=1, weight=d1), R.vcountPDict(orac, oset, collapse=1, weight=d1)
    assert x.dtype == np.float64 and (x == y).all()
    x, y = bs.vcountPDict(prod, pset, collapse=2, weight=d2), R.vcountPDict(orac, oset, collapse=2, weight=d2)
    assert (x == y).all()
    wa, wb = bs.vwhichPDict(prod, pset), R.vwhichPDict(orac, oset)
    assert [v.tolist() for v in wa] == [v.tolist() for v in wb]
    # with a tail and mismatches
    prod, orac = both(pats, tb_start=1, tb_end=12)
    a, b = bs.vcountPDict(prod, pset, max_mismatch=2), R.vcountPDict(orac, oset, max_mismatch=2)
    assert (a == b).all()
    # MTB: only vwhich is supported (R/matchPDict.R:481-482)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prod, orac = both(pats, max_mismatch=1)
    wa, wb = bs.vwhichPDict(prod, pset, max_mismatch=1), R.vwhichPDict(orac, oset, max_mismatch=1)
    assert [v.tolist() for v in wa] == [v.tolist() for v in wb]
    with pytest.raises(ValueError, match="not supported yet"):
        bs.vcountPDict(prod, pset, max_mismatch=1)


def test_iupac_subject_full_tb():
    g = rng(51)
    s = random_dna(g, 30_000)
    amb = g.random(s.size) < 0.004
    s[amb] = IUPAC_AMBIG[g.integers(0, len(IUPAC_AMBIG), int(amb.sum()))]
    s[1000:1003] = 15
    s[2000] = 16
    clean = s.copy()
    clean[~np.isin(clean, [1, 2, 4, 8])] = 1
    pats = make_dict(g, clean, 2000, 15, frac_planted=0.6)
    prod, orac = both(pats)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n0 = check_all(prod, orac, s, s, fixed=True)
        n1 = check_all(prod, orac, s, s, fixed=False)
        check_all(prod, orac, s, s, fixed="pattern")
    assert n1 > n0 > 0
    # dense ambiguity (a run of N): every Trusted-Band word matches every window, as in the
    # reference, whose node set then holds the whole trie level
    dense = np.concatenate([s[:300], np.full(60, 15, dtype=np.uint8), s[300:600]])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n2 = check_all(prod, orac, dense, dense, fixed=False)
    assert n2 > 40 * 1000
    # ... up to a work budget, beyond which the reference's own message is raised
    # (its cap is 5M live trie nodes, src/match_pdict_ACtree2.c:1024,1051-1054)
    import os
    os.environ["BSGPU_IUPAC_BRUTE_BUDGET"] = "1000"
    try:
        with pytest.raises(bs.BsgpuError, match="too many IUPAC ambiguity letters"):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                bs.countPDict(prod, dense, fixed=False)
    finally:
        del os.environ["BSGPU_IUPAC_BRUTE_BUDGET"]


def test_chunked_subject_equals_whole():
    """Sharding a long subject into overlap chunks gives the same answer."""
    from biostrings_b200 import matchpdict as mp
    from biostrings_b200._lib import MATCHES_AS_COUNTS, MATCHES_AS_ENDS
    g = rng(61)
    s = subject_with_n_runs(g, 50_000)
    pats = make_dict(g, s, 1500, 30, nsub=2, sub_lo=12)
    prod, _ = both(pats, tb_start=3, tb_end=12)
    part = prod.parts()[0]
    halo_l, halo_r = 12 - 1 + 2, 18
    with mp.DeviceSubject.from_xstring(s) as whole:
        want_c = mp.match_parts([part], whole, 2, 0, (True, True), MATCHES_AS_COUNTS)
        want_o, want_e = mp.match_parts([part], whole, 2, 0, (True, True), MATCHES_AS_ENDS)
    G = 4
    bounds = np.linspace(0, s.size, G + 1).astype(int)
    got_c = np.zeros_like(want_c)
    rows = [[] for _ in pats]
    for r in range(G):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        a, b = max(0, lo - halo_l), min(s.size, hi + halo_r)
        with mp.DeviceSubject.from_chunk(s[a:b], a, s.size, lo, hi) as ch:
            got_c += mp.match_parts([part], ch, 2, 0, (True, True), MATCHES_AS_COUNTS)
            o, e = mp.match_parts([part], ch, 2, 0, (True, True), MATCHES_AS_ENDS)
            for i in range(len(pats)):
                rows[i].extend(e[o[i]:o[i + 1]].tolist())
    assert (got_c == want_c).all()
    for i in range(len(pats)):
        assert rows[i] == want_e[want_o[i]:want_o[i + 1]].tolist()
    # a chunk without enough halo is refused
    with mp.DeviceSubject.from_chunk(s[1000:2000], 1000, s.size, 1000, 2000) as ch:
        with pytest.raises(bs.BsgpuError, match="insufficient halo"):
            mp.match_parts([part], ch, 2, 0, (True, True), MATCHES_AS_COUNTS)
