"""The exact-dictionary engines against the reference on the GPU: scan_byteseed_kernel (16-letter
seeds, S = w - 15 from 2 to 32, NP = 8 / 4 / 2) and, for w >= 25, scan_xseed_kernel (10-letter
core + 18-letter extended seeds, S = w - 9 from 16 to 32, aligned and unaligned tiles, NP = 4 / 2,
4 or 8 filter bits per seed):
every stride / tile regime,
subjects spanning many TMA tiles, sequence sets and views (separators), overlap chunks
(ownership by end position), non-base letters in the subject, and repeats dense enough to
overflow the candidate slices (the confirm-in-place fallback)."""
import warnings

import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import _lib
from oracle import ref, rlevel as R
from tests.helpers import assert_same_ends, random_dna, rng

pytestmark = pytest.mark.gpu
BACKEND = "ref" if ref.available() else "port"


def both(pats, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return bs.PDict(bs.DNAStringSet(pats), **kw), R.PDict(R.SeqSet(pats), backend=BACKEND, **kw)


def plan():
    return _lib.lib().bsgpu_last_plan().decode()


def subject(g, n):
    s = random_dna(g, n)
    for st in g.integers(0, max(1, n - 200), 6):             # runs of N and single IUPAC letters
        s[st:st + int(g.integers(1, 120))] = 15
    s[g.integers(0, n, 40)] = np.array([3, 5, 6, 9, 10, 12, 7, 16], dtype=np.uint8)[g.integers(0, 8, 40)]
    return s


def cut(g, s, w, k):
    pats = []
    while len(pats) < k:
        i = int(g.integers(0, s.size - w))
        p = s[i:i + w].copy()
        if ((p == 1) | (p == 2) | (p == 4) | (p == 8)).all():
            pats.append(p)
    return pats


ENGINES = {"xseed": {}, "xseed_np2": {"BSGPU_XSEED_NP": "2"}, "xseed_np4": {"BSGPU_XSEED_NP": "4"},
           "xseed_kbits8": {"BSGPU_XSEED_KBITS": "8"}, "xseed_split": {"BSGPU_XSEED_SPLIT_CONFIRM": "1"},
           "xseed_unaligned": {"BSGPU_XSEED_UNALIGNED": "1"}, "byteseed": {"BSGPU_NO_XSEED": "1"}}


@pytest.fixture(params=["xseed", "byteseed"])
def engine(request, monkeypatch):
    for k, v in ENGINES[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


def expect_engine(engine, w):
    assert plan().startswith("byteseed_hash"), plan()
    assert ("xseed" in plan()) == (engine.startswith("xseed") and w >= 25), plan()


@pytest.mark.parametrize("variant", ["xseed_np2", "xseed_np4", "xseed_kbits8", "xseed_unaligned", "xseed_split"])
@pytest.mark.parametrize("w", [25, 31])
def test_xseed_tile_variants(w, variant, monkeypatch):
    for k, v in ENGINES[variant].items():
        monkeypatch.setenv(k, v)
    g = rng(7050 + w)
    s = subject(g, 300_000)
    pats = cut(g, s, w, 600) + [random_dna(g, w) for _ in range(300)]
    prod, orac = both(pats)
    got = bs.countPDict(prod, s)
    expect_engine(variant, w)
    assert got.tolist() == R.countPDict(orac, s).tolist()
    assert_same_ends(bs.matchPDict(prod, s).endIndex(), R.matchPDict(orac, s).endIndex())


@pytest.mark.parametrize("w", [17, 18, 25, 26, 27, 28, 38, 39, 40, 41, 42, 46, 47, 48, 64, 65, 100, 130])
def test_exact_dictionary_all_strides(w, engine):
    g = rng(7000 + w)
    s = subject(g, 150_000)
    pats = cut(g, s, w, 400) + [random_dna(g, w) for _ in range(200)]
    pats += [s[:w].copy(), s[s.size - w:].copy()]
    for p in pats[-2:]:
        p[~np.isin(p, [1, 2, 4, 8])] = 1
    prod, orac = both(pats)
    got = bs.countPDict(prod, s)
    expect_engine(engine, w)
    assert got.tolist() == R.countPDict(orac, s).tolist()
    assert bs.whichPDict(prod, s).tolist() == R.whichPDict(orac, s).tolist()
    assert_same_ends(bs.matchPDict(prod, s).endIndex(), R.matchPDict(orac, s).endIndex())


@pytest.mark.parametrize("w", [33, 64, 65, 100, 140])
def test_long_trusted_band_on_the_plain_scan(w, monkeypatch):
    """Trusted Bands wider than the 32-letter hash seed, with a tail and mismatches (tb_hash +
    verify_flanks), and the same whole-pattern dictionary with the byte-seed engine off."""
    g = rng(7150 + w)
    s = subject(g, 40_000)
    pats = cut(g, s, w + 10, 300)
    for p in pats[:150]:
        p[w + int(g.integers(0, 10))] = 2
    pats += [pats[0].copy(), np.concatenate([pats[1][:w], random_dna(g, 10)])]      # duplicate, shared TB
    prod, orac = both(pats, tb_start=1, tb_end=w)
    for mm in (0, 2):
        assert bs.countPDict(prod, s, max_mismatch=mm).tolist() == R.countPDict(orac, s, max_mismatch=mm).tolist()
        assert_same_ends(bs.matchPDict(prod, s, max_mismatch=mm).endIndex(),
                         R.matchPDict(orac, s, max_mismatch=mm).endIndex())
    assert plan().startswith("tb_hash"), plan()
    monkeypatch.setenv("BSGPU_NO_BYTESEED", "1")
    prod, orac = both(cut(g, s, w, 300))
    assert bs.countPDict(prod, s).tolist() == R.countPDict(orac, s).tolist()
    assert plan().startswith("tb_hash"), plan()
    for fixed in ("subject", False):
        if fixed is False:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                assert bs.countPDict(prod, s[:8000], fixed=fixed).tolist() == R.countPDict(orac, s[:8000], fixed=fixed).tolist()


def test_short_widths_keep_the_plain_scan():
    g = rng(7100)
    s = subject(g, 20_000)
    pats = cut(g, s, 16, 200)
    prod, orac = both(pats)
    assert bs.countPDict(prod, s).tolist() == R.countPDict(orac, s).tolist()
    assert plan().startswith("tb_hash"), plan()


def test_sets_views_and_collapse(engine):
    g = rng(7200)
    genome = subject(g, 60_000)
    pats = cut(g, genome, 25, 300)
    reads = [genome[i:i + int(g.integers(0, 180))].copy() for i in g.integers(0, 59_000, 700)]
    prod, orac = both(pats)
    pset, oset = bs.DNAStringSet(reads), R.SeqSet(reads)
    a, b = bs.vcountPDict(prod, pset), R.vcountPDict(orac, oset)
    assert a.shape == b.shape and (a == b).all() and a.sum() > 50
    expect_engine(engine, 25)
    for collapse in (1, 2):
        w = g.integers(1, 5, len(reads) if collapse == 1 else len(pats)).astype(np.int32)
        got = bs.vcountPDict(prod, pset, collapse=collapse, weight=w)
        want = R.vcountPDict(orac, oset, collapse=collapse, weight=w)
        assert np.array_equal(np.asarray(got), np.asarray(want))
    gw, ww = bs.vwhichPDict(prod, pset), R.vwhichPDict(orac, oset)
    assert [v.tolist() for v in gw] == [v.tolist() for v in ww]
    starts = g.integers(1, 59_000, 60).astype(np.int32)
    widths = g.integers(0, 400, 60).astype(np.int32)
    vp = bs.XStringViews(genome, starts, widths)
    vo = R.Views(genome, starts, widths)
    assert bs.countPDict(prod, vp).tolist() == R.countPDict(orac, vo).tolist()
    assert_same_ends(bs.matchPDict(prod, vp).endIndex(), R.matchPDict(orac, vo).endIndex())


@pytest.mark.parametrize("w", [25, 40])
def test_chunks_own_matches_by_end(w, engine):
    from biostrings_b200 import matchpdict as mp
    g = rng(7300 + w)
    s = subject(g, 90_000)
    pats = cut(g, s, w, 600)
    prod, _ = both(pats)
    part = prod.parts()[0]
    with mp.DeviceSubject.from_xstring(s) as whole:
        want = mp.match_parts([part], whole, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS)
        wo, we = mp.match_parts([part], whole, 0, 0, (True, True), _lib.MATCHES_AS_ENDS)
    bounds = [0, 1, 4097, 30_000, 30_010, 65_536, s.size]
    got = np.zeros_like(want)
    rows = [[] for _ in pats]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        a, b = max(0, lo - (w - 1)), hi
        with mp.DeviceSubject.from_chunk(s[a:b], a, s.size, lo, hi) as ch:
            got += mp.match_parts([part], ch, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS)
            o, e = mp.match_parts([part], ch, 0, 0, (True, True), _lib.MATCHES_AS_ENDS)
            for i in range(len(pats)):
                rows[i].extend(e[o[i]:o[i + 1]].tolist())
    assert (got == want).all()
    for i in range(len(pats)):
        assert rows[i] == we[wo[i]:wo[i + 1]].tolist()


def test_dense_repeats_overflow_the_candidate_slices(engine):
    """Every probe of a periodic subject is a seed hit: the candidate slices fill up and the
    scan confirms in place."""
    g = rng(7400)
    unit = random_dna(g, 7)
    s = np.tile(unit, 40_000)                                 # 280 kbp, period 7
    pats = [np.tile(np.roll(unit, r), 4)[:25].copy() for r in range(7)] + cut(g, random_dna(g, 5000), 25, 50)
    prod, orac = both(pats)
    got = bs.countPDict(prod, s)
    expect_engine(engine, 25)
    want = R.countPDict(orac, s)
    assert got.tolist() == want.tolist()
    assert int(got[:7].sum()) == s.size - 24
    assert_same_ends(bs.matchPDict(prod, s).endIndex(), R.matchPDict(orac, s).endIndex())
