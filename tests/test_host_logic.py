"""Host-side logic of the product (no GPU): containers, threebands, Dups, argument
normalisation, MIndex CSR accessors, sharding arithmetic, synthetic inputs."""
import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import matchpdict as mp, pdict as pdm, sharding, synth
from biostrings_b200._lib import NA_INTEGER
from oracle import rlevel as R


def test_encoding_roundtrip():
    s = "ACGTMRWSYKVHDBN-+."
    codes = bs.encode_dna(s)
    assert codes.tolist() == [1, 2, 4, 8, 3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15, 16, 32, 64]
    assert bs.decode_dna(codes) == s
    assert bs.encode_dna("acgt").tolist() == [1, 2, 4, 8]
    with pytest.raises(ValueError, match="not in lookup table"):
        bs.encode_dna("ACGX")
    assert (bs.encode_dna(s) == R.encode(s)).all()


def test_threebands_matches_oracle_restatement():
    g = np.random.Generator(np.random.PCG64(3))
    seqs = ["".join("ACGT"[i] for i in g.integers(0, 4, int(n))) for n in g.integers(8, 20, 50)]
    x, y = bs.DNAStringSet(seqs), R.SeqSet(seqs)
    for kw in ({}, {"start": 2, "end": 6}, {"start": 1, "width": 8}, {"end": -1, "width": 3},
               {"start": 3}, {"end": 5}, {"start": -6, "end": -2}):
        a = pdm.threebands(x, **kw)
        b = R.threebands(y, kw.get("start"), kw.get("end"), kw.get("width"))
        for p, q in zip(a, b):
            assert (p.widths == q.widths).all() and (p.concat == q.concat).all()
    for kw in ({"start": 25}, {"start": 1, "end": 3, "width": 3}, {"width": 3}, {"width": -1, "start": 1}):
        with pytest.raises(ValueError, match="solving row"):
            pdm.threebands(x, **kw)


def test_dups():
    h2l = np.array([NA_INTEGER, NA_INTEGER, 1, NA_INTEGER, 2, 1], dtype=np.int32)
    d, o = pdm.Dups(h2l), R.Dups(h2l)
    assert d.duplicated().tolist() == [False, False, True, False, True, True]
    assert d.togroup().tolist() == [1, 2, 1, 4, 2, 1]
    assert d.togrouplength().tolist() == [3, 2, 3, 1, 2, 3] == o.togrouplength().tolist()
    assert d.low2high(1).tolist() == [3, 6] and d.low2high(4) is None
    assert d.members([1, 4]).tolist() == [1, 3, 4, 6] == o.members([1, 4]).tolist()
    assert d.members([]).tolist() == []


def test_normarg():
    assert mp.normarg_fixed(True) == (True, True)
    assert mp.normarg_fixed(False) == (False, False)
    assert mp.normarg_fixed("pattern") == (True, False)
    assert mp.normarg_fixed("subject") == (False, True)
    assert mp.normarg_fixed(["pattern", "subject"]) == (True, True)
    assert mp.normarg_fixed([True, False]) == (True, False)
    assert mp.normarg_fixed({"subject": True, "pattern": False}) == (False, True)
    for bad in (["pattern", "pattern"], "both", [True, True, True], 3):
        with pytest.raises((ValueError, TypeError)):
            mp.normarg_fixed(bad)
    for f in (True, False, "pattern", "subject"):
        assert mp.normarg_fixed(f) == tuple(R.normarg_fixed(f))
    assert mp.normarg_max_mismatch(2.0) == 2
    with pytest.raises(ValueError, match="non-negative"):
        mp.normarg_max_mismatch(-1)
    with pytest.raises(ValueError, match="must be <= 'max.mismatch'"):
        mp.normarg_min_mismatch(2, 1)
    assert mp.normarg_collapse(False) == 0 and mp.normarg_collapse(2) == 2
    with pytest.raises(ValueError, match="FALSE, 1 or 2"):
        mp.normarg_collapse(3)


def test_mindex_csr_accessors():
    # the example of R/MIndex-class.R:165-176
    width0 = np.array([9, 10, 8, 4, 10], dtype=np.int32)
    dups0 = pdm.Dups(np.array([NA_INTEGER] * 4 + [2], dtype=np.int32))
    m = bs.ByPos_MIndex(width0, [0, 0, 2, 2, 2, 2], [199, 402], names=list("abcde"), dups0=dups0)
    ends = m.endIndex()
    assert ends[0] is None and ends[1].tolist() == [199, 402] and ends[4].tolist() == [199, 402]
    st = m.startIndex()
    assert st[1].tolist() == [190, 393] and st[4].tolist() == [190, 393] and st[2] is None
    assert m.elementNROWS().tolist() == [0, 2, 0, 0, 2]
    s, e, w = m[4]
    assert s.tolist() == [190, 393] and w.tolist() == [10, 10]
    with pytest.raises(IndexError):
        m[5]
    o = R.MIndex(width0, [None, np.array([199, 402]), None, None, None], dups0=R.Dups(dups0.high2low))
    assert [None if x is None else x.tolist() for x in o.startIndex()] == \
        [None if x is None else x.tolist() for x in st]
    # unlist / coverage / extractAllMatches on the CSR (R/MIndex-class.R:56-103)
    us, ue, un = m.unlist(use_names=True)
    assert us.tolist() == [190, 393, 190, 393] and ue.tolist() == [199, 402, 199, 402]
    assert un.tolist() == ["b", "b", "e", "e"]
    cov = m.coverage(410)
    assert cov.size == 410 and cov[188] == 0 and cov[189] == 2 and cov[198] == 2 and cov[199] == 0
    assert int(cov.sum()) == 40
    g = np.random.default_rng(5)
    rows = [np.sort(g.integers(30, 900, int(g.integers(0, 6)))).astype(np.int32) for _ in range(50)]
    offs = np.concatenate([[0], np.cumsum([r.size for r in rows])])
    wid = g.integers(1, 30, 50).astype(np.int32)
    mm = bs.ByPos_MIndex(wid, offs, np.concatenate(rows) if offs[-1] else np.zeros(0, np.int32))
    s2, e2 = mm.unlist()
    want_s = np.concatenate([r - w + 1 for r, w in zip(rows, wid)])
    assert s2.tolist() == want_s.tolist() and e2.tolist() == np.concatenate(rows).tolist()
    brute = np.zeros(1000, dtype=np.int32)
    for r, w in zip(rows, wid):
        for e_ in r:
            brute[max(e_ - w, 0):e_] += 1
    assert mm.coverage(1000).tolist() == brute.tolist()
    subj = bs.DNAString("ACGT" * 250)
    v = bs.extractAllMatches(subj, mm)
    assert len(v) == s2.size and v.start.tolist() == s2.tolist() and v.width.tolist() == (e2 - s2 + 1).tolist()
    with pytest.raises(ValueError, match="MIndex"):
        bs.extractAllMatches(subj, [1, 2])


def test_masked_views():
    m = bs.MaskedDNAString("ATGGATACACAA", 4, 8).as_views()
    assert m.start.tolist() == [1, 9] and m.width.tolist() == [3, 4]
    m = bs.MaskedDNAString("ACGTACGTAC", [1, 3, 9], [2, 5, 20]).as_views()
    assert m.start.tolist() == [6] and m.width.tolist() == [3]
    o = R.Masked("ACGTACGTAC", [1, 3, 9], [2, 5, 20]).as_views()
    assert o.start.tolist() == [6] and o.width.tolist() == [3]
    v = bs.XStringViews("ACGTACGT", [0, 7, 3], [3, 5, 2]).to_set()     # trimmed, not an error
    assert v.widths.tolist() == [2, 2, 2]


def test_shard_bounds_and_chunks():
    for n, world in ((10, 3), (250_000_000, 8), (7, 8), (0, 2)):
        parts = [sharding.shard_bounds(n, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    assert sharding.chunk_range(100, 200, 1000, 24, 13) == (76, 213)
    assert sharding.chunk_range(0, 200, 210, 24, 13) == (0, 210)
    cuts = sharding.split_set_by_letters([5, 5, 5, 5, 100, 5, 5], 2)
    assert cuts[0][0] == 0 and cuts[-1][1] == 7 and cuts[0][1] == cuts[1][0]


def test_synth_is_block_deterministic():
    a = synth.subject_region(0, 3_000_000, 250_000_000)
    b = synth.subject_region(1_500_000, 2_500_000, 250_000_000)
    assert (a[1_500_000:2_500_000] == b).all()
    assert (a[:10_000] == 15).all() and set(np.unique(a[20_000:1_000_000]).tolist()) <= {1, 2, 4, 8, 15}
    gc = np.isin(a[a != 15], [2, 4]).mean()
    assert abs(gc - 0.41) < 0.01
    whole = synth.subject_region(0, 250_000_000 // 50, 250_000_000 // 50)
    assert 0.05 < (whole == 15).mean() < 0.09
    p = synth.dictionary(20_000, 25, 250_000_000, 2)
    q = synth.dictionary(20_000, 25, 250_000_000, 2)
    assert p.shape == (20_000, 25) and (p == q).all() and np.isin(p, [1, 2, 4, 8]).all()
    assert len(np.unique(p, axis=0)) < 20_000           # duplicates are present


def test_seed_shape_table_is_complete():
    """Every row of csrc/bsgpu_shapes.inc must be a complete Hamming filter: for every set
    of k mismatch positions in a window of m letters some listed shift of the shape is
    clean.  (A wrong row would silently lose matches of an MTB_PDict.)"""
    import itertools
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "biostrings_b200", "csrc", "bsgpu_shapes.inc")
    rows = re.findall(r"\{(\d+), (\d+), (\d+), (\d+), (0x[0-9a-f]+)ull, (0x[0-9a-f]+)ull\}", open(path).read())
    assert len(rows) >= 15
    for m, k, q, span, mask, shifts in rows:
        m, k, q, span, mask, shifts = int(m), int(k), int(q), int(span), int(mask, 16), int(shifts, 16)
        assert bin(mask).count("1") == q and mask.bit_length() == span and mask & 1 and span <= 32
        ts = [t for t in range(64) if (shifts >> t) & 1]
        assert ts and max(ts) + span <= m and len(ts) <= 32
        placed = [mask << t for t in ts]
        for mm in itertools.combinations(range(m), k):
            bad = sum(1 << x for x in mm)
            assert any((pl & bad) == 0 for pl in placed), (m, k, mm)


def test_plain_dictionary_argument_checks():
    """selectAlgo / .valid.algos (R/match-utils.R:36-85) as mirrored on the host side, product and
    oracle restatement alike; no device needed."""
    from oracle import rlevel as R
    w = np.array([5, 9, 70], dtype=np.int32)
    for sel in (mp.select_algo, R.select_algo):
        assert sel("auto", w[:2], 0, 0, False, (True, True)) == "boyer-moore"
        assert sel("auto", w[:2], 1, 0, False, (True, True)) == "shift-or"
        assert sel("auto", w, 1, 0, False, (True, True)) == "naive-inexact"          # > 64 letters
        assert sel("auto", w[:2], 1, 1, False, (True, True)) == "naive-inexact"      # min.mismatch
        assert sel("auto", w[:2], 1, 0, False, (True, False)) == "naive-inexact"     # fixed differs
        assert sel("auto", w[:2], 0, 0, False, (False, False)) == "shift-or"
        assert sel("auto", w[:2], 2, 0, True, (True, True)) == "indels"
        assert sel("naive-exact", w[:2], 0, 0, False, (True, True)) == "naive-exact"
        with pytest.raises(ValueError, match="valid algos for your string matching problem"):
            sel("boyer-moore", w[:2], 1, 0, False, (True, True))
        with pytest.raises(ValueError, match="'min.mismatch' must be 0 when 'with.indels' is TRUE"):
            sel("auto", w[:2], 2, 1, True, (True, True))
        with pytest.raises(ValueError, match="empty patterns are not supported"):
            sel("auto", np.array([3, 0]), 0, 0, False, (True, True))
        with pytest.raises(ValueError, match="more than 20000 letters"):
            sel("auto", np.array([3, 20001]), 0, 0, False, (True, True))
    assert mp.normarg_algorithm("shift") == "shift-or" and mp.normarg_algorithm("auto") == "auto"
    with pytest.raises(ValueError):
        mp.normarg_algorithm("naive")                 # ambiguous abbreviation
    with pytest.raises(ValueError, match="TRUE or FALSE"):
        mp.normarg_with_indels(1)
    # the restatement of match_naive_inexact on the examples of R/matchPattern.R:183-185
    enc = R.encode
    assert R.naive_match_ends(enc("TG"), enc("GTGACGTGCAT"), 0, 0, True, True) == [3, 8]
    assert R.naive_match_ends(enc("---"), enc("ACGTGCA"), 3, 0, True, True) == list(range(1, 10))
    assert R.naive_match_ends(enc("---"), enc("A"), 0, 0, True, True) == []


def test_weight_recycling_follows_recycleNumericArg():
    """R/matchPDict.R:541-561: 'weight' goes through recycleNumericArg (numeric, no NA, not empty,
    not longer than its target) when collapsing, and is warned about when it is not."""
    import warnings
    from biostrings_b200 import matchpdict as mp
    assert mp.recycle_numeric_arg(2, "weight", 4).tolist() == [2, 2, 2, 2]
    assert mp.recycle_numeric_arg([1, 2], "weight", 4).tolist() == [1, 2, 1, 2]
    assert mp.recycle_numeric_arg(np.array([0.5]), "weight", 3).dtype == np.float64
    with pytest.raises(ValueError, match="'weight' must be a numeric vector"):
        mp.recycle_numeric_arg("a", "weight", 3)
    with pytest.raises(ValueError, match="'weight' must be a numeric vector"):
        mp.recycle_numeric_arg(True, "weight", 3)
    with pytest.raises(ValueError, match="'weight' has no elements"):
        mp.recycle_numeric_arg([], "weight", 3)
    with pytest.raises(ValueError, match="'weight' is longer"):
        mp.recycle_numeric_arg([1, 2, 3, 4], "weight", 3)
    with pytest.raises(ValueError, match="'weight' contains NAs"):
        mp.recycle_numeric_arg([1.0, float("nan")], "weight", 2)
    with pytest.warns(UserWarning, match="not a multiple"):
        assert mp.recycle_numeric_arg([1, 2], "weight", 3).tolist() == [1, 2, 1]
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        mp.check_weight_without_collapse(1)
    with pytest.warns(UserWarning, match="'weight' is ignored when 'collapse=FALSE'"):
        mp.check_weight_without_collapse(2)
    with pytest.warns(UserWarning, match="'weight' is ignored when 'collapse=FALSE'"):
        mp.check_weight_without_collapse(1.0)
