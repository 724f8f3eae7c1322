"""BASELINE.json's full sizes (1M 25-mers against 250 Mbp, the C2 / C3' inputs of bench.py) on
one GPU, checked through size-independent properties, because the reference needs minutes per
Mbp-million-patterns there:

  * every ends row is ascending, has `count` entries and every reported end spells the pattern
    (max.mismatch 0) or is within max.mismatch substitutions of it (MTB, max.mismatch 2);
  * sum(counts) == number of reported ends; which == patterns with a non-zero count;
  * duplicated patterns get their representative's row;
  * the whole subject == the sum over 5 overlap chunks (the sharding of SURVEY 8e);
  * max.mismatch 2 finds everything max.mismatch 0 finds, and min.mismatch = 1 removes
    exactly those;
  * a 5 Mbp window of it, counted by the compiled reference, agrees bit for bit.
"""
import warnings

import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp, sharding, synth
from oracle import ref, rlevel as R

pytestmark = pytest.mark.gpu
N, M, W = 250_000_000, 1_000_000, 25


@pytest.fixture(scope="module")
def world():
    subj = synth.subject_region(0, N, N)
    pats = synth.dictionary(M, W, N, 1)
    dset = bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(M, W, np.int32))
    return subj, pats, dset


def mismatches(subj, pats, rows, ends):
    """# of letters by which pattern rows[i] differs from the subject window ending at ends[i]"""
    idx = (ends.astype(np.int64) - W)[:, None] + np.arange(W)[None, :]
    return (subj[idx] != pats[rows]).sum(axis=1)


def test_exact_full_size(world):
    subj, pats, dset = world
    pd = bs.PDict(dset)
    with mp.DeviceSubject.from_xstring(subj) as ds:
        counts = bs.countPDict(pd, ds)
        assert _lib.lib().bsgpu_last_plan().decode().startswith("byteseed_hash")
        which = bs.whichPDict(pd, ds)
        mi = bs.matchPDict(pd, ds)
    n = mi.elementNROWS()
    assert (n == counts).all() and counts.sum() > 90_000
    assert which.tolist() == (np.flatnonzero(counts) + 1).tolist()
    s, e = mi.unlist()
    rows = np.repeat(np.arange(M), n)
    assert (e - s + 1 == W).all() and (mismatches(subj, pats, rows, e) == 0).all()
    first = np.concatenate([[True], rows[1:] != rows[:-1]])
    assert (np.diff(e.astype(np.int64))[~first[1:]] > 0).all()         # ascending within a row
    # duplicates: same counts as their representative
    h2l = pd.dups().high2low
    dup = np.flatnonzero(h2l != _lib.NA_INTEGER)
    assert dup.size > 5000 and (counts[dup] == counts[h2l[dup] - 1]).all()
    # planted patterns: every base-only 25-mer of the region the dictionary was cut from that
    # equals a pattern is reported (check through a hash of the windows of the first 2 Mbp)
    lo, hi = 20_000, 1_020_000
    win = np.lib.stride_tricks.sliding_window_view(subj[lo:hi], W)
    key = (win.astype(np.uint64) * (np.uint64(0x9E3779B97F4A7C15) ** np.arange(1, W + 1, dtype=np.uint64))).sum(axis=1)
    pkey = (pats.astype(np.uint64) * (np.uint64(0x9E3779B97F4A7C15) ** np.arange(1, W + 1, dtype=np.uint64))).sum(axis=1)
    hit_pos = np.flatnonzero(np.isin(key, pkey)) + lo + W                  # 1-based ends
    got = np.unique(e[(e >= lo + W) & (e < hi + 1)])
    assert set(hit_pos.tolist()) <= set(got.tolist()) and hit_pos.size > 1000
    # sharded == whole
    part = pd.parts()
    total = np.zeros(M, dtype=np.int64)
    for r in range(5):
        a0, b0 = sharding.shard_bounds(N, 5, r)
        a, b = sharding.chunk_range(a0, b0, N, W - 1, 0)
        with mp.DeviceSubject.from_chunk(subj[a:b], a, N, a0, b0) as ch:
            total += mp.match_parts(part, ch, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS)
    # the representative rows carry the counts at the .Call level
    rep = np.flatnonzero(h2l == _lib.NA_INTEGER)
    assert (total[rep] == counts[rep]).all()
    if ref.available():
        lo2, n2 = 30_000_000, 5_000_000
        sample = subj[lo2:lo2 + n2]
        sub = np.sort(np.random.default_rng(1).choice(M, 200_000, replace=False))
        od = R.PDict(R.SeqSet(concat=pats[sub].reshape(-1), widths=np.full(sub.size, W, np.int32)), backend="ref")
        pdsub = bs.PDict(bs.DNAStringSet(concat=pats[sub].reshape(-1), widths=np.full(sub.size, W, np.int32)))
        assert (bs.countPDict(pdsub, sample) == R.countPDict(od, sample)).all()


def test_mtb_full_size(world):
    subj, pats, dset = world
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pd2 = bs.PDict(dset, max_mismatch=2)
    pd0 = bs.PDict(dset)
    with mp.DeviceSubject.from_xstring(subj) as ds:
        c2 = bs.countPDict(pd2, ds, max_mismatch=2)
        assert _lib.lib().bsgpu_last_plan().decode().startswith("qgram_hamming")
        c21 = bs.countPDict(pd2, ds, max_mismatch=2, min_mismatch=1)
        c0 = bs.countPDict(pd0, ds)
        mi = bs.matchPDict(pd2, ds, max_mismatch=2)
    assert (c2 >= c0).all() and (c2 - c21 == c0).all() and c2.sum() > c0.sum()
    n = mi.elementNROWS()
    assert (n == c2).all()
    s, e = mi.unlist()
    rows = np.repeat(np.arange(M), n)
    inside = (s >= 1) & (e <= N)
    assert inside.all()                                  # 40 N-runs at both ends: nothing overhangs
    assert (mismatches(subj, pats, rows, e) <= 2).all()
