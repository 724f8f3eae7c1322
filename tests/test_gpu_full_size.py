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
  * a 5 Mbp window of it (200 000 of the patterns), counted by the compiled reference, agrees bit
    for bit -- for max.mismatch 0 AND for PDict(max.mismatch=2), which is the recall check of
    the q-gram filter at size (a missed alignment cannot hide behind the precision properties).
BASELINE configurations 3, 4, 5 at their full dictionary sizes: the same kind of sample parity
against the compiled reference (biostrings_b200/workloads.py generates the inputs).
"""
import warnings

import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp, sharding, synth, workloads
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
    if ref.available():
        # recall at size: every alignment the reference's three Trusted Bands (8/8/9) find
        # (tests/testthat/test-matchPDict.R:63-76 semantics), counts AND ends
        lo2, n2 = 30_000_000, 5_000_000
        sample = subj[lo2:lo2 + n2]
        sub = np.sort(np.random.default_rng(1).choice(M, 200_000, replace=False))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            od = R.PDict(R.SeqSet(concat=pats[sub].reshape(-1), widths=np.full(sub.size, W, np.int32)),
                         max_mismatch=2, backend="ref")
            pdsub = bs.PDict(bs.DNAStringSet(concat=pats[sub].reshape(-1), widths=np.full(sub.size, W, np.int32)),
                             max_mismatch=2)
        want = R.countPDict(od, sample, max_mismatch=2)
        got = bs.countPDict(pdsub, sample, max_mismatch=2)
        assert _lib.lib().bsgpu_last_plan().decode().startswith("qgram_hamming")
        assert want.sum() > 300 and (got == want).all()
        # the same window inside the resident 250 Mbp subject: counts of the sub-dictionary's rows
        # restricted to ends inside the window agree with the whole-subject ends
        sel = (e > lo2 + W) & (e <= lo2 + n2 - W)
        whole = np.bincount(rows[sel], minlength=M)[sub]
        we = R.matchPDict(od, sample, max_mismatch=2).endIndex()
        inner = np.array([0 if v is None else int(((v > W) & (v <= n2 - W)).sum()) for v in we])
        assert (whole == inner).all()


def _sample_parity(pd_gpu, pd_ref, sample, **kw):
    got = bs.countPDict(pd_gpu, sample, **kw)
    want = R.countPDict(pd_ref, sample, **kw)
    assert (got == want).all(), (int((got != want).sum()), int(got.sum()), int(want.sum()))
    return int(want.sum())


@pytest.mark.skipif(not ref.available(), reason="needs the compiled reference")
def test_c3_solexa_shape_sample_parity():
    """BASELINE config 3: 1M 36-mers, PDict(tb.start=1, tb.end=12), countPDict(max.mismatch=2) on
    100 Mbp: properties at size, and 200 000 of the patterns x a 4 Mbp window against the reference."""
    n = 100_000_000
    pats = workloads.c3_dictionary(1_000_000, n)
    subj = workloads.c3_subject(0, n, n)
    dset = bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(pats.shape[0], 36, np.int32))
    pd = bs.PDict(dset, tb_start=1, tb_end=12)
    with mp.DeviceSubject.from_xstring(subj) as ds:
        c0 = bs.countPDict(pd, ds)
        c2 = bs.countPDict(pd, ds, max_mismatch=2)
        c21 = bs.countPDict(pd, ds, max_mismatch=2, min_mismatch=1)
    assert (c2 - c21 == c0).all() and c2.sum() > c0.sum() > 20_000
    lo2, n2 = 48_000_000, 4_000_000
    sub = np.sort(np.random.default_rng(2).choice(pats.shape[0], 200_000, replace=False))
    sset = dict(concat=pats[sub].reshape(-1), widths=np.full(sub.size, 36, np.int32))
    pg = bs.PDict(bs.DNAStringSet(**sset), tb_start=1, tb_end=12)
    po = R.PDict(R.SeqSet(**sset), tb_start=1, tb_end=12, backend="ref")
    sample = subj[lo2:lo2 + n2]
    assert _sample_parity(pg, po, sample, max_mismatch=2) > 300
    assert _sample_parity(pg, po, sample, max_mismatch=2, min_mismatch=1) > 100
    assert _sample_parity(pg, po, sample) > 100


@pytest.mark.skipif(not ref.available(), reason="needs the compiled reference")
def test_c4_reads_sample_parity():
    """BASELINE config 4: vcountPDict of 100k 25-mers against 150 bp reads, collapse=1 / 2 on a
    100 000-read slice against the reference (weights 1 and a seeded integer weight)."""
    probes = workloads.c4_probes(100_000)
    reads = workloads.c4_reads(3_000_000, 3_100_000, probes)
    pg = bs.PDict(bs.DNAStringSet(concat=probes.reshape(-1), widths=np.full(probes.shape[0], 25, np.int32)))
    po = R.PDict(R.SeqSet(concat=probes.reshape(-1), widths=np.full(probes.shape[0], 25, np.int32)), backend="ref")
    rg = bs.DNAStringSet(concat=reads.reshape(-1), widths=np.full(reads.shape[0], 150, np.int32))
    ro = R.SeqSet(concat=reads.reshape(-1), widths=np.full(reads.shape[0], 150, np.int32))
    g = np.random.default_rng(4)
    # (the reference restarts its walk for every read: ~30 s per call on this slice)
    for collapse, w in ((1, g.integers(1, 9, reads.shape[0]).astype(np.int32)),
                        (2, np.ones(probes.shape[0], dtype=np.int32))):
        got = np.asarray(bs.vcountPDict(pg, rg, collapse=collapse, weight=w))
        want = np.asarray(R.vcountPDict(po, ro, collapse=collapse, weight=w))
        assert got.sum() > 4000 and np.array_equal(got, want)
    wg, wo = bs.vwhichPDict(pg, rg), R.vwhichPDict(po, ro)
    assert [v.tolist() for v in wg] == [v.tolist() for v in wo]


@pytest.mark.skipif(not ref.available(), reason="needs the compiled reference")
def test_c5_variable_width_iupac_sample_parity():
    """BASELINE config 5: 2M patterns of 15-40 letters, PDict(tb.start=1, tb.width=15), fixed=FALSE
    on a genome with isolated IUPAC letters: the whole dictionary on one 16 Mbp stretch (properties)
    and 200 000 of the patterns x a 2 Mbp window against the reference, fixed=FALSE and fixed=TRUE."""
    concat, widths = workloads.c5_dictionary(2_000_000)
    offs = np.concatenate([[0], np.cumsum(widths, dtype=np.int64)])
    B = synth.BLOCK
    nb = 3_100_000_000 // B - 1
    step = -(-nb // 64)
    lo = 20 * step * B                                   # a planted block, away from the ends
    subj = workloads.c5_genome_region(lo - 4_000_000, lo + 12_000_000)
    assert 8_000 < int((~np.isin(subj, synth.BASES)).sum()) < 24_000          # ~1e-3 ambiguity letters
    pd = bs.PDict(bs.DNAStringSet(concat=concat, widths=widths), tb_start=1, tb_width=15)
    with mp.DeviceSubject.from_xstring(subj) as ds, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cf = bs.countPDict(pd, ds, fixed=False)
        ct = bs.countPDict(pd, ds)
    assert (cf >= ct).all() and cf.sum() > ct.sum() > 1000
    sub = np.sort(np.random.default_rng(5).choice(widths.size, 200_000, replace=False))
    sc = np.concatenate([concat[offs[i]:offs[i + 1]] for i in sub.tolist()])
    sw = widths[sub]
    pg = bs.PDict(bs.DNAStringSet(concat=sc, widths=sw), tb_start=1, tb_width=15)
    po = R.PDict(R.SeqSet(concat=sc, widths=sw), tb_start=1, tb_width=15, backend="ref")
    sample = subj[4_000_000 - 500_000:4_000_000 + 1_500_000]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert _sample_parity(pg, po, sample, fixed=False) > 50
        assert _sample_parity(pg, po, sample) > 50
        assert _sample_parity(pg, po, sample, max_mismatch=1, fixed=False) > 50
