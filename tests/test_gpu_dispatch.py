"""The drop-in boundary, tested where the reference draws it: the `.Call` entry points.

biostrings_b200/csrc/match_pdict_gpu.c implements build_Twobit / ACtree2_build /
match_PDict3Parts_XString / match_PDict3Parts_XStringViews / vmatch_PDict3Parts_XStringSet /
ByPos_MIndex_* with the reference's names and SEXP signatures on top of libbsgpu.  It is
compiled against the same stand-in R headers as the reference's own C and driven by the same
harness (oracle/ref_harness.c), so here both libraries get the same R objects and their SEXP
results are compared directly -- including the R-level restatement (oracle/rlevel.py) run on
top of each of them.
"""
import numpy as np
import pytest

from oracle import ref, rlevel as R
from tests.helpers import IUPAC_AMBIG, assert_same_ends, random_dna, rng

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref.available("ref") and ref.available("dispatch")),
                                 reason="needs oracle/_ref/libbiostrings_ref.so and "
                                        "libbiostrings_gpu_dispatch.so")]


def pair(pats, **kw):
    """The same PDict() over the reference's C and over the GPU dispatch."""
    ref.use("ref")
    a = R.PDict(R.SeqSet(pats), backend="ref", **kw)
    ref.use("dispatch")
    try:
        b = R.PDict(R.SeqSet(pats), backend="ref", **kw)
    finally:
        ref.use("ref")
    return a, b


def same(a, b, subj, **kw):
    assert R.countPDict(a, subj, **kw).tolist() == R.countPDict(b, subj, **kw).tolist()
    assert R.whichPDict(a, subj, **kw).tolist() == R.whichPDict(b, subj, **kw).tolist()
    ma, mb = R.matchPDict(a, subj, **kw), R.matchPDict(b, subj, **kw)
    assert_same_ends(ma.ends, mb.ends)             # raw SEXP lists, NULL placement included
    assert_same_ends(ma.endIndex(), mb.endIndex())


def test_golden_vectors_through_the_dot_call_boundary():
    a, b = pair(["acgt", "gt", "cgt", "ac"], tb_end=2)
    for mm in (0, 1, 2):
        same(a, b, "acggaccg", max_mismatch=mm)
    assert R.countPDict(b, "acggaccg", max_mismatch=1).tolist() == [1, 0, 2, 2]
    a, b = pair(["TACCNG", "TAGT", "CGGNT", "AGTAG", "TAGT"], tb_end=3)
    for kw in ({}, {"fixed": False}, {"max_mismatch": 1}, {"max_mismatch": 1, "fixed": False},
               {"max_mismatch": 2}):
        same(a, b, "TAGTACCAGTTTCGGG", **kw)
    assert R.countPDict(b, "TAGTACCAGTTTCGGG", max_mismatch=2).tolist() == [1, 1, 1, 2, 1]
    a, b = pair(["AA", "AT", "TA", "TT"], algorithm="Twobit")
    same(a, b, "ATTATTAATTAA")
    assert b.duplicated().tolist() == a.duplicated().tolist()


def test_error_messages_through_the_dot_call_boundary():
    ref.use("dispatch")
    try:
        with pytest.raises(ref.RefError, match="non base DNA letter found in Trusted Band for pattern 2"):
            R.PDict(["ACGT", "ANGT"], backend="ref")
        with pytest.raises(ref.RefError, match="element 2 in Trusted Band has a different length"):
            R.PDict(["ACGT", "ACG"], backend="ref")
        with pytest.raises(ref.RefError, match="must be <= 14"):
            R.PDict(["ACGTACGTACGTACGT"], algorithm="Twobit", backend="ref")
        pd = R.PDict(["ACGT", "CCGT"], algorithm="Twobit", tb_end=3, backend="ref")
        with pytest.raises(ref.RefError, match="cannot treat IUPAC extended letters in the subject"):
            R.countPDict(pd, "ACGTACGT", fixed="pattern")
        pd = R.PDict(["ACGTA", "CCGTA"], tb_end=3, backend="ref")
        with pytest.raises(ref.RefError, match="out of limits"):
            R.countPDict(pd, R.Views("ACGTACGT", [5], [10]))
    finally:
        ref.use("ref")


@pytest.mark.parametrize("algo", ["ACtree2", "Twobit"])
def test_random_dictionaries(algo):
    g = rng(303)
    s = random_dna(g, 30_000)
    s[700:720] = 15
    amb = g.random(s.size) < 0.003
    s[amb] = IUPAC_AMBIG[g.integers(0, len(IUPAC_AMBIG), int(amb.sum()))]
    pats = [s[i:i + 30].copy() for i in g.integers(0, 29_000, 600)] + [random_dna(g, 30) for _ in range(400)]
    for p in pats:
        p[~np.isin(p, [1, 2, 4, 8])] = 1
    for p in pats[:300]:
        p[int(g.integers(12, 30))] = 4
    pats[11] = pats[3].copy()
    a, b = pair(pats, tb_start=1, tb_end=12, algorithm=algo)
    for mm in (0, 2):
        same(a, b, s, max_mismatch=mm)
    same(a, b, s, max_mismatch=2, min_mismatch=1)
    if algo == "ACtree2":
        same(a, b, s, max_mismatch=1, fixed=False)
    v = R.Views(s, [1, 9000, 5000], [8000, 3000, 6000])
    same(a, b, v, max_mismatch=1)
    # whole-pattern TB: dups0 + blanked dups
    a, b = pair(pats, algorithm=algo) if algo == "ACtree2" else pair([p[:12] for p in pats], algorithm=algo)
    same(a, b, s)
    # MTB_PDict: the union is done by ByPos_MIndex_combine of the dispatch library
    a, b = pair(pats, max_mismatch=2, algorithm=algo)        # bands of 10 letters
    same(a, b, s, max_mismatch=2)
    same(a, b, s, max_mismatch=1)


def test_sets():
    g = rng(404)
    genome = random_dna(g, 20_000)
    reads = [genome[i:i + int(g.integers(0, 120))].copy() for i in g.integers(0, 19_800, 150)]
    pats = [genome[i:i + 20].copy() for i in g.integers(0, 19_900, 300)]
    a, b = pair(pats, tb_end=10)
    sa = R.SeqSet(reads)
    for coll, w in ((False, 1), (1, g.integers(1, 4, 150)), (2, g.integers(1, 4, 300)),
                    (1, g.random(150)), (2, g.random(300))):
        x = R.vcountPDict(a, sa, max_mismatch=1, collapse=coll, weight=w)
        y = R.vcountPDict(b, sa, max_mismatch=1, collapse=coll, weight=w)
        assert x.dtype == y.dtype and np.array_equal(x, y)
    assert [v.tolist() for v in R.vwhichPDict(a, sa, max_mismatch=1)] == \
        [v.tolist() for v in R.vwhichPDict(b, sa, max_mismatch=1)]


def test_mindex_combine_entry_point():
    """ByPos_MIndex_combine (src/MIndex_class.c:230) of both libraries on the per-band
    results of an MTB_PDict, called the way R/matchPDict.R:335-377 calls it."""
    g = rng(505)
    s = random_dna(g, 8000)
    pats = [s[i:i + 24].copy() for i in g.integers(0, 7900, 200)]
    for p in pats[:120]:
        p[int(g.integers(0, 24))] = 8
    outs = {}
    for which in ("ref", "dispatch"):
        ref.use(which)
        try:
            pd = R.PDict(R.SeqSet(pats), max_mismatch=2, backend="ref")
            for tp in pd.threeparts_list:
                tp.pptb.match_xstring(tp.head_or_none(), tp.tail_or_none(), s, 2, 0, (True, True),
                                      ref.MATCHES_AS_ENDS, keep=True)
                ref.stash_result()
            outs[which] = ref.mindex_combine()
        finally:
            ref.use("ref")
    assert_same_ends(outs["ref"], outs["dispatch"])
    assert sum(e is not None for e in outs["ref"]) > 50


def test_oligo_frequency_through_the_dot_call_boundary():
    """XString_oligo_frequency / XStringSet_oligo_frequency (src/letter_frequency.c:634,667)
    of biostrings_b200/csrc/letter_frequency_gpu.c against the reference's own: values, dim,
    names / dimnames of every result shape, and the error text."""
    g = rng(1234)
    x = random_dna(g, 20_000)
    x[g.integers(0, x.size, 300)] = IUPAC_AMBIG[g.integers(0, len(IUPAC_AMBIG), 300)]
    widths = g.integers(0, 200, 150).astype(np.int32)
    concat = x[:int(widths.sum())]

    def both(fn, *a, **kw):
        out = []
        for which in ("ref", "dispatch"):
            ref.use(which)
            try:
                out.append(fn(*a, **kw))
            finally:
                ref.use("ref")
        return out

    def same_table(a, b):
        (va, da, la), (vb, db, lb) = a, b
        assert da == db and la == lb
        if isinstance(va, list):
            assert len(va) == len(vb)
            for p, q in zip(va, vb):
                assert p.dtype == q.dtype and np.array_equal(p, q)
        else:
            assert va.dtype == vb.dtype and np.array_equal(va, vb)

    for width, step in [(1, 1), (2, 1), (3, 3), (4, 2), (6, 1), (8, 5)]:
        for kw in ({}, {"as_prob": True}, {"fast_moving_side": "left"}, {"with_labels": False},
                   {"as_array": True, "fast_moving_side": "left"}):
            if kw.get("as_array") and width > 4:
                continue
            same_table(*both(ref.oligo_frequency_xstring, x, width, step, **kw))
        for simplify_as in ("matrix", "list", "collapsed"):
            for kw in ({}, {"as_prob": True}, {"with_labels": False}):
                if width > 6:
                    continue
                same_table(*both(ref.oligo_frequency_set, concat, widths, width, step,
                                 simplify_as=simplify_as, **kw))
    same_table(*both(ref.oligo_frequency_set, concat, widths, 2, simplify_as="list", as_array=True,
                     fast_moving_side="left"))
    ref.use("dispatch")
    try:
        with pytest.raises(ref.RefError, match="'buflength' must be >= 1 and <= 15"):
            ref.oligo_frequency_xstring(x, 16)
    finally:
        ref.use("ref")


def test_read_fasta_files_through_the_dot_call_boundary():
    """read_fasta_files (src/read_fasta_files.c:412) of biostrings_b200/csrc/read_fasta_files_gpu.c
    against the reference's own on the same in-memory files: widths, letters, names, the
    invalid-letter warning, nrec / skip / seek.first.rec and the error texts."""
    from tests.test_fasta_hostlogic import TEXTS, rand_fasta

    def both(text, **kw):
        out = []
        for which in ("ref", "dispatch"):
            ref.use(which)
            try:
                out.append(ref.read_fasta(text, **kw))
            finally:
                ref.use("ref")
        return out

    def same_set(a, b):
        assert a[0].tolist() == b[0].tolist() and np.array_equal(a[1], b[1])
        assert a[2] == b[2] and a[3] == b[3]

    for text in TEXTS:
        same_set(*both(text))
        same_set(*both(text, use_names=False))
    g = rng(808)
    for k in range(6):
        text = rand_fasta(g, int(g.integers(1, 30)), crlf=bool(k % 2), line=int(g.integers(1, 600)))
        same_set(*both(text))
        for nrec, skip in [(3, 0), (-1, 2), (2, 4), (0, 0), (50, 1)]:
            same_set(*both(text, nrec=nrec, skip=skip))
            same_set(*both(b"junk\n" + text, nrec=nrec, skip=skip, seek_first_rec=True))
    ref.use("dispatch")
    try:
        with pytest.raises(ref.RefError, match='reading FASTA file <memory>: ">" expected at beginning of line 4'):
            ref.read_fasta(b"\n;c\n\nAC\n>r\n")
        with pytest.raises(ref.RefError, match="no FASTA record found"):
            ref.read_fasta(b"junk\n", seek_first_rec=True)
    finally:
        ref.use("ref")


def _dispatch_lib():
    ref.use("dispatch")
    try:
        return ref.lib()
    finally:
        ref.use("ref")


@pytest.mark.parametrize("algo", ["ACtree2", "Twobit"])
def test_device_index_lives_in_the_object(algo):
    """The device index hangs on an external pointer of the PreprocessedTB object (finalizer frees
    it); an object that comes back without it (readRDS) is rebuilt from its own slots, including
    which patterns the build skipped (pp_exclude), and gives the same answers."""
    g = rng(310)
    s = random_dna(g, 20_000)
    w = 12 if algo == "Twobit" else 20
    pats = [s[i:i + w + 6].copy() for i in g.integers(0, 19_000, 120)]
    pats += [pats[3].copy(), pats[3].copy(), np.concatenate([pats[5][:w], random_dna(g, 6)])]   # dups0 + a TB duplicate
    a, b = pair(pats, algorithm=algo, tb_start=1, tb_end=w)
    L = _dispatch_lib()
    live0 = L.ref_dispatch_live_indexes()
    assert live0 >= 1
    want = {mm: R.countPDict(a, s, max_mismatch=mm).tolist() for mm in (0, 2)}
    for mm in (0, 2):
        assert R.countPDict(b, s, max_mismatch=mm).tolist() == want[mm]
    # "readRDS": the pointer is gone, the next call rebuilds the index from the object
    ref.use("dispatch")
    try:
        for tp in b.threeparts_list:
            assert L.ref_pptb_detach_device(tp.pptb._h) == 0
        assert L.ref_dispatch_live_indexes() == live0 - len(b.threeparts_list)
        for mm in (0, 2):
            assert R.countPDict(b, s, max_mismatch=mm).tolist() == want[mm]
            assert R.whichPDict(b, s, max_mismatch=mm).tolist() == R.whichPDict(a, s, max_mismatch=mm).tolist()
        assert L.ref_dispatch_live_indexes() == live0
        # garbage collection of the object frees the device index
        for tp in b.threeparts_list:
            tp.pptb.close()
        assert L.ref_dispatch_live_indexes() == live0 - len(b.threeparts_list)
    finally:
        ref.use("ref")


def test_indels_through_the_dot_call_boundary():
    g = rng(311)
    s = random_dna(g, 5_000)
    pats = []
    for _ in range(30):
        a0 = int(g.integers(0, 4_900))
        p = s[a0:a0 + int(g.integers(8, 25))].copy()
        if g.random() < 0.5:
            p = np.delete(p, int(g.integers(0, len(p))))
        pats.append(p)
    pc, pw = np.concatenate(pats), np.array([len(p) for p in pats], dtype=np.int32)
    reads = [s[i:i + 200].copy() for i in g.integers(0, 4_700, 12)]
    sc, sw = np.concatenate(reads), np.array([len(r) for r in reads], dtype=np.int32)
    out = {}
    for which in ("ref", "dispatch"):
        ref.use(which)
        try:
            out[which] = (
                ref.match_set_xstring(pc, pw, s, 2, 0, True, (True, True), "indels", ref.MATCHES_AS_COUNTS).tolist(),
                ref.match_set_xstring(pc, pw, s, 1, 0, True, (True, True), "indels", ref.MATCHES_AS_WHICH).tolist(),
                ref.match_set_views(pc, pw, s, [1, 2000], [1500, 2500], 2, 0, True, (True, True), "indels",
                                    ref.MATCHES_AS_COUNTS).tolist(),
                np.asarray(ref.vmatch_set_set(pc, pw, sc, sw, 2, 0, True, (True, True), "indels", 0,
                                              np.ones(1, np.int32), ref.MATCHES_AS_COUNTS)).tolist(),
                np.asarray(ref.vmatch_set_set(pc, pw, sc, sw, 2, 0, True, (True, True), "indels", 1,
                                              np.arange(1, len(reads) + 1, dtype=np.int32), ref.MATCHES_AS_COUNTS)).tolist(),
                np.asarray(ref.vmatch_set_set(pc, pw, sc, sw, 2, 0, True, (True, True), "indels", 2,
                                              np.linspace(0.5, 2.0, len(pats)), ref.MATCHES_AS_COUNTS)).tolist())
        finally:
            ref.use("ref")
    assert out["ref"] == out["dispatch"] and sum(out["ref"][0]) > 10
