"""MIndex consumers (R/MIndex-class.R:52-103, src/MIndex_class.c:134-158) on the
GPU result against the oracle's restatement of the R methods run on the reference's own ends:
elementNROWS, startIndex / endIndex, x[[i]], unlist (with names), coverage, extractAllMatches --
for a Trusted-Band dictionary with whole-pattern duplicates and names, an MTB_PDict, a
variable-width dictionary with out-of-limits matches, and a non-preprocessed dictionary."""
import warnings

import numpy as np
import pytest

import biostrings_b200 as bs
from oracle import ref, rlevel as R
from tests.helpers import assert_same_ends, random_dna, rng

pytestmark = pytest.mark.gpu
BACKEND = "ref" if ref.available() else "port"


def consumers_agree(mg, mo, subject):
    assert mg.elementNROWS().tolist() == mo.elementNROWS().tolist()
    assert_same_ends(mg.endIndex(), mo.endIndex())
    assert_same_ends(mg.startIndex(), mo.startIndex())
    s, e, names = mg.unlist(use_names=True)
    os_, oe, onames = mo.unlist()
    assert s.tolist() == os_.tolist() and e.tolist() == oe.tolist()
    assert (None if names is None else list(names)) == onames
    width = len(subject) + 40
    assert mg.coverage(width).tolist() == mo.coverage(width).tolist()
    assert mg.coverage().tolist() == mo.coverage().tolist()
    v = bs.extractAllMatches(bs.DNAString(subject), mg)
    vs, vw, vn = R.extractAllMatches(subject, mo)
    assert v.start.tolist() == vs.tolist() and v.width.tolist() == vw.tolist()
    assert (None if v.names is None else list(v.names)) == vn
    for i in np.flatnonzero(mo.elementNROWS())[:20].tolist():
        gs, ge, gw = mg[i]
        assert ge.tolist() == mo.endIndex()[i].tolist() and gs.tolist() == mo.startIndex()[i].tolist()
        assert (gw == mg.width0[i]).all()


def test_tb_dictionary_with_duplicates_and_names():
    g = rng(8100)
    s = random_dna(g, 60_000)
    s[g.integers(0, s.size, 30)] = 15
    pats = [s[i:i + 25].copy() for i in g.integers(0, 59_000, 300)]
    for p in pats:
        p[p == 15] = 1
    pats += [pats[4].copy(), pats[4].copy(), pats[17].copy()] + [random_dna(g, 25) for _ in range(50)]
    names = [f"p{i}" for i in range(len(pats))]
    pg = bs.PDict(bs.DNAStringSet(pats, names=names))
    po = R.PDict(R.SeqSet(pats, names=names), backend=BACKEND)
    consumers_agree(bs.matchPDict(pg, s), R.matchPDict(po, s), s)
    pg = bs.PDict(bs.DNAStringSet(pats, names=names), tb_start=1, tb_end=10)
    po = R.PDict(R.SeqSet(pats, names=names), tb_start=1, tb_end=10, backend=BACKEND)
    consumers_agree(bs.matchPDict(pg, s, max_mismatch=2), R.matchPDict(po, s, max_mismatch=2), s)


def test_mtb_variable_width_and_plain_dictionaries():
    g = rng(8200)
    s = random_dna(g, 30_000)
    pats = [s[i:i + 25].copy() for i in g.integers(0, 29_000, 200)] + [s[:25].copy(), s[-25:].copy()]
    for p in pats[::2]:
        p[int(g.integers(0, 25))] = 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pg = bs.PDict(bs.DNAStringSet(pats), max_mismatch=2)
        po = R.PDict(R.SeqSet(pats), max_mismatch=2, backend=BACKEND)
    consumers_agree(bs.matchPDict(pg, s, max_mismatch=2), R.matchPDict(po, s, max_mismatch=2), s)
    # variable width, matches hanging over both ends of the subject
    vp = [s[i:i + int(g.integers(12, 40))].copy() for i in g.integers(0, 29_000, 150)] + [s[:30].copy(), s[-30:].copy()]
    pg = bs.PDict(bs.DNAStringSet(vp), tb_start=5, tb_width=8)
    po = R.PDict(R.SeqSet(vp), tb_start=5, tb_width=8, backend=BACKEND)
    consumers_agree(bs.matchPDict(pg, s, max_mismatch=3), R.matchPDict(po, s, max_mismatch=3), s)
    # non-preprocessed dictionary
    pl = vp[:40]
    consumers_agree(bs.matchPDict(bs.DNAStringSet(pl), s, max_mismatch=1),
                    R.matchPDict(R.SeqSet(pl), s, max_mismatch=1, backend=BACKEND), s)
