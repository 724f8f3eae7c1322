"""Committed golden fixtures (tests/golden/pdict_golden.json, generated from the reference's
own C by tests/golden/make_golden.py) against (a) our C restatement of the path, on CPU, and
(b) the CUDA path through the C ABI, on the GPU."""
import json
import os
import warnings

import numpy as np
import pytest

from oracle import rlevel as R

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "pdict_golden.json")) as f:
    GOLD = json.load(f)["cases"]


def _norm(ends):
    return [None if e is None else np.asarray(e).tolist() for e in ends]


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_port_oracle_reproduces_reference_fixtures(case):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pd = R.PDict(R.SeqSet(case["patterns"]), backend="port", **case["pdict"])
    for call in case["calls"]:
        kw = call["args"]
        if "set" in case:
            ss = R.SeqSet(case["set"])
            assert R.vcountPDict(pd, ss, **kw).tolist() == call["vcount"]
            assert [v.tolist() for v in R.vwhichPDict(pd, ss, **kw)] == call["vwhich"]
            continue
        subj = R.Views(case["subject"], *case["views"]) if "views" in case else case["subject"]
        assert R.countPDict(pd, subj, **kw).tolist() == call["counts"]
        assert R.whichPDict(pd, subj, **kw).tolist() == call["which"]
        assert _norm(R.matchPDict(pd, subj, **kw).endIndex()) == call["ends"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_cuda_path_reproduces_reference_fixtures(case):
    import biostrings_b200 as bs
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pd = bs.PDict(case["patterns"], **case["pdict"])
        for call in case["calls"]:
            kw = call["args"]
            if "set" in case:
                ss = bs.DNAStringSet(case["set"])
                assert bs.vcountPDict(pd, ss, **kw).tolist() == call["vcount"]
                assert [v.tolist() for v in bs.vwhichPDict(pd, ss, **kw)] == call["vwhich"]
                continue
            subj = bs.XStringViews(case["subject"], *case["views"]) if "views" in case else case["subject"]
            assert bs.countPDict(pd, subj, **kw).tolist() == call["counts"]
            assert bs.whichPDict(pd, subj, **kw).tolist() == call["which"]
            assert _norm(bs.matchPDict(pd, subj, **kw).endIndex()) == call["ends"]
