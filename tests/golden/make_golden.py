"""Generates tests/golden/pdict_golden.json from the REFERENCE ITSELF: the reference's own C
(compiled unmodified into oracle/_ref by oracle/Makefile, which needs /root/reference) driven
through the R-level restatement in oracle/rlevel.py.  The fixtures let every other
implementation (our C restatement, the CUDA path) be checked against reference outputs even
where /root/reference and oracle/_ref are absent.

    python tests/golden/make_golden.py      # rewrites pdict_golden.json

Inputs are stored literally (letters), outputs as counts / which / ends.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, rlevel as R          # noqa: E402

DEC = np.zeros(256, dtype=np.uint8)
for k, v in R.DNA_CODES.items():
    DEC[v] = ord(k)


def letters(codes):
    return DEC[np.asarray(codes, dtype=np.uint8)].tobytes().decode()


def rnd_dna(g, n):
    return np.array([1, 2, 4, 8], dtype=np.uint8)[g.integers(0, 4, n)]


def case(name, pats, subject, pdict_kw, calls, views=None, sset=None):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pd = R.PDict(R.SeqSet(pats), backend="ref", **pdict_kw)
    out = {"name": name, "patterns": [letters(p) for p in pats], "subject": letters(subject),
           "pdict": pdict_kw, "calls": []}
    if views is not None:
        out["views"] = views
    if sset is not None:
        out["set"] = [letters(x) for x in sset]
    for kw in calls:
        rec = {"args": kw}
        if sset is not None:
            rec["vcount"] = R.vcountPDict(pd, R.SeqSet(sset), **kw).tolist()
            rec["vwhich"] = [v.tolist() for v in R.vwhichPDict(pd, R.SeqSet(sset), **kw)]
        else:
            subj = R.Views(subject, views[0], views[1]) if views is not None else subject
            rec["counts"] = R.countPDict(pd, subj, **kw).tolist()
            rec["which"] = R.whichPDict(pd, subj, **kw).tolist()
            rec["ends"] = [None if e is None else e.tolist()
                           for e in R.matchPDict(pd, subj, **kw).endIndex()]
        out["calls"].append(rec)
    return out


def main():
    assert ref.available(), "oracle/_ref is not built (needs /root/reference)"
    g = np.random.Generator(np.random.PCG64(20261019))
    cases = []
    s = rnd_dna(g, 3000)
    s[100:130] = 15
    s[2990:] = 15
    amb = np.array([3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15], dtype=np.uint8)
    # exact, whole-pattern TB, with duplicates
    pats = [s[i:i + 25].copy() for i in g.integers(140, 2900, 60)] + [rnd_dna(g, 25) for _ in range(40)]
    pats[7] = pats[2].copy()
    pats[50] = pats[2].copy()
    cases.append(case("exact_25mers", pats, s, {}, [{}]))
    cases.append(case("exact_12mers_twobit", [p[:12] for p in pats], s, {"algorithm": "Twobit"}, [{}]))
    # TB + tail, mismatches, min.mismatch
    pats = [s[i:i + 36].copy() for i in g.integers(140, 2900, 80)]
    for p in pats[:50]:
        for _ in range(int(g.integers(0, 4))):
            p[int(g.integers(12, 36))] = [1, 2, 4, 8][int(g.integers(4))]
    cases.append(case("tb12_tail24", pats, s, {"tb_start": 1, "tb_end": 12},
                      [{"max_mismatch": 0}, {"max_mismatch": 2}, {"max_mismatch": 3, "min_mismatch": 2}]))
    # head + TB + tail, variable width, IUPAC letters in flanks and subject, all fixed modes
    s2 = s.copy()
    pos = g.integers(140, 2900, 40)
    s2[pos] = amb[g.integers(0, len(amb), 40)]
    s2[500] = 16
    pats = []
    for i in g.integers(140, 2900, 80):
        w = int(g.integers(15, 41))
        p = s[i:i + w].copy()
        p[p == 15] = 1
        if g.random() < 0.4:
            p[int(g.integers(12, w))] = amb[int(g.integers(len(amb)))]
        if g.random() < 0.3:
            p[int(g.integers(0, 2))] = 15
        pats.append(p)
    cases.append(case("var_width_iupac", pats, s2, {"tb_start": 3, "tb_width": 10},
                      [{"max_mismatch": 1, "fixed": f} for f in (True, False, "pattern", "subject")]))
    # MTB_PDict
    pats = [s[i:i + 25].copy() for i in g.integers(140, 2900, 80)]
    for p in pats:
        for _ in range(int(g.integers(0, 4))):
            p[int(g.integers(0, 25))] = [1, 2, 4, 8][int(g.integers(4))]
    pats[9] = pats[1].copy()
    cases.append(case("mtb_k2_25mers", pats, s, {"max_mismatch": 2},
                      [{"max_mismatch": 2}, {"max_mismatch": 1}, {"max_mismatch": 2, "min_mismatch": 1}]))
    # out-of-limits matches
    pats = [np.concatenate([rnd_dna(g, 6), s[0:14]]), np.concatenate([s[2970:2990], rnd_dna(g, 5)])]
    s3 = s[:2990]
    cases.append(case("out_of_limits", pats, s3, {"tb_start": 7, "tb_end": 14},
                      [{"max_mismatch": 0}, {"max_mismatch": 6}]))
    # views (overlapping, unordered, empty)
    pats = [s[i:i + 20].copy() for i in g.integers(140, 2900, 60)]
    cases.append(case("views", pats, s, {"tb_end": 10}, [{"max_mismatch": 1}],
                      views=([1500, 1, 1400, 200], [800, 1000, 300, 0])))
    # sets
    reads = [s[i:i + int(g.integers(0, 90))].copy() for i in g.integers(140, 2800, 40)]
    cases.append(case("sets", pats, s, {"tb_end": 10}, [{"max_mismatch": 1}], sset=reads))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pdict_golden.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference C) + oracle/rlevel.py",
                   "cases": cases}, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
