"""Dense-hit regime (the SolexaYi2 shape: /root/reference/vignettes/SolexaYi2-benchmark.txt:127-162 --
4.47 M 35-mer reads against mouse chr1, 110.6 M exact matches = 0.56 per bp; 52 s for matchPDict and
23 s for elementNROWS in the reference): M reads of 35 letters against a synthetic subject whose
repeat families give every repeat-derived read ~150 matches, spread over the whole subject.

    python tools/bench_dense.py [--mbp 200] [--reads 4000000] [--families 2000] [--copies 200]

One JSON line: countPDict and matchPDict (ends) times on the resident subject, hits, hits per
second, the MIndex consumers on the CSR, and the parity of a sample (a slice of the dictionary x a
window of the subject) against the compiled reference.
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biostrings_b200 as bs                                    # noqa: E402
from biostrings_b200 import _lib, matchpdict as mp, synth       # noqa: E402

W = 35
UNIT = 300


def make_inputs(n, M, families, copies, seed=21):
    g = np.random.Generator(np.random.PCG64(seed))
    subj = synth.subject_region(0, n, n, seed=seed, with_n=False)
    units = synth.BASES[g.integers(0, 4, size=(families, UNIT), dtype=np.uint8)]
    # copies of every family at random places over the whole subject (a later copy may clip an earlier one)
    pos = g.integers(0, n - UNIT, size=families * copies)
    fam = np.repeat(np.arange(families), copies)
    for p, f in zip(pos.tolist(), fam.tolist()):
        subj[p:p + UNIT] = units[f]
    n_rep = int(M * 0.15)
    reads = np.empty((M, W), dtype=np.uint8)
    f = g.integers(0, families, n_rep)
    o = g.integers(0, UNIT - W + 1, n_rep)
    reads[:n_rep] = units[f[:, None], o[:, None] + np.arange(W)[None, :]]
    n_uni = int(M * 0.35)
    st = g.integers(0, n - W, n_uni)
    reads[n_rep:n_rep + n_uni] = subj[st[:, None] + np.arange(W)[None, :]]
    reads[n_rep + n_uni:] = synth.BASES[g.integers(0, 4, size=(M - n_rep - n_uni, W), dtype=np.uint8)]
    g.shuffle(reads, axis=0)
    return subj, reads


def timed(fn, reps=3):
    import torch
    best, out = None, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=200.0)
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--families", type=int, default=2000)
    ap.add_argument("--copies", type=int, default=200)
    a = ap.parse_args()
    n, M = int(a.mbp * 1e6), a.reads
    t0 = time.perf_counter()
    subj, reads = make_inputs(n, M, a.families, a.copies)
    t_gen = time.perf_counter() - t0
    dset = bs.DNAStringSet(concat=reads.reshape(-1), widths=np.full(M, W, np.int32))
    t0 = time.perf_counter()
    pd = bs.PDict(dset)
    t_pdict = time.perf_counter() - t0
    parts = pd.parts()
    L = _lib.lib()
    with mp.DeviceSubject.from_xstring(subj) as ds:
        t0 = time.perf_counter()
        mp.match_parts(parts, ds, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS)        # builds the engine's index
        t_first = time.perf_counter() - t0
        counts, t_counts = timed(lambda: mp.match_parts(parts, ds, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS))
        plan = L.bsgpu_last_plan().decode()
        (off, ends), t_ends = timed(lambda: mp.match_parts(parts, ds, 0, 0, (True, True), _lib.MATCHES_AS_ENDS), 2)
        mi, t_match = timed(lambda: bs.matchPDict(pd, ds), 1)
    hits = int(counts.sum())
    t0 = time.perf_counter()
    nrows = mi.elementNROWS()
    t_nrows = time.perf_counter() - t0
    t0 = time.perf_counter()
    s, e = mi.unlist()
    t_unlist = time.perf_counter() - t0
    t0 = time.perf_counter()
    cov = mi.coverage(n)
    t_cov = time.perf_counter() - t0
    assert int(nrows.sum()) == s.size and int(off[-1]) == hits
    # every reported end spells its read (a sample of the rows)
    g = np.random.default_rng(3)
    rows = np.repeat(np.arange(M), np.diff(mi.offsets))
    take = g.choice(mi.ends.size, min(mi.ends.size, 2_000_000), replace=False)
    idx = (mi.ends[take].astype(np.int64) - W)[:, None] + np.arange(W)[None, :]
    spelled = bool((subj[idx] == reads[rows[take]]).all())
    parity = None
    try:
        from oracle import ref, rlevel as R
        if ref.available():
            sub = np.sort(g.choice(M, 100_000, replace=False))
            lo = n // 3
            sample = subj[lo:lo + 2_000_000]
            skw = dict(concat=reads[sub].reshape(-1), widths=np.full(sub.size, W, np.int32))
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                po = R.PDict(R.SeqSet(**skw), backend="ref")
                pg = bs.PDict(bs.DNAStringSet(**skw))
            t0 = time.perf_counter()
            want = R.countPDict(po, sample)
            t_ref = time.perf_counter() - t0
            got = bs.countPDict(pg, sample)
            we = R.matchPDict(po, sample).endIndex()
            ge = bs.matchPDict(pg, sample).endIndex()
            same_ends = all((a is None) == (b is None) and (a is None or a.tolist() == b.tolist())
                            for a, b in zip(ge, we))
            parity = {"counts_equal": bool((got == want).all()), "ends_equal": bool(same_ends),
                      "sample_patterns": int(sub.size), "sample_letters": int(sample.size),
                      "sample_matches": int(want.sum()), "reference_seconds_counts": t_ref}
    except Exception as ex:                                     # noqa: BLE001
        parity = {"error": str(ex)}
    print(json.dumps({
        "workload": f"{M} reads of {W} letters (15 % from {a.families} repeat families with {a.copies} copies each, "
                    f"35 % unique cuts, 50 % random) x {n / 1e6:g} Mbp, exact",
        "hits": hits, "hits_per_bp": hits / n, "plan": plan,
        "countPDict": {"seconds": t_counts, "Gbp_s": n / t_counts / 1e9, "Mhits_s": hits / t_counts / 1e6},
        "matchPDict_ends_C_call": {"seconds": t_ends, "Gbp_s": n / t_ends / 1e9, "Mhits_s": hits / t_ends / 1e6},
        "matchPDict_python": {"seconds": t_match},
        "mindex_consumers_s": {"elementNROWS": t_nrows, "unlist": t_unlist, "coverage": t_cov},
        "every_sampled_end_spells_its_read": spelled, "parity_with_reference_on_sample": parity,
        "build": {"generate_s": t_gen, "pdict_s": t_pdict, "first_call_s": t_first},
        "reference_published": {"matchPDict_s": 52.4, "elementNROWS_s": 23.2, "hits": 110578790,
                                "source": "vignettes/SolexaYi2-benchmark.txt:127-136 (4.47 M 35-mers x mouse chr1, other hardware)"},
    }), flush=True)


if __name__ == "__main__":
    main()
