"""Times the other BASELINE.json configurations (C3, C3', C4, C5) on one
GPU -- they are parity-test shapes, not the headline bench line, but their timings tell
which kernel needs work next.  Prints one JSON line per configuration.

    python tools/bench_configs.py [--scale 1.0] [--only C3,C4]
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
from biostrings_b200 import _lib, matchpdict as mp, synth, workloads as W       # noqa: E402

L = _lib.lib()


def timed(fn, reps=3):
    import torch
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out, best


def dna_set(pats):
    return bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(pats.shape[0], pats.shape[1], np.int32))


def c3(scale):
    """countPDict of 1M 36-mers, tb.start=1, tb.end=12, max.mismatch=2, 100 Mbp."""
    n = int(100e6 * scale)
    M = int(1e6 * scale) or 1000
    subj = W.c3_subject(0, n, n)
    pats = W.c3_dictionary(M, n)
    t0 = time.perf_counter()
    pd = bs.PDict(dna_set(pats), tb_start=1, tb_end=12)
    t_build = time.perf_counter() - t0
    ds = mp.DeviceSubject.from_xstring(subj)
    out = {}
    for mm in (0, 2):
        c, dt = timed(lambda: mp.match_parts(pd.parts(), ds, mm, 0, (True, True), _lib.MATCHES_AS_COUNTS))
        out[f"mm{mm}"] = {"seconds": dt, "Gbp_s": n / dt / 1e9, "matches": int(c.sum()),
                          "plan": L.bsgpu_last_plan().decode()}
    return {"config": "C3", "letters": n, "patterns": M, "pdict_build_s": t_build, **out}


def c3prime(scale):
    """countPDict of 1M 25-mers as PDict(max.mismatch=2) (bands 8/8/9), 250 Mbp."""
    n = int(250e6 * scale)
    M = int(1e6 * scale) or 1000
    subj = synth.subject_region(0, n, n)
    pats = synth.dictionary(M, 25, n, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        pd = bs.PDict(dna_set(pats), max_mismatch=2)
        t_build = time.perf_counter() - t0
    ds = mp.DeviceSubject.from_xstring(subj)
    c, dt = timed(lambda: mp.match_parts(pd.parts(), ds, 2, 0, (True, True), _lib.MATCHES_AS_COUNTS))
    (_, e), dte = timed(lambda: mp.match_parts(pd.parts(), ds, 2, 0, (True, True), _lib.MATCHES_AS_ENDS), 1)
    return {"config": "C3'", "letters": n, "patterns": M, "pdict_build_s": t_build,
            "counts": {"seconds": dt, "Gbp_s": n / dt / 1e9, "matches": int(c.sum())},
            "ends": {"seconds": dte, "hits": int(e.size)}, "plan": L.bsgpu_last_plan().decode()}


def c4(scale):
    """vcountPDict of 100k 25-mers against 10M x 150 bp reads, collapse=1 and 2."""
    nreads = int(1e7 * scale)
    M = int(1e5 * scale) or 1000
    probes = W.c4_probes(M)
    reads = W.c4_reads(0, nreads, probes)
    t0 = time.perf_counter()
    pd = bs.PDict(dna_set(probes))
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    ds = mp.DeviceSubject.from_set(dna_set(reads))
    t_up = time.perf_counter() - t0
    out = {}
    for coll in (1, 2):
        w = np.ones(nreads if coll == 1 else M, dtype=np.int32)
        c, dt = timed(lambda: mp.vmatch_parts(pd.parts(), ds, 0, 0, (True, True), coll, w,
                                              _lib.MATCHES_AS_COUNTS))
        out[f"collapse{coll}"] = {"seconds": dt, "Gbp_s": nreads * 150 / dt / 1e9,
                                  "matches": int(c.sum())}
    return {"config": "C4", "reads": nreads, "patterns": M, "pdict_build_s": t_build,
            "upload_s": t_up, **out, "plan": L.bsgpu_last_plan().decode()}


def c5(scale):
    """2M patterns of 15-40 letters, tb.start=1, tb.width=15, fixed=FALSE, one 1/8 shard
    (387.5 Mbp) of a 3.1 Gbp genome with isolated IUPAC letters at rate 1e-3."""
    n = int(387.5e6 * scale)
    M = int(2e6 * scale) or 1000
    subj = W.c5_genome_region(0, n)
    concat, widths = W.c5_dictionary(M, max(n, 64 << 20))
    t0 = time.perf_counter()
    pd = bs.PDict(bs.DNAStringSet(concat=concat, widths=widths), tb_start=1, tb_width=15)
    t_build = time.perf_counter() - t0
    ds = mp.DeviceSubject.from_xstring(subj)
    out = {}
    for fixed, name in (((True, True), "fixed_TRUE"), ((False, False), "fixed_FALSE")):
        c, dt = timed(lambda: mp.match_parts(pd.parts(), ds, 0, 0, fixed, _lib.MATCHES_AS_COUNTS), 4)
        out[name] = {"seconds": dt, "Gbp_s": n / dt / 1e9, "matches": int(c.sum()),
                     "plan": L.bsgpu_last_plan().decode()}
    return {"config": "C5 (one of 8 shards)", "letters": n, "patterns": M, "pdict_build_s": t_build, **out}


def p1(scale):
    """Non-preprocessed dictionary: 1000 patterns of 15-40 letters, 2 % IUPAC
    letters, countPDict(max.mismatch=2, fixed=FALSE) against 10 Mbp; the reference's own loop
    (one matchPattern per pattern) is timed on one host core over 20 patterns."""
    n = int(10e6 * scale)
    M = max(20, int(1000 * scale))
    g = np.random.Generator(np.random.PCG64(11))
    subj = synth.subject_region(0, n, n, seed=12, with_n=False)
    widths = g.integers(15, 41, M)
    pats = []
    for w in widths.tolist():
        st = int(g.integers(0, n - w))
        q = subj[st:st + w].copy()
        amb = g.random(w) < 0.02
        q[amb] = np.array([15, 3, 5, 10], dtype=np.uint8)[g.integers(0, 4, int(amb.sum()))]
        pats.append(q)
    pset = bs.DNAStringSet(pats)
    ds = mp.DeviceSubject.from_xstring(subj)
    out = {}
    for mm, fixed in ((0, True), (2, False)):
        c, dt = timed(lambda: bs.countPDict(pset, ds, max_mismatch=mm, fixed=fixed), 2)
        out[f"mm{mm}_fixed{fixed}"] = {"seconds": dt, "Gbp_s": n / dt / 1e9,
                                       "pattern_x_letters_per_s": M * n / dt,
                                       "matches": int(c.sum()), "plan": L.bsgpu_last_plan().decode()}
    cpu = None
    try:
        from oracle import ref, rlevel as R
        if ref.available():
            k = 20
            oset = R.SeqSet(pats[:k])
            t0 = time.perf_counter()
            want = R.countPDict(oset, subj, max_mismatch=2, fixed=False, backend="ref")
            dt = time.perf_counter() - t0
            got = bs.countPDict(bs.DNAStringSet(pats[:k]), ds, max_mismatch=2, fixed=False)
            cpu = {"patterns": k, "seconds": dt, "pattern_x_letters_per_s": k * n / dt,
                   "algo": "shift-or (auto)", "cores": 1, "parity_with_gpu": bool((got == want).all())}
    except Exception as e:                                      # noqa: BLE001
        cpu = {"error": str(e)}
    return {"config": "P1 (non-preprocessed dictionary)", "letters": n, "patterns": M, **out,
            "cpu_reference": cpu}


def p2(scale):
    """A LARGE non-preprocessed dictionary: the headline's 1M base-only 25-mers handed over as a plain
    DNAStringSet (no PDict() call), countPDict against 250 Mbp.  The library gives it a Trusted-Band
    index of its own on first use (plain_auto_index); the brute force would need M x L = 2.5e14 letter
    pairs (~11 min at 3.8e11 per second)."""
    n = int(250e6 * scale)
    M = max(1000, int(1e6 * scale))
    pats = synth.dictionary(M, 25, n, 1)
    subj = synth.subject_region(0, n, n)
    pset = dna_set(pats)
    ds = mp.DeviceSubject.from_xstring(subj)
    t0 = time.perf_counter()
    c = bs.countPDict(pset, ds)
    t_first = time.perf_counter() - t0
    c, dt = timed(lambda: bs.countPDict(pset, ds), 3)
    return {"config": "P2 (1M base-only 25-mers as a plain DNAStringSet)", "letters": n, "patterns": M,
            "first_call_s": t_first, "seconds": dt, "Gbp_s": n / dt / 1e9, "matches": int(c.sum()),
            "plan": L.bsgpu_last_plan().decode(),
            "bruteforce_estimate_s": M * n / 3.8e11}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="C3,C3p,C4,C5,P1,P2")
    a = ap.parse_args()
    todo = {"C3": c3, "C3p": c3prime, "C4": c4, "C5": c5, "P1": p1, "P2": p2}
    for name in a.only.split(","):
        t0 = time.perf_counter()
        res = todo[name](a.scale)
        res["wall_s"] = time.perf_counter() - t0
        print(json.dumps(res), flush=True)
