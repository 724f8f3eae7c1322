"""`with.indels=TRUE` (the part of SURVEY 8f rank 1 the CUDA path still refuses): the Python
restatement oracle/indels.py -- split into the per-position candidate step and the sequential
best-local-match filter, the shape a device implementation needs -- against the reference's own
_match_pattern_indels (src/match_pattern_indels.c, compiled into oracle/_ref) through the
reference's entry point match_XStringSet_XString, on the hand-checkable example and on seeded
inputs (substitutions, insertions and deletions planted; all four `fixed` modes)."""
import numpy as np
import pytest

from oracle import indels as OI
from oracle import ref
from tests.helpers import random_dna, rng

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
ENC = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15}


def enc(s):
    return np.array([ENC[c] for c in s], dtype=np.uint8)


def ref_counts(pats, S, k, fixed=(True, True)):
    concat = np.concatenate(pats) if pats else np.zeros(0, np.uint8)
    widths = np.array([len(p) for p in pats], dtype=np.int32)
    return ref.match_set_xstring(concat, widths, S, k, 0, True, fixed, "indels",
                                 ref.MATCHES_AS_COUNTS).tolist()


def test_hand_checkable_example():
    """A small case that can be followed by hand (letters outside the DNA alphabet on purpose:
    fixed=TRUE compares raw bytes).  The reference keeps a driver for such cases,
    test_match_pattern_indels() at src/match_pattern_indels.c:7-20, but no expectations; the
    starts below are what its _match_pattern_indels reports (the counts are checked against it
    here, the starts against the best-local-match definition at :22-34)."""
    for p, s, k, starts in [("ABCDE", "BCDExAxBCDDxDABCxExxABDCExExAABCDEE", 0, [30]),
                            ("ABCDE", "BCDExAxBCDDxDABCxExxABDCExExAABCDEE", 1, [1, 14, 30]),
                            ("ABCDE", "BCDExAxBCDDxDABCxExxABDCExExAABCDEE", 2, [1, 8, 14, 21, 30])]:
        P = np.frombuffer(p.encode(), dtype=np.uint8)
        S = np.frombuffer(s.encode(), dtype=np.uint8)
        if k == 0:
            continue            # the indels algorithm is not entered for max.mismatch = 0
        got = [a for a, w in OI.match_pattern_indels(P, S, k)]
        assert got == starts
        assert ref_counts([P], S, k) == [len(starts)]


@pytest.mark.parametrize("k", [1, 2, 3])
def test_restatement_equals_reference_counts(k):
    g = rng(100 + k)
    S = random_dna(g, 3000)
    S[g.integers(0, 3000, 20)] = 15
    pats = []
    for _ in range(40):
        w = int(g.integers(k + 1, 25))
        a = int(g.integers(0, 3000 - 30))
        p = S[a:a + w].copy()
        for _ in range(int(g.integers(0, k + 2))):          # plant edits
            kind, at = int(g.integers(3)), int(g.integers(0, len(p)))
            if kind == 0:
                p[at] = [1, 2, 4, 8][int(g.integers(4))]
            elif kind == 1 and len(p) > k + 1:
                p = np.delete(p, at)
            else:
                p = np.insert(p, at, [1, 2, 4, 8][int(g.integers(4))])
        p[p == 15] = 1
        pats.append(p.astype(np.uint8))
    pats.append(random_dna(g, k))                           # not longer than k: naive loop
    assert OI.count_indels(pats, S, k) == ref_counts(pats, S, k)


@pytest.mark.parametrize("fixed", [(True, True), (False, False), (True, False), (False, True)])
def test_fixed_modes_and_iupac(fixed):
    g = rng(7)
    codes = np.array([1, 2, 4, 8, 1, 2, 4, 8, 15, 3, 5, 10], dtype=np.uint8)
    S = codes[g.integers(0, len(codes), 1200)]
    pats = [codes[g.integers(0, len(codes), int(g.integers(4, 14)))] for _ in range(25)]
    pats += [S[a:a + 12].copy() for a in g.integers(0, 1100, 10)]
    assert OI.count_indels(pats, S, 2, *fixed) == ref_counts(pats, S, 2, fixed)


def test_edges():
    S = enc("ACGTACGTAC")
    for p, k in [("ACGTACGTACGT", 2), ("ACGTACGTACGTA", 2), ("A", 1), ("AC", 1), ("TTTT", 3), ("ACGTAC", 5)]:
        assert OI.count_indels([enc(p)], S, k) == ref_counts([enc(p)], S, k), (p, k)
    assert OI.count_indels([enc("ACG")], enc(""), 1) == ref_counts([enc("ACG")], enc(""), 1)


# ---- the device-ready core (biostrings_b200/csrc/bsgpu_indels.cuh) compiled for the CPU ----------

@pytest.fixture(scope="module")
def hostcheck(tmp_path_factory):
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path_factory.mktemp("hostcheck") / "indels_hostcheck.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(root, "tests", "hostcheck", "indels_hostcheck.cpp")])
    L = C.CDLL(so)
    L.indels_hostcheck.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p]
    return L


def core_counts(L, pats, S, k, fixed=(True, True)):
    concat = np.ascontiguousarray(np.concatenate(pats), dtype=np.uint8)
    widths = np.array([len(p) for p in pats], dtype=np.int32)
    S = np.ascontiguousarray(S, dtype=np.uint8)
    counts = np.zeros(len(pats), dtype=np.int64)
    rc = L.indels_hostcheck(concat.ctypes.data, widths.ctypes.data, len(pats), S.ctypes.data, S.size,
                            k, int(fixed[0]), int(fixed[1]), counts.ctypes.data)
    assert rc == 0
    return counts.tolist()


@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_device_core_equals_reference(hostcheck, k):
    g = rng(500 + k)
    S = random_dna(g, 4000)
    S[g.integers(0, 4000, 30)] = 15
    pats = []
    for _ in range(60):
        w = int(g.integers(k + 1, 30))
        a = int(g.integers(0, 4000 - 40))
        p = S[a:a + w].copy()
        for _ in range(int(g.integers(0, k + 2))):
            kind, at = int(g.integers(3)), int(g.integers(0, len(p)))
            if kind == 0:
                p[at] = [1, 2, 4, 8][int(g.integers(4))]
            elif kind == 1 and len(p) > k + 1:
                p = np.delete(p, at)
            else:
                p = np.insert(p, at, [1, 2, 4, 8][int(g.integers(4))])
        pats.append(p.astype(np.uint8))
    pats += [S[:k + 1].copy(), S[-(k + 3):].copy(), np.concatenate([random_dna(g, 3), S[:10]])]   # subject edges
    assert core_counts(hostcheck, pats, S, k) == ref_counts(pats, S, k)


@pytest.mark.parametrize("fixed", [(True, True), (False, False), (True, False), (False, True)])
def test_device_core_fixed_modes(hostcheck, fixed):
    g = rng(9)
    codes = np.array([1, 2, 4, 8, 1, 2, 4, 8, 15, 3, 5, 10], dtype=np.uint8)
    S = codes[g.integers(0, len(codes), 1500)]
    pats = [codes[g.integers(0, len(codes), int(g.integers(4, 16)))] for _ in range(30)]
    pats += [S[a:a + 12].copy() for a in g.integers(0, 1400, 10)]
    assert core_counts(hostcheck, pats, S, 3, fixed) == ref_counts(pats, S, 3, fixed)


def test_device_core_short_subjects(hostcheck):
    for s, p, k in [("ACGTACGTAC", "ACGTACGTACGT", 2), ("ACGTACGTAC", "ACGTACGTACGTA", 2), ("AC", "ACG", 1),
                    ("A", "AC", 1), ("ACGT", "TTTTT", 4)]:
        assert core_counts(hostcheck, [enc(p)], enc(s), k) == ref_counts([enc(p)], enc(s), k), (s, p, k)


def test_engine_host_step_on_several_sequences(hostcheck):
    """bsgpu_count_indels() minus the device: candidates of every (letter, pattern) of a
    multi-sequence subject in the concat layout, in scrambled order, through the host step
    (sort by (pattern, position), filter restarted per sequence) -- summed and per sequence."""
    import ctypes as C
    g = rng(4711)
    S = random_dna(g, 5000)
    cuts = [(0, 1800), (1700, 2600), (2600, 2600), (4000, 5000)]     # overlapping, empty, tail
    seqs = [S[a:b] for a, b in cuts]
    k = 2
    pats = []
    for _ in range(30):
        a = int(g.integers(0, 4900))
        p = S[a:a + int(g.integers(k + 1, 25))].copy()
        if g.random() < 0.5 and len(p) > k + 2:
            p = np.delete(p, int(g.integers(0, len(p))))
        pats.append(p.astype(np.uint8))
    concat = np.ascontiguousarray(np.concatenate(pats))
    widths = np.array([len(p) for p in pats], dtype=np.int32)
    sconcat = np.ascontiguousarray(np.concatenate(seqs))
    slen = np.array([len(s) for s in seqs], dtype=np.int32)
    hostcheck.indels_engine_hostcheck.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                  C.c_int32] + [C.c_int] * 4 + [C.c_void_p]
    per = np.stack([np.array(ref_counts(pats, s, k)) for s in seqs], axis=1)
    for per_sequence in (0, 1):
        out = np.zeros(len(pats) * (len(seqs) if per_sequence else 1), dtype=np.int32)
        rc = hostcheck.indels_engine_hostcheck(concat.ctypes.data, widths.ctypes.data, len(pats),
                                               sconcat.ctypes.data, slen.ctypes.data, len(seqs), k, 1, 1,
                                               per_sequence, out.ctypes.data)
        assert rc == 0
        if per_sequence:
            assert np.array_equal(out.reshape((len(pats), len(seqs)), order="F"), per)
        else:
            assert out.tolist() == per.sum(axis=1).tolist()
