"""bsgpu_count_host(): countPDict on host-resident letters with the upload cut into slices and
scanned slice by slice (the end-to-end path of bench.py).  Every engine is run with slices small
enough that a 400 kbp subject takes 6 of them, and compared with the resident-subject call and
the reference: exact dictionaries (extended-seed and byte-seed engines), Trusted Band + tail with
mismatches, MTB_PDict (q-gram engine), a non-preprocessed dictionary, fixed=FALSE (uploaded whole),
and an overlap chunk with its report range and halo."""
import warnings

import numpy as np
import pytest

import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp
from oracle import ref, rlevel as R
from tests.helpers import random_dna, rng

pytestmark = pytest.mark.gpu
BACKEND = "ref" if ref.available() else "port"


@pytest.fixture(autouse=True)
def small_slices(monkeypatch):
    monkeypatch.setenv("BSGPU_HOST_SLICE_MB", "0.0625")         # 65536 letters


def subject(g, n):
    s = random_dna(g, n)
    for st in g.integers(0, n - 300, 8):
        s[st:st + int(g.integers(1, 200))] = 15
    s[g.integers(0, n, 60)] = np.array([3, 5, 6, 9, 10, 12, 7, 16], dtype=np.uint8)[g.integers(0, 8, 60)]
    return s


def cut(g, s, w, k, spread=None):
    pats = []
    while len(pats) < k:
        i = int(g.integers(0, s.size - w))
        if spread is not None:                  # around the slice boundaries
            i = int(np.clip(spread[len(pats) % len(spread)] + g.integers(-2 * w, 2 * w), 0, s.size - w))
        p = s[i:i + w].copy()
        p[~np.isin(p, [1, 2, 4, 8])] = 1
        pats.append(p)
    return pats


def both(pats, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return bs.PDict(bs.DNAStringSet(pats), **kw), R.PDict(R.SeqSet(pats), backend=BACKEND, **kw)


BOUNDS = [65536 * k for k in range(1, 6)]


@pytest.mark.parametrize("case", ["xseed25", "byteseed20", "tb_tail_mm2", "mtb_k2", "wide40"])
def test_slices_equal_whole_and_reference(case):
    g = rng(9100 + len(case))
    s = subject(g, 400_000)
    kw, mkw = {}, {}
    if case == "xseed25":
        pats = cut(g, s, 25, 300, BOUNDS) + cut(g, s, 25, 300) + [random_dna(g, 25) for _ in range(300)]
    elif case == "byteseed20":
        pats = cut(g, s, 20, 300, BOUNDS) + cut(g, s, 20, 300)
    elif case == "wide40":
        pats = cut(g, s, 40, 300, BOUNDS) + cut(g, s, 40, 300)
    elif case == "tb_tail_mm2":
        pats = cut(g, s, 36, 300, BOUNDS) + cut(g, s, 36, 300)
        for p in pats[::2]:
            p[int(g.integers(12, 36))] = 2
        kw, mkw = {"tb_start": 1, "tb_end": 12}, {"max_mismatch": 2}
    else:
        pats = cut(g, s, 25, 300, BOUNDS) + cut(g, s, 25, 300)
        for p in pats[::2]:
            p[int(g.integers(0, 25))] = 4
        kw, mkw = {"max_mismatch": 2}, {"max_mismatch": 2}
    prod, orac = both(pats, **kw)
    got = bs.countPDict(prod, s, **mkw)                      # host letters: bsgpu_count_host, 7 slices
    with mp.DeviceSubject.from_xstring(s) as ds:
        whole = bs.countPDict(prod, ds, **mkw)               # resident subject: bsgpu_match
    assert got.tolist() == whole.tolist()
    assert got.tolist() == R.countPDict(orac, s, **mkw).tolist() and got.sum() > 500
    assert bs.whichPDict(prod, s, **mkw).tolist() == R.whichPDict(orac, s, **mkw).tolist()


def test_plain_dictionary_and_fixed_false():
    g = rng(9200)
    s = subject(g, 300_000)
    pats = cut(g, s, 22, 40, BOUNDS[:4]) + [np.array([1, 15, 2, 2, 8, 4, 4, 1], dtype=np.uint8)]
    pset, oset = bs.DNAStringSet(pats), R.SeqSet(pats)
    for mkw in ({}, {"max_mismatch": 1, "fixed": False}):
        got = bs.countPDict(pset, s, **mkw)
        assert got.tolist() == R.countPDict(oset, s, backend=BACKEND, **mkw).tolist()
    prod, orac = both(cut(g, s, 30, 200, BOUNDS[:4]), tb_start=1, tb_width=15)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert bs.countPDict(prod, s, fixed=False).tolist() == R.countPDict(orac, s, fixed=False).tolist()


@pytest.mark.parametrize("mm", [0, 2])
def test_chunks_through_the_host_call(mm):
    """Overlap chunks of a long subject through bsgpu_count_host (report range + halo), summed over
    the chunks == the whole subject, with the counts accumulated on the device (d_counts)."""
    import torch
    g = rng(9300 + mm)
    s = subject(g, 500_000)
    pats = cut(g, s, 25, 500, [100_000, 250_000, 250_010, 400_000]) + cut(g, s, 25, 300)
    for p in pats[::3]:
        p[int(g.integers(0, 25))] = 8
    prod, orac = both(pats, **({"max_mismatch": 2} if mm else {}))
    parts = prod.parts()
    bounds = [0, 100_000, 250_000, 250_010, 400_000, s.size]
    acc = torch.zeros(len(pats), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()
    halo = 24
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        a, b = max(0, lo - halo), min(s.size, hi + (halo if mm else 0))
        mp.count_parts_host(parts, s[a:b], mm, 0, (True, True), chunk_start=a, subject_len=s.size,
                            report=(lo, hi), d_counts=acc.data_ptr(), stream=stream.cuda_stream)
    torch.cuda.synchronize()
    got = acc.cpu().numpy()
    with mp.DeviceSubject.from_xstring(s) as ds:
        want = mp.match_parts(parts, ds, mm, 0, (True, True), _lib.MATCHES_AS_COUNTS)
    assert got.tolist() == want.tolist() and got.sum() >= 250


def _peer_child(handle_bytes, pats, chunk, a, subject_len, lo, hi, q):
    """A second process on the same GPU: maps the owner's count vector and counts its chunk into it."""
    import ctypes as C
    try:
        L = _lib.lib()
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64)(*handle_bytes)
        _lib.check(L.bsgpu_shared_counts_open(buf, C.byref(ptr)))
        prod = bs.PDict(bs.DNAStringSet(pats))
        mp.count_parts_host(prod.parts(), chunk, 0, 0, (True, True), chunk_start=a, subject_len=subject_len,
                            report=(lo, hi), d_counts=ptr.value, stream=0)
        import torch
        torch.cuda.synchronize()
        _lib.check(L.bsgpu_shared_counts_close(ptr, 0))
        q.put("ok")
    except Exception as ex:             # noqa: BLE001 -- reported to the parent
        q.put(repr(ex))


def test_one_count_vector_for_two_processes():
    """bsgpu_shared_counts_*: the owner creates the vector, a second process maps it through its IPC
    handle and counts its overlap chunk into it (system-scope reductions), the owner counts the other
    chunk: the sum equals the whole subject.  (On a box with several GPUs bench.py does the same with
    one process per GPU and checks it against a 1-GPU recount.)"""
    import ctypes as C
    import multiprocessing as mpc
    import torch
    g = rng(9400)
    s = subject(g, 300_000)
    pats = cut(g, s, 25, 400, [150_000]) + cut(g, s, 25, 200)
    prod, _ = both(pats)
    parts = prod.parts()
    L = _lib.lib()
    ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
    _lib.check(L.bsgpu_shared_counts_create(len(pats), C.byref(ptr), handle))
    try:
        mid, halo = 150_000, 24
        ctx = mpc.get_context("spawn")
        q = ctx.Queue()
        child = ctx.Process(target=_peer_child, args=(bytes(handle), pats, s[mid - halo:], mid - halo, s.size, mid, s.size, q))
        child.start()
        mp.count_parts_host(parts, s[:mid], 0, 0, (True, True), chunk_start=0, subject_len=s.size,
                            report=(0, mid), d_counts=ptr.value, stream=0)
        torch.cuda.synchronize()
        msg = q.get(timeout=300)
        child.join(timeout=60)
        assert msg == "ok", msg
        # read the vector back through torch (the owner's own allocation)
        class _A:
            __cuda_array_interface__ = {"shape": (len(pats),), "typestr": "<i4", "data": (ptr.value, False),
                                        "version": 2}
        got = torch.as_tensor(_A(), device="cuda").cpu().numpy()
    finally:
        _lib.check(L.bsgpu_shared_counts_close(ptr, 1))
    with mp.DeviceSubject.from_xstring(s) as ds:
        want = mp.match_parts(parts, ds, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS)
    assert got.tolist() == want.tolist() and got.sum() >= 250
