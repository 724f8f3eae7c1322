import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp, synth
L = _lib.lib()
n = 250_000_000
subj = synth.subject_region(0, n, n)
pats = synth.dictionary(1_000_000, 25, n, 1)
pd = bs.PDict(bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(len(pats), 25, np.int32)))
ds = mp.DeviceSubject.from_xstring(subj)
parts = pd.parts()
arr = mp._parts_array(parts)
for mode, name in ((_lib.MATCHES_AS_COUNTS, "counts"), (_lib.MATCHES_AS_ENDS, "ends")):
    for k in range(4):
        res = _lib.Result()
        t0 = time.perf_counter()
        _lib.check(L.bsgpu_match(arr, len(parts), ds.handle, 0, 0, 1, 1, mode, C.byref(res)))
        t1 = time.perf_counter()
        out = mp._take_result(res, mode)
        t2 = time.perf_counter()
        print(name, k, f"C call {1e3*(t1-t0):.3f} ms, take_result {1e3*(t2-t1):.3f} ms")
