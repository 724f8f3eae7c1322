import ctypes as C, sys, time, warnings
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp, synth
L = _lib.lib()
n = 250_000_000
subj = synth.subject_region(0, n, n)
pats = synth.dictionary(1_000_000, 25, n, 1)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    pd = bs.PDict(bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(len(pats), 25, np.int32)), max_mismatch=2)
ds = mp.DeviceSubject.from_xstring(subj)
parts = pd.parts()
arr = mp._parts_array(parts)
for mode, name in ((_lib.MATCHES_AS_COUNTS, "counts"), (_lib.MATCHES_AS_ENDS, "ends")):
    for k in range(4):
        res = _lib.Result()
        torch.cuda.synchronize()
        L.bsgpu_kernel_timing(1)
        t0 = time.perf_counter()
        _lib.check(L.bsgpu_match(arr, len(parts), ds.handle, 2, 0, 1, 1, mode, C.byref(res)))
        t1 = time.perf_counter()
        out = mp._take_result(res, mode)
        t2 = time.perf_counter()
        kt = _lib.kernel_times()
        L.bsgpu_kernel_timing(0)
        print(name, k, f"C call {1e3*(t1-t0):.3f} ms, take_result {1e3*(t2-t1):.3f} ms", [(a, round(b, 3)) for a, b in kt], flush=True)
        del out
