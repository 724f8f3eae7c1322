import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp, workloads as W
L = _lib.lib()
n = int(387.5e6); M = 2_000_000
subj = W.c5_genome_region(0, n)
concat, widths = W.c5_dictionary(M, max(n, 64 << 20))
pd = bs.PDict(bs.DNAStringSet(concat=concat, widths=widths), tb_start=1, tb_width=15)
ds = mp.DeviceSubject.from_xstring(subj)
parts = pd.parts()
arr = mp._parts_array(parts)
for fixed in ((1, 1), (0, 0)):
    for k in range(4):
        res = _lib.Result()
        torch.cuda.synchronize()
        L.bsgpu_kernel_timing(1)
        t0 = time.perf_counter()
        _lib.check(L.bsgpu_match(arr, len(parts), ds.handle, 0, 0, fixed[0], fixed[1], _lib.MATCHES_AS_COUNTS, C.byref(res)))
        t1 = time.perf_counter()
        out = mp._take_result(res, _lib.MATCHES_AS_COUNTS)
        t2 = time.perf_counter()
        kt = _lib.kernel_times()
        L.bsgpu_kernel_timing(0)
        print(fixed, k, f"C call {1e3*(t1-t0):.3f} ms, take_result {1e3*(t2-t1):.3f} ms", [(a, round(b, 3)) for a, b in kt], flush=True)
        del out
