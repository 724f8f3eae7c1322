import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp, synth
L = _lib.lib()
n = 250_000_000
subj = synth.subject_region(0, n, n)
pats = synth.dictionary(1_000_000, 25, n, 1)
pd = bs.PDict(bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(len(pats), 25, np.int32)))
parts = pd.parts()
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.numpy()[:] = subj
for name, ds in (("xstring", mp.DeviceSubject.from_xstring(subj)), ("chunk", mp.DeviceSubject.from_chunk(host.numpy(), 0, n, 0, n))):
    for k in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o, e = mp.match_parts(parts, ds, 0, 0, (True, True), _lib.MATCHES_AS_ENDS)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print(name, k, f"{1e3*(t1-t0):.3f} ms", e.size, flush=True)
        del o, e
