"""Times the phases of the host-buffer call on the GPU box (diagnostic)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biostrings_b200 as bs
from biostrings_b200 import _lib, matchpdict as mp, synth
L = _lib.lib()
n = 250_000_000
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
chunk = host.numpy(); synth.subject_region(0, n, n, out=chunk)
pats = synth.dictionary(1_000_000, 25, n, 1)
pd = bs.PDict(bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(len(pats), 25, np.int32)))
def T(f, name, reps=3):
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
        print(f"{name}: {(time.perf_counter()-t)*1e3:.2f} ms")
    return r
d = torch.empty(n, dtype=torch.uint8, device="cuda")
T(lambda: d.copy_(host, non_blocking=True), "torch H2D 250MB pinned")
ds = T(lambda: mp.DeviceSubject.from_xstring(chunk), "from_xstring")
T(lambda: mp.match_parts(pd.parts(), ds, 0, 0, (True, True), _lib.MATCHES_AS_COUNTS), "match_parts counts")
T(lambda: ds.close(), "close", 1)
T(lambda: bs.countPDict(pd, chunk), "countPDict(host)")
