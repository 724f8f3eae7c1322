import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import biostrings_b200 as bs
from biostrings_b200 import matchpdict as mp
g = np.random.default_rng(1)
B = np.array([1, 2, 4, 8], dtype=np.uint8)
unit = B[g.integers(0, 4, 7)]
n = 49_000_000
s = np.tile(unit, n // 7)
pats = [np.tile(np.roll(unit, r), 4)[:25].copy() for r in range(7)] + [B[g.integers(0, 4, 25)] for _ in range(1000)]
pd = bs.PDict(bs.DNAStringSet(pats))
ds = mp.DeviceSubject.from_xstring(s)
for k in range(3):
    t0 = time.perf_counter(); c = bs.countPDict(pd, ds); dt = time.perf_counter() - t0
    print("dense", k, f"{dt*1e3:.2f} ms", int(c[:7].sum()), s.size - 24, int(c[7:].sum()))
m = bs.matchPDict(pd, ds)
print("ends", int(m.elementNROWS().sum()))
