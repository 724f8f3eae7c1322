import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biostrings_b200 as bs
g = np.random.Generator(np.random.PCG64(5))
B = np.array([1, 2, 4, 8], dtype=np.uint8)
s = B[g.integers(0, 4, 5000)]
starts = np.concatenate([np.arange(200, 1200, 7), np.arange(2000, 2200, 1), np.arange(3000, 3300, 3)])
pats = [s[a:a + 25].copy() for a in starts]
pd = bs.PDict(bs.DNAStringSet(pats))
m = bs.matchPDict(pd, s)
ends = m.endIndex()
lost = []
for i, a in enumerate(starts):
    want = a + 25
    got = [] if ends[i] is None else ends[i].tolist()
    if want not in got:
        lost.append((i, int(a), int(a) % 4, got))
print("stride", os.environ.get("BSGPU_SAMPLE_STRIDE"), "lost", len(lost), "of", len(starts))
print(lost[:20])
from collections import Counter
print(Counter((a + 128) % 4 for _, a, _, _ in lost), Counter((a+128) % 32 for _, a, _, _ in lost))
