"""Measures the two memory rooflines that bound the PDict scan on this GPU:
streaming copy (GB/s) and random 32-byte-sector gathers (G sectors/s) for several table
sizes (L2-resident up to ~100 MB on B200, HBM beyond).  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biostrings_b200 import _lib  # noqa: E402

L = _lib.lib()
out = {}
for mb in (2, 8, 16, 32, 48, 64, 96, 128, 256, 1024):
    c, g = C.c_double(), C.c_double()
    _lib.check(L.bsgpu_probe_peaks(C.byref(c) if mb == 2 else None, C.byref(g), mb << 20))
    if mb == 2:
        out["copy_gbs"] = round(c.value, 1)
    out[f"gather_Gsectors_{mb}MB"] = round(g.value, 2)
print(json.dumps(out))
