import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biostrings_b200 as bs
from biostrings_b200 import _lib, synth
from biostrings_b200.letterfreq import oligo_frequency_device
L = _lib.lib()
x = synth.subject_region(0, 250_000_000, 250_000_000)
with bs.DeviceSubject.from_xstring(x) as d:
    for env in ("", "1"):
        if env: os.environ["BSGPU_OLIGO_NO_SLICES"] = "1"
        for w in (8, 9):
            oligo_frequency_device(d, w, 1, False, True, False)
            L.bsgpu_kernel_timing(1)
            for _ in range(3):
                oligo_frequency_device(d, w, 1, False, True, False)
            torch.cuda.synchronize()
            print(env or "slices", w, [(a, round(b, 3)) for a, b in _lib.kernel_times()][-1], L.bsgpu_last_plan().decode())
            L.bsgpu_kernel_timing(0)
