"""Times the exact-dictionary scan engines on the headline workload (1M 25-mers x 250 Mbp, resident
letters) under a list of engine settings, in one process: one JSON line per setting with the
library's own kernel events, the matches found (must agree across settings) and the plan.

    python tools/bench_exact.py [--mbp 250] [--patterns 1000000] [--steps 20] [--only name,name]
"""
import argparse
import ctypes as C
import json
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biostrings_b200 as bs                                    # noqa: E402
from biostrings_b200 import _lib, matchpdict as mp, synth       # noqa: E402

SETTINGS = {
    "byteseed_v6": {"BSGPU_NO_XSEED": "1"},
    "xseed": {},
    "xseed_np2": {"BSGPU_XSEED_NP": "2"},
    "xseed_kbits8": {"BSGPU_XSEED_KBITS": "8"},
    "xseed_kbits8_np2": {"BSGPU_XSEED_KBITS": "8", "BSGPU_XSEED_NP": "2"},
    "xseed_unaligned": {"BSGPU_XSEED_UNALIGNED": "1"},
    "xseed_split_confirm": {"BSGPU_XSEED_SPLIT_CONFIRM": "1"},
    # candidates dropped (results wrong by construction): what the scan alone costs
    "xseed_var16_no_candidates": {"BSGPU_XSEED_VARIANT": "16"},
}
KNOBS = sorted({k for v in SETTINGS.values() for k in v})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=250.0)
    ap.add_argument("--patterns", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=25)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    L = _lib.lib()
    n = int(args.mbp * 1e6)
    M = args.patterns
    pats = synth.dictionary(M, args.width, n, 1)
    subj = synth.subject_region(0, n, n)
    ds = mp.DeviceSubject.from_xstring(subj)
    counts = torch.zeros(M, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    nl = C.c_int(0)
    names = [s for s in args.only.split(",") if s] or list(SETTINGS)
    ref_counts = None
    for name in names:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(SETTINGS[name])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pd = bs.PDict(bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(M, args.width, np.int32)))
        parts = pd.parts()
        parr = mp._parts_array(parts)

        def run(k):
            counts.zero_()
            for _ in range(k):
                _lib.check(L.bsgpu_count_device(parr, len(parts), ds.handle, 0, 0, 1, 1,
                                                C.c_void_p(counts.data_ptr()), sptr, C.byref(nl)))
        run(3)
        torch.cuda.synchronize()
        L.bsgpu_kernel_timing(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(args.steps)
        e1.record(stream)
        torch.cuda.synchronize()
        acc = {}
        for kname, ms in _lib.kernel_times():
            acc.setdefault(kname, []).append(ms)
        L.bsgpu_kernel_timing(0)
        c = counts.cpu().numpy() // args.steps
        if ref_counts is None:
            ref_counts = c
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"setting": name, "env": SETTINGS[name], "ms_per_step": ms,
                          "Gbp_s": n / ms / 1e6, "frac_hbm_6555": n / ms / 1e6 / 6555.8,
                          "kernel_ms": {k: [min(v), sum(v) / len(v)] for k, v in acc.items()},
                          "matches": int(c.sum()), "same_counts_as_first": bool((c == ref_counts).all()),
                          "plan": L.bsgpu_last_plan().decode()}), flush=True)
        del pd, parts, parr


if __name__ == "__main__":
    main()
