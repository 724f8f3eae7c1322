"""Measurements for the SURVEY 8f rows beyond the dictionary path (one JSON line each):

  oligo   oligonucleotideFrequency(x, w) on a resident 250 Mbp chr1-shaped subject, w = 3, 6, 8, 12:
          pack_subject_kernel + oligo_freq_kernel (CUDA events recorded by the library around its
          launches), algorithmic bytes = 1 B per letter against the HBM copy peak; the reference's
          own C (src/letter_frequency.c compiled into oracle/_ref) timed on one host core over a
          20 Mbp sample beside it, and compared bit for bit on that sample.
  fasta   readDNAStringSet-style ingestion of a 250 MB FASTA text (60 letters per line) from host
          memory into a resident subject: the two kernels, the whole call, and the reference's
          read_fasta_files() on one core over a 20 MB sample.

    python tools/bench_widen.py [--mbp 250] > profiles/r01_widen.jsonl
    ncu --set full ... -k regex:"oligo_freq|fasta_" -c 4 python tools/bench_widen.py --profile --mbp 50
        (--profile: one launch each of oligo_freq_kernel w=6 / w=12 and the two FASTA kernels)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import biostrings_b200 as bs  # noqa: E402
from biostrings_b200 import _lib, synth  # noqa: E402
from biostrings_b200.letterfreq import oligo_frequency_device  # noqa: E402

HBM_GBS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6555.8) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6555.8


def kernel_ms(fn, reps):
    """mean per-call milliseconds of every library kernel launched by fn(), after one warm-up."""
    fn()
    _lib.lib().bsgpu_kernel_timing(1)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    times = _lib.kernel_times()
    wall = (time.perf_counter() - t0) / reps * 1e3
    _lib.lib().bsgpu_kernel_timing(0)
    out = {}
    for name, ms in times:
        out[name] = out.get(name, 0.0) + ms / reps
    return out, wall


def bench_oligo(x, sample):
    from oracle import ref
    L = x.size
    lines = []
    with bs.DeviceSubject.from_xstring(x) as d:
        for w in (3, 6, 8, 9, 12):
            def run():
                _lib.lib().bsgpu_subject_repack(d.handle, None)         # pack inside the measurement
                return oligo_frequency_device(d, w, 1, False, True, False)
            k, wall = kernel_ms(run, 5)
            scan = k.get("oligo_freq_kernel", float("nan"))
            pack = k.get("pack_subject_kernel", float("nan"))
            line = {"what": "oligonucleotideFrequency", "width": w, "letters": L,
                    "kernel_ms": k, "call_ms_host_api": wall,
                    "Gbp_s_kernels": L / ((scan + pack) * 1e-3) / 1e9,
                    "roofline": {"bound": "hbm", "unit": "GB/s", "algorithmic_bytes": L,
                                 "achieved": L / ((scan + pack) * 1e-3) / 1e9, "peak": HBM_GBS,
                                 "frac": L / ((scan + pack) * 1e-3) / 1e9 / HBM_GBS},
                    "plan": _lib.lib().bsgpu_last_plan().decode()}
            if ref.available():
                xs = x[:sample]
                t0 = time.perf_counter()
                r = ref.oligo_frequency_xstring(xs, w, with_labels=False)[0]
                dt = time.perf_counter() - t0
                with bs.DeviceSubject.from_xstring(xs) as ds:
                    g = oligo_frequency_device(ds, w, 1, False, True, False)
                line["cpu_baseline"] = {"value": sample / dt / 1e9, "unit": "Gbp/s", "cores": 1,
                                        "kind": "reference", "sample": f"first {sample} letters, {dt:.2f} s",
                                        "parity_with_gpu_on_sample": bool(np.array_equal(r, g))}
            lines.append(line)
    return lines


def bench_letters(x):
    """DNAString(<character>) -> resident subject: letters translated on the device against the
    host-side table lookup (numpy) followed by the upload of the codes."""
    dec = np.zeros(256, np.uint8)
    for letter, code in bs.xstring.DNA_CODES.items():
        dec[code] = ord(letter)
    letters = dec[x]

    def run():
        bs.device_subject_from_letters(letters).close()
    k, wall = kernel_ms(run, 3)
    t0 = time.perf_counter()
    codes = bs.encode_dna(letters.tobytes())
    t_host = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    bs.DeviceSubject.from_xstring(codes).close()
    t_up = (time.perf_counter() - t0) * 1e3
    return [{"what": "DNAString(<letters>) -> resident subject", "letters": int(x.size), "kernel_ms": k,
             "call_ms_host_api": wall, "host_numpy_encode_ms": t_host, "host_encoded_upload_ms": t_up,
             "roofline": {"bound": "hbm", "unit": "GB/s", "algorithmic_bytes": 2 * int(x.size),
                          "achieved": 2 * x.size / (sum(k.values()) * 1e-3) / 1e9, "peak": HBM_GBS,
                          "frac": 2 * x.size / (sum(k.values()) * 1e-3) / 1e9 / HBM_GBS}}]


def fasta_text(x, line=60):
    n = x.size
    dec = np.zeros(256, np.uint8)
    for letter, code in bs.xstring.DNA_CODES.items():
        dec[code] = ord(letter)
    full = n - n % line
    body = np.empty((full // line, line + 1), dtype=np.uint8)
    body[:, :line] = dec[x[:full]].reshape(-1, line)
    body[:, line] = ord("\n")
    tail = dec[x[full:]].tobytes() + (b"\n" if n % line else b"")
    return b">chr1 synthetic\n" + body.tobytes() + tail


def bench_fasta(x, sample):
    from oracle import ref
    L = x.size
    text = fasta_text(x)

    def run():
        dset, widths, names = bs.read_fasta_to_device(text)
        dset.close()
        return widths
    k, wall = kernel_ms(run, 3)
    dev = sum(k.values())
    line = {"what": "FASTA ingestion (text in host memory -> resident subject)", "text_bytes": len(text),
            "letters": L, "kernel_ms": k, "call_ms_host_api": wall,
            "GB_s_kernels": len(text) / (dev * 1e-3) / 1e9, "GB_s_call": len(text) / (wall * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "unit": "GB/s",
                         "algorithmic_bytes": 2 * len(text) + L,      # text read by both passes + letters written
                         "achieved": (2 * len(text) + L) / (dev * 1e-3) / 1e9, "peak": HBM_GBS,
                         "frac": (2 * len(text) + L) / (dev * 1e-3) / 1e9 / HBM_GBS},
            "plan": _lib.lib().bsgpu_last_plan().decode()}
    # the same text in pinned host memory: the upload runs at PCIe speed
    try:
        import torch
        pinned = torch.empty(len(text), dtype=torch.uint8, pin_memory=True)
        pview = pinned.numpy()
        pview[:] = np.frombuffer(text, dtype=np.uint8)

        def run_pinned():
            dset, widths, names = bs.read_fasta_to_device(pview)
            dset.close()
        _, wall_p = kernel_ms(run_pinned, 3)
        line["call_ms_pinned_text"] = wall_p
        line["GB_s_call_pinned_text"] = len(text) / (wall_p * 1e-3) / 1e9
    except Exception as e:                          # measurement only
        line["pinned_error"] = repr(e)
    if ref.available():
        ts = fasta_text(x[:sample])
        t0 = time.perf_counter()
        w, c, names, warn = ref.read_fasta(ts)
        dt = time.perf_counter() - t0
        dset, widths, _ = bs.read_fasta_to_device(ts)
        with dset:
            got = bs.download_set(dset)
        line["cpu_baseline"] = {"value": len(ts) / dt / 1e9, "unit": "GB/s of text", "cores": 1,
                                "kind": "reference", "sample": f"{len(ts)} bytes, {dt:.2f} s (two passes, as "
                                "read_fasta_files does)",
                                "parity_with_gpu_on_sample": bool(np.array_equal(c, got.concat)
                                                                  and w.tolist() == got.widths.tolist())}
    return [line]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=int, default=250)
    ap.add_argument("--sample-mbp", type=int, default=20)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    L, sample = a.mbp * 1_000_000, a.sample_mbp * 1_000_000
    x = synth.subject_region(0, L, L)
    if a.profile:
        with bs.DeviceSubject.from_xstring(x) as d:
            for w in (6, 12):
                oligo_frequency_device(d, w, 1, False, True, False)
        dset, _, _ = bs.read_fasta_to_device(fasta_text(x))
        dset.close()
        sys.exit(0)
    for ln in bench_oligo(x, sample):
        print(json.dumps(ln), flush=True)
    for ln in bench_fasta(x, sample):
        print(json.dumps(ln), flush=True)
    for ln in bench_letters(x):
        print(json.dumps(ln), flush=True)
