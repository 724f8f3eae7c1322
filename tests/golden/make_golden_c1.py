"""Generates tests/golden/c1_yeast_chr1.npz: BASELINE config 1 (the reference's own CPU-runnable
case) as a committed fixture, from the REFERENCE ITSELF -- its data file
data/yeastSEQCHR1.rda (gzip'd RDX2; the 230,208-letter string starts at offset 63 of the
stream) and its own C compiled into oracle/_ref, driven through oracle/rlevel.py.

    python tests/golden/make_golden_c1.py

Stored: the sequence packed 2 bits per letter (it holds A/C/G/T only), and for the seeded
dictionary of tests/test_golden_c1.py::dictionary() -- 10,000 random 25-mers + 1,000 cut from
the subject, PDict(tb.start=1, tb.end=12, algorithm="Twobit") -- the reference's countPDict,
matchPDict ends (CSR) and countPDict(max.mismatch=2); plus the reference's
oligonucleotideFrequency(yeast1, 4), (4, step=3) and (6) of man/nucleotideFrequency.Rd:188-203.
"""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, rlevel as R                     # noqa: E402
from tests.test_golden_c1 import dictionary, pack2      # noqa: E402

YEAST = "/root/reference/data/yeastSEQCHR1.rda"


def main():
    assert ref.available(), "oracle/_ref is not built (needs /root/reference)"
    raw = gzip.open(YEAST).read()
    seq = raw[63:63 + 230_208].decode("ascii")
    assert set(seq) <= set("ACGT") and seq.count("A") == 69_830 and seq.count("C") == 44_643
    s = R.encode(seq)
    pats = dictionary(s)
    pd = R.PDict(R.SeqSet(pats), tb_start=1, tb_end=12, algorithm="Twobit", backend="ref")
    counts = R.countPDict(pd, s)
    ends = R.matchPDict(pd, s).endIndex()
    offsets = np.zeros(len(pats) + 1, dtype=np.int64)
    np.cumsum([0 if e is None else len(e) for e in ends], out=offsets[1:])
    flat = np.concatenate([e for e in ends if e is not None]).astype(np.int32)
    counts_mm2 = R.countPDict(pd, s, max_mismatch=2)
    out = dict(subject_2bit=pack2(s), n=np.int64(len(s)), counts=counts.astype(np.int32),
               ends_offsets=offsets, ends=flat, counts_mm2=counts_mm2.astype(np.int32),
               oligo4=ref.oligo_frequency_xstring(s, 4, with_labels=False)[0],
               oligo4_step3=ref.oligo_frequency_xstring(s, 4, 3, with_labels=False)[0],
               oligo6=ref.oligo_frequency_xstring(s, 6, with_labels=False)[0])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_yeast_chr1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", int(counts.sum()), "exact matches,",
          int(counts_mm2.sum()), "with max.mismatch=2")


if __name__ == "__main__":
    main()
