"""BASELINE config 1 -- "countPDict + matchPDict of 10k random 25-mer probes (Twobit PDict,
exact) against yeastSEQCHR1" -- as a committed fixture generated from the reference's own data
file and its own C (tests/golden/make_golden_c1.py): (a) our C restatement on the CPU, (b) the
CUDA path through the C ABI on the GPU, with both of the reference's algorithms; plus the
oligonucleotideFrequency(yeast1, ...) calls of man/nucleotideFrequency.Rd:188-203."""
import os

import numpy as np
import pytest

from tests.helpers import random_dna, rng

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "c1_yeast_chr1.npz")
BASES = np.array([1, 2, 4, 8], dtype=np.uint8)


def pack2(codes):
    """A/C/G/T codes (1, 2, 4, 8) -> 2 bits per letter, 4 letters per byte."""
    two = (codes >> 1) - (codes >> 3)
    pad = (-len(two)) % 4
    two = np.concatenate([two, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return (two[:, 0] | (two[:, 1] << 2) | (two[:, 2] << 4) | (two[:, 3] << 6)).astype(np.uint8)


def unpack2(packed, n):
    two = np.stack([(packed >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)[:n]
    return BASES[two]


def dictionary(s):
    """10,000 random 25-mers + 1,000 cut from the subject (seed 1)."""
    g = rng(1)
    return [random_dna(g, 25) for _ in range(10_000)] + \
           [s[i:i + 25].copy() for i in g.integers(0, len(s) - 25, 1000)]


@pytest.fixture(scope="module")
def gold():
    z = np.load(PATH)
    s = unpack2(z["subject_2bit"], int(z["n"]))
    assert s.size == 230_208 and int((s == 1).sum()) == 69_830 and int((s == 2).sum()) == 44_643
    return z, s


def _ends(z, i):
    a, b = z["ends_offsets"][i], z["ends_offsets"][i + 1]
    return None if a == b else z["ends"][a:b].tolist()


def test_port_oracle_reproduces_config1(gold):
    from oracle import letterfreq as LF, rlevel as R
    z, s = gold
    pats = dictionary(s)
    pd = R.PDict(R.SeqSet(pats), tb_start=1, tb_end=12, algorithm="Twobit", backend="port")
    assert R.countPDict(pd, s).tolist() == z["counts"].tolist()
    assert z["counts"][10_000:].min() >= 1
    ends = R.matchPDict(pd, s).endIndex()
    for i in range(len(pats)):
        assert (None if ends[i] is None else ends[i].tolist()) == _ends(z, i)
    assert R.countPDict(pd, s, max_mismatch=2).tolist() == z["counts_mm2"].tolist()
    assert LF.xstring_oligo_frequency(s, 4).tolist() == z["oligo4"].tolist()
    assert LF.xstring_oligo_frequency(s, 4, 3).tolist() == z["oligo4_step3"].tolist()
    assert LF.xstring_oligo_frequency(s, 6).tolist() == z["oligo6"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("algorithm", ["Twobit", "ACtree2"])
def test_cuda_path_reproduces_config1(gold, algorithm):
    import biostrings_b200 as bs
    z, s = gold
    pats = dictionary(s)
    pd = bs.PDict(bs.DNAStringSet(pats), tb_start=1, tb_end=12, algorithm=algorithm)
    subject = bs.DNAString(s)
    assert bs.countPDict(pd, subject).tolist() == z["counts"].tolist()
    ends = bs.matchPDict(pd, subject).endIndex()
    for i in range(len(pats)):
        assert (None if ends[i] is None else ends[i].tolist()) == _ends(z, i)
    assert bs.countPDict(pd, subject, max_mismatch=2).tolist() == z["counts_mm2"].tolist()
    assert bs.whichPDict(pd, subject).tolist() == (np.flatnonzero(z["counts"]) + 1).tolist()


@pytest.mark.gpu
def test_cuda_oligonucleotide_frequency_of_yeast_chr1(gold):
    import biostrings_b200 as bs
    z, s = gold
    x = bs.DNAString(s)
    assert np.asarray(bs.oligonucleotideFrequency(x, 4)).tolist() == z["oligo4"].tolist()
    assert np.asarray(bs.oligonucleotideFrequency(x, 4, step=3)).tolist() == z["oligo4_step3"].tolist()
    assert np.asarray(bs.oligonucleotideFrequency(x, 6)).tolist() == z["oligo6"].tolist()
    # the reference's identity on its own data (R/letterFrequency.R:541-553)
    pd = bs.PDict(bs.mkAllStrings(bs.DNA_BASES, 6))
    assert bs.countPDict(pd, x).tolist() == z["oligo6"].tolist()
