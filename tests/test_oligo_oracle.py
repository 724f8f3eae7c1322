"""oligonucleotideFrequency: the oracle pinned on CPU.

The numpy restatement (oracle/letterfreq.py) against the reference's own C compiled in place
(oracle/_ref: XString_oligo_frequency / XStringSet_oligo_frequency,
src/letter_frequency.c:634,667) on seeded inputs covering the three step regimes, both
two-bit orders, as.prob, matrix / list / collapsed, labels and the error text; and both
against what tests/testthat/test-letterFrequency.R:229-292 asserts.  That file draws its
sequence with R's RNG (set.seed(999) + sample), which cannot be replayed without R, so its
expectations are re-expressed on our own seeded sequence: equality with a direct table() of
the substrings, the codon table for step=3, one row per view, masked ranges excluded.
Also: the plane arithmetic oligo_freq_kernel uses (valid-run mask, 2w-bit field, digit
reversal) modelled in numpy, against the same references -- the kernel itself is covered by
tests/test_gpu_oligo.py.
"""
from collections import Counter

import numpy as np
import pytest

from oracle import letterfreq as LF
from oracle import ref
from tests.helpers import random_dna, rng

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
LETTERS = "ACGT"
ENC = {"A": 1, "C": 2, "G": 4, "T": 8}
DEC = {v: k for k, v in ENC.items()}


def all_strings(width, invert=False):
    out = []
    for i in range(4 ** width):
        digits = [(i // 4 ** (width - 1 - j)) % 4 for j in range(width)]
        if invert:
            digits = digits[::-1]
        out.append("".join(LETTERS[d] for d in digits))
    return out


def noisy_dna(g, n, frac=0.03):
    """bases with a sprinkle of N / IUPAC / gap letters (all reset the signature)."""
    x = random_dna(g, n)
    k = max(1, int(n * frac)) if n else 0
    if k:
        x[g.integers(0, n, k)] = g.choice(np.array([15, 3, 5, 10, 16], dtype=np.uint8), k)
    return x


def table(s, width, step=1):
    """table(substring(...)) over the windows made of base letters only."""
    c = Counter(s[i:i + width] for i in range(0, len(s) - width + 1, step))
    return [c.get(w, 0) for w in all_strings(width)]


# ---- restatement vs the reference's own C ------------------------------------------------------

@needs_ref
@pytest.mark.parametrize("width,step", [(1, 1), (2, 1), (3, 1), (3, 2), (3, 3), (4, 7), (5, 1),
                                        (6, 4), (8, 1), (11, 1), (12, 5), (13, 20)])
@pytest.mark.parametrize("invert", [False, True])
def test_xstring_restatement_equals_reference(width, step, invert):
    g = rng(1000 + 31 * width + step)
    x = noisy_dna(g, 5000)
    side = "left" if invert else "right"
    got, dim, labels = ref.oligo_frequency_xstring(x, width, step, fast_moving_side=side,
                                                   with_labels=width <= 6)
    want = LF.xstring_oligo_frequency(x, width, step, invert=invert)
    assert got.dtype == np.int32 and np.array_equal(got, want)
    if width <= 6:
        assert labels[0] == all_strings(width, invert)
    p, _, _ = ref.oligo_frequency_xstring(x, width, step, as_prob=True, fast_moving_side=side,
                                          with_labels=False)
    assert np.array_equal(p, LF.xstring_oligo_frequency(x, width, step, as_prob=True, invert=invert))


@needs_ref
@pytest.mark.parametrize("simplify_as", ["matrix", "list", "collapsed"])
@pytest.mark.parametrize("as_prob", [False, True])
def test_set_restatement_equals_reference(simplify_as, as_prob):
    g = rng(77)
    widths = g.integers(0, 60, 40).astype(np.int32)
    widths[[3, 17]] = 0                                     # empty elements
    concat = noisy_dna(g, int(widths.sum()))
    for width, step in [(1, 1), (2, 1), (3, 3), (4, 2), (7, 1)]:
        got, dim, labels = ref.oligo_frequency_set(concat, widths, width, step, as_prob=as_prob,
                                                   simplify_as=simplify_as)
        want = LF.xstringset_oligo_frequency(concat, widths, width, step, as_prob=as_prob,
                                             simplify_as=simplify_as)
        if simplify_as == "matrix":
            assert dim == (len(widths), 4 ** width)
            assert labels[0] is None and labels[1] == all_strings(width)
            assert np.array_equal(got.reshape(dim, order="F"), want)
        elif simplify_as == "list":
            assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
        else:
            assert np.array_equal(got, want) and labels[0] == all_strings(width)


@needs_ref
def test_reference_as_array_layout_and_errors():
    x = np.array([ENC[c] for c in "ACGTACGTTACGTT"], dtype=np.uint8)
    flat, dim, labels = ref.oligo_frequency_xstring(x, 3, as_array=True, fast_moving_side="left")
    assert dim == (4, 4, 4) and labels == [list(LETTERS)] * 3
    arr = flat.reshape(dim, order="F")
    assert arr[0, 1, 2] == 3                                # "ACG": first letter = first index
    assert np.array_equal(flat, LF.xstring_oligo_frequency(x, 3, invert=True))
    with pytest.raises(ref.RefError, match="'buflength' must be >= 1 and <= 15"):
        ref.oligo_frequency_xstring(x, 16)
    with pytest.raises(ValueError, match="'buflength' must be >= 1 and <= 15"):
        LF.oligo_counts(x, 16)


# ---- tests/testthat/test-letterFrequency.R:229-292, re-expressed ------------------------------

BACKENDS = ["port"] + (["ref"] if ref.available() else [])


def freq(backend, x, width, step=1):
    if backend == "ref":
        return ref.oligo_frequency_xstring(x, width, step)[0].tolist()
    return LF.xstring_oligo_frequency(x, width, step).tolist()


@pytest.mark.parametrize("backend", BACKENDS)
def test_reference_expectations(backend):
    g = rng(999)
    ln = 102
    d = random_dna(g, ln)
    ds = "".join(DEC[int(c)] for c in d)
    # :236-241 width 1 == alphabetFrequency(baseOnly=TRUE)[1:4]
    assert freq(backend, d, 1) == [ds.count(c) for c in LETTERS]
    # :243-258 di- and trinucleotides == table of the substrings
    assert freq(backend, d, 2) == table(ds, 2)
    assert freq(backend, d, 3) == table(ds, 3)
    # :260-266 step=3 == table(codons(d))
    assert freq(backend, d, 3, step=3) == table(ds, 3, step=3)
    # :268-274 on Views (codons): one row per view, a single 1 in the column of its codon
    widths = np.full(ln // 3, 3, dtype=np.int32)
    if backend == "ref":
        flat, dim, labels = ref.oligo_frequency_set(d, widths, 3)
        mat = flat.reshape(dim, order="F")
    else:
        mat = LF.xstringset_oligo_frequency(d, widths, 3)
    names = all_strings(3)
    exp = np.zeros((ln // 3, 64), dtype=np.int32)
    for i in range(ln // 3):
        exp[i, names.index(ds[3 * i:3 * i + 3])] = 1
    assert np.array_equal(mat, exp)
    # :276-286 masked ranges [10,19] and [50,57]: windows touching '#' are not counted
    masked = ds[:9] + "#" * 10 + ds[19:49] + "#" * 8 + ds[57:]
    want = Counter(masked[i:i + 2] for i in range(ln - 1))
    want = [want.get(w, 0) for w in all_strings(2)]
    gaps_w = np.array([9, 30, ln - 57], dtype=np.int32)
    gaps = np.concatenate([d[:9], d[19:49], d[57:]])
    if backend == "ref":
        got = ref.oligo_frequency_set(gaps, gaps_w, 2, simplify_as="collapsed")[0].tolist()
    else:
        got = LF.xstringset_oligo_frequency(gaps, gaps_w, 2, simplify_as="collapsed").tolist()
    assert got == want


# ---- the kernel's plane arithmetic, modelled ---------------------------------------------------

M32 = 0xFFFFFFFF


def _brev(x):
    return int(f"{x & M32:032b}"[::-1], 2)


def _rev16(x):
    r = _brev(x)
    return ((r & 0x55555555) << 1 | (r >> 1) & 0x55555555) & M32


def _shf_r(lo, hi, s):                    # __funnelshift_r (shift & 31) / _rc (clamped to 32)
    return ((hi << 32 | lo) >> s) & M32


def model_fast_word(lo, hi, starts, width, invert, ncol):
    """The unrolled path of oligo_freq_kernel for one word: the table offsets it bumps."""
    if invert:
        q0, q1, q2 = lo & M32, lo >> 32, hi & M32
    else:
        r0, r1, r2 = _rev16(hi & M32), _rev16(lo >> 32), _rev16(lo & M32)
        s0 = 2 * (17 - width)
        q0, q1, q2 = _shf_r(r0, r1, min(s0, 32)), _shf_r(r1, r2, min(s0, 32)), _shf_r(r2, 0, min(s0, 32))
        starts = _brev(starts)
    out = []
    for i in range(32):
        if not (starts >> i) & 1:
            continue
        f = _shf_r(q0, q1, 2 * (i & 15)) if i < 16 else _shf_r(q1, q2, 2 * (i & 15))
        out.append(f & (ncol - 1))
    return out


def model_plane_counts(segs, width, step, invert, fast=False):
    """What oligo_freq_kernel computes, word by word, on the concat layout
    [128 guard][seg][sep]...: returns rows (one per segment); fast = its unrolled path
    (step 1, one row: everything lands in row 0)."""
    guard = 128
    pos, starts_ = guard, []
    for s in segs:
        starts_.append(pos)
        pos += len(s) + 1
    total = ((pos + guard + 127) // 128) * 128
    buf = np.full(total + 64, 0x80, dtype=np.uint8)
    for st, s in zip(starts_, segs):
        buf[st:st + len(s)] = s
    two = LF._TWOBIT[buf[:total]]
    nwords = total // 32
    codes = [0] * (nwords + 4)
    valid = [0] * (nwords + 4)
    for j in range(nwords):
        for i in range(32):
            t = int(two[32 * j + i])
            if t >= 0:
                codes[j] |= t << (2 * i)
                valid[j] |= 1 << i
    ncol = 1 << (2 * width)
    rows = np.zeros((len(segs), ncol), dtype=np.int64)
    seg_start = np.array(starts_, dtype=np.int64)
    for j in range(nwords):
        vv = valid[j] | (valid[j + 1] << 32)
        run = vv
        for k in range(1, width):
            run &= vv >> k
        st = run & 0xFFFFFFFF
        lo, hi = codes[j], codes[j + 1]
        if fast:
            a, b = 128 - 32 * j - width + 1, pos - 32 * j - width + 1      # scan_lo, scan_hi
            if a > 0:
                st = 0 if a >= 32 else st & (M32 << a) & M32
            if b < 32:
                st = 0 if b <= 0 else st & (M32 >> (32 - b))
            for field in model_fast_word(lo, hi, st, width, invert, ncol):
                rows[0, field] += 1
            continue
        for i in range(32):
            if not (st >> i) & 1:
                continue
            p = 32 * j + i
            seg = int(np.searchsorted(seg_start, p, side="right")) - 1
            if (p - int(seg_start[seg])) % step:
                continue
            f = lo if i == 0 else ((lo >> (2 * i)) | (hi << (64 - 2 * i))) & ((1 << 64) - 1)
            field = f & (ncol - 1)
            if not invert:
                r = int(f"{field:032b}"[::-1], 2) >> (32 - 2 * width)          # __brev
                field = ((r & 0x55555555) << 1) | ((r >> 1) & 0x55555555)
            rows[seg, field] += 1
    return rows


@pytest.mark.parametrize("width,step,invert", [(1, 1, False), (2, 3, False), (3, 1, True), (5, 2, False),
                                               (9, 1, False), (11, 1, False), (11, 4, True)])
def test_plane_arithmetic_of_the_kernel(width, step, invert):
    g = rng(4242 + width)
    segs = [noisy_dna(g, int(n)) for n in (0, 1, 31, 32, 33, 97, 200, 0, 64)]
    rows = model_plane_counts(segs, width, step, invert)
    for s, row in zip(segs, rows):
        assert np.array_equal(row, LF.oligo_counts(s, width, step, invert))


@pytest.mark.parametrize("width", list(range(1, 16)))
@pytest.mark.parametrize("invert", [False, True])
def test_unrolled_path_of_the_kernel(width, invert):
    """field extraction of the step = 1 / one-row path: digit reversal of the three words, the
    pre-shift by 2 * (17 - width), compile-time offsets -- for every width."""
    g = rng(777 + width)
    segs = [noisy_dna(g, int(n)) for n in (0, 1, 47, 32, 33, 97, 300, 0, 64)]
    if width <= 10:
        rows = model_plane_counts(segs, width, 1, invert, fast=True)
        want = sum(LF.oligo_counts(s, width, 1, invert) for s in segs)
        assert np.array_equal(rows[0], want)
        return
    # wide windows: compare the offsets themselves instead of 4^width-entry tables
    s = random_dna(g, 96)
    two = LF._TWOBIT[s]
    lo = sum(int(two[i]) << (2 * i) for i in range(32))
    hi = sum(int(two[32 + i]) << (2 * i) for i in range(32))
    got = model_fast_word(lo, hi, M32, width, invert, 1 << (2 * width))
    want = []
    for i in range(32):
        digits = [int(two[i + k]) for k in range(width)]
        want.append(sum(d * 4 ** (k if invert else width - 1 - k) for k, d in enumerate(digits)))
    assert sorted(got) == sorted(want)
