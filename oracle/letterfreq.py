"""numpy restatement of the reference's oligonucleotideFrequency arithmetic
(TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this).

Follows src/letter_frequency.c:183-224 (update_int_oligo_freqs: the three step regimes),
src/utils.c:101-146 (_new_TwobitEncodingBuffer / _shift_twobit_signature: a non-base letter
resets the signature, endianness 0 = first letter most significant, 1 = least significant),
:286-301 (normalize_oligo_freqs) and :634-738 (the two .Call entry points: vector, column-major
matrix, list, collapsed).  Pinned against the reference's own C (oracle/_ref, ref.py's
oligo_frequency_*) and against the expectations of tests/testthat/test-letterFrequency.R:229-292
in tests/test_oligo_oracle.py.
"""
import numpy as np

_TWOBIT = np.full(256, -1, dtype=np.int64)
for _code, _two in ((1, 0), (2, 1), (4, 2), (8, 3)):       # base_codes = c(A=1, C=2, G=4, T=8)
    _TWOBIT[_code] = _two


def oligo_counts(codes, width, step=1, invert=False):
    """Counts of one sequence: int64[4^width].  Window [s, s+width) (0-based s) is counted iff
    s % step == 0, it fits, and every letter is a base letter -- which is what all three
    branches of update_int_oligo_freqs() do (rolling with reset for step < width, one fresh
    signature per start for step >= width)."""
    if not 1 <= width <= 15:
        raise ValueError("_new_TwobitEncodingBuffer(): 'buflength' must be >= 1 and <= 15")
    if step < 1:
        raise ValueError("'step' must be a positive integer")
    codes = np.asarray(codes, dtype=np.uint8)
    ncol = 1 << (2 * width)
    n = codes.size - width + 1
    if n <= 0:
        return np.zeros(ncol, dtype=np.int64)
    two = _TWOBIT[codes]
    cs = np.concatenate([[0], np.cumsum(two >= 0)])
    ok = (cs[width:] - cs[:-width]) == width
    starts = np.flatnonzero(ok[::step]) * step
    offset = np.zeros(starts.size, dtype=np.int64)
    for i in range(width):
        weight = 4 ** i if invert else 4 ** (width - 1 - i)
        offset += two[starts + i] * weight
    return np.bincount(offset, minlength=ncol).astype(np.int64)


def _normalize(row):
    """normalize_oligo_freqs(): sum left to right in double, rows summing to 0 untouched."""
    row = row.astype(np.float64)
    total = 0.0
    for v in row.tolist():
        total += v
    return row if total == 0.0 else row / total


def xstring_oligo_frequency(codes, width, step=1, as_prob=False, invert=False):
    """XString_oligo_frequency(): int32 counts or float64 frequencies, 4^width entries."""
    c = oligo_counts(codes, width, step, invert)
    return _normalize(c) if as_prob else c.astype(np.int32)


def xstringset_oligo_frequency(concat, widths, width, step=1, as_prob=False, invert=False,
                               simplify_as="matrix"):
    """XStringSet_oligo_frequency(): "matrix" -> (N, 4^width) array (the reference's
    column-major matrix), "list" -> list of vectors, "collapsed" -> one vector."""
    concat = np.asarray(concat, dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum(np.asarray(widths, dtype=np.int64))])
    rows = [oligo_counts(concat[off[i]:off[i + 1]], width, step, invert) for i in range(len(widths))]
    ncol = 1 << (2 * width)
    if simplify_as == "collapsed":
        total = np.sum(rows, axis=0) if rows else np.zeros(ncol, dtype=np.int64)
        return _normalize(total) if as_prob else total.astype(np.int32)
    if as_prob:
        rows = [_normalize(r) for r in rows]
    else:
        rows = [r.astype(np.int32) for r in rows]
    if simplify_as == "list":
        return rows
    dtype = np.float64 if as_prob else np.int32
    return np.array(rows, dtype=dtype).reshape(len(rows), ncol)
