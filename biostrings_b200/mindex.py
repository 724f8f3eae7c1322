"""ByPos_MIndex: the result of matchPDict(), kept in CSR form.

Mirror of R/MIndex-class.R:18-25,121-219 and src/MIndex_class.c:134-158.  The reference
stores `ends` as an R list of integer vectors (NULL when a pattern has no match); here the
same content is one offsets[M+1] + one ends[] array as it comes off the device, and the
list-shaped accessors are derived from it.  Whole-pattern duplicates are resolved lazily
through dups0, as in the reference.
"""
import numpy as np

from ._lib import NA_INTEGER


class ByPos_MIndex:
    def __init__(self, width0, offsets, ends, names=None, dups0=None):
        self.width0 = np.asarray(width0, dtype=np.int32)
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.ends = np.asarray(ends, dtype=np.int32)
        self.NAMES = names
        self.dups0 = dups0

    def __len__(self):
        return int(self.width0.size)

    def names(self):
        return self.NAMES

    def _row(self, i):
        """0-based pattern index -> row holding its ends (duplicates point at their low)."""
        if self.dups0 is not None:
            low = int(self.dups0.high2low[i])
            if low != NA_INTEGER:
                return low - 1
        return i

    def _ends_of(self, i):
        r = self._row(i)
        return self.ends[self.offsets[r]:self.offsets[r + 1]]

    def _as_list(self, values):
        """values (aligned with self.ends) cut into one array per pattern, None where the reference
        has NULL; one numpy split over the non-empty rows instead of a Python loop over all M
        patterns (a 4.5 M-pattern result spends seconds in such a loop; the reference's
        elementNROWS(endIndex(x)) on it is the 23-31 s of SolexaYi2-benchmark.txt:127-136)."""
        rows = self._rows()
        lo, hi = self.offsets[rows], self.offsets[rows + 1]
        out = [None] * len(self)
        nz = np.flatnonzero(hi > lo)
        if nz.size == 0:
            return out
        if self.dups0 is None:
            # rows are the patterns themselves, in CSR order: one split of the whole array
            pieces = np.split(values, hi[nz][:-1])
            pieces[0] = pieces[0][lo[nz[0]]:] if lo[nz[0]] else pieces[0]
            for i, p in zip(nz.tolist(), pieces):
                out[i] = p
            return out
        for i in nz.tolist():
            out[i] = values[lo[i]:hi[i]].copy()
        return out

    def endIndex(self):
        """list of int arrays, None where the reference has NULL
        (ByPos_MIndex_endIndex(high2low, ends, NULL), src/MIndex_class.c:134-158)."""
        return self._as_list(self.ends)

    def startIndex(self):
        """ends - width0 + 1 per pattern (src/MIndex_class.c:149-154); a duplicate gets its low
        element's starts (same width)."""
        n = np.diff(self.offsets)
        return self._as_list(self.ends - np.repeat(self.width0, n).astype(np.int32) + 1)

    def elementNROWS(self):
        """Number of matches per pattern, from the CSR offsets (no list is built)."""
        n = np.diff(self.offsets).astype(np.int32)
        if self.dups0 is not None:
            n = n[self.dups0.togroup() - 1]
        return n

    lengths = elementNROWS

    def __getitem__(self, i):
        """x[[i]] (0-based here): (start, end, width) arrays of the i-th pattern's matches
        (R/MIndex-class.R:149-162)."""
        if not 0 <= i < len(self):
            raise IndexError("subscript out of bounds")
        r = self._row(i)
        e = self.ends[self.offsets[r]:self.offsets[r + 1]]
        w = int(self.width0[i])
        return e - w + 1, e.copy(), np.full(e.size, w, dtype=np.int32)

    def _rows(self):
        """Row of the CSR holding the ends of every pattern (duplicates -> their low)."""
        rows = np.arange(len(self), dtype=np.int64)
        if self.dups0 is not None:
            h2l = np.asarray(self.dups0.high2low, dtype=np.int64)
            dup = h2l != NA_INTEGER
            rows[dup] = h2l[dup] - 1
        return rows

    def unlist(self, use_names=False):
        """start / end of all matches, pattern-major (unlist(MIndex), R/MIndex-class.R:56-79),
        computed on the CSR without building the per-pattern lists; with use_names also the
        pattern name of every match."""
        rows = self._rows()
        n = (self.offsets[rows + 1] - self.offsets[rows]).astype(np.int64)
        total = int(n.sum())
        if total == 0:
            z = np.zeros(0, dtype=np.int32)
            return (z, z, None) if use_names else (z, z)
        if self.dups0 is None:
            e = self.ends                   # the CSR is already pattern-major
        else:
            # gather the CSR rows in pattern order
            first = np.repeat(self.offsets[rows] - np.concatenate([[0], np.cumsum(n)[:-1]]), n)
            e = self.ends[first + np.arange(total, dtype=np.int64)]
        w0 = self.width0.astype(np.int32)
        if w0.size and (w0 == w0[0]).all():
            s = e - (int(w0[0]) - 1)        # one width: no per-match width vector
        else:
            s = e - np.repeat(w0, n) + 1
        if not use_names:
            return s, e
        names = None if self.NAMES is None else np.repeat(np.asarray(self.NAMES, dtype=object), n)
        return s, e, names

    def coverage(self, width=None):
        """coverage(MIndex): how many matches cover each subject position 1..width (the method
        IRanges supplies for the IntegerRanges that unlist() returns); width defaults to the
        largest end."""
        s, e = self.unlist()
        if width is None:
            width = int(e.max()) if e.size else 0
        if e.size == 0 or width <= 0:
            return np.zeros(max(width, 0), dtype=np.int32)
        if int(s.min()) >= 1 and int(e.max()) <= width:
            lo, hi = s, e + 1               # everything inside 1..width (the usual case): no clipping, no mask
        else:
            lo = np.clip(s.astype(np.int64), 1, width + 1)
            hi = np.clip(e.astype(np.int64) + 1, 1, width + 1)
            keep = hi > lo
            lo, hi = lo[keep], hi[keep]
        # +1 at every start, -1 behind every end, running sum; the difference array in int32
        d = np.bincount(lo, minlength=width + 2).astype(np.int32)
        d -= np.bincount(hi, minlength=width + 2).astype(np.int32)
        return np.cumsum(d[1:width + 1], dtype=np.int32)


def extractAllMatches(subject, mindex):
    """XStringViews of all the matches on the (unmasked) subject, pattern-major, named after
    the patterns (R/MIndex-class.R:91-103)."""
    from .xstring import DNAString, MaskedDNAString, XStringViews
    if isinstance(subject, MaskedDNAString):
        subject = subject.unmasked()
    elif not isinstance(subject, (DNAString, np.ndarray, str, bytes)):
        raise ValueError("'subject' must be an XString or MaskedXString object")
    if not isinstance(mindex, ByPos_MIndex):
        raise ValueError("'mindex' must be an MIndex object")
    s, e, names = mindex.unlist(use_names=True)
    v = XStringViews(subject, s, e - s + 1)     # unsafe.newXStringViews: no limits check
    v.names = names
    return v
