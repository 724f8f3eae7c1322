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

    def endIndex(self):
        """list of int arrays, None where the reference has NULL
        (ByPos_MIndex_endIndex(high2low, ends, NULL), src/MIndex_class.c:134-158)."""
        out = []
        for i in range(len(self)):
            e = self._ends_of(i)
            out.append(e.copy() if e.size else None)
        return out

    def startIndex(self):
        """ends - width0 + 1 per pattern (src/MIndex_class.c:149-154)."""
        out = []
        for i in range(len(self)):
            r = self._row(i)
            e = self.ends[self.offsets[r]:self.offsets[r + 1]]
            out.append(e + (1 - int(self.width0[r])) if e.size else None)
        return out

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

    def unlist(self):
        """start/end/width of all matches, pattern-major (unlist(MIndex))."""
        ends = self.endIndex()
        s, e = [], []
        for i, x in enumerate(ends):
            if x is not None:
                e.append(x)
                s.append(x - int(self.width0[i]) + 1)
        if not e:
            z = np.zeros(0, dtype=np.int32)
            return z, z
        return np.concatenate(s), np.concatenate(e)
