"""PDict(): preprocessed pattern dictionaries with a device-resident Trusted-Band index.

Mirror of R/PDict-class.R for the path that matters: the S4 classes become small Python
classes, the C preprocessing (build_Twobit, src/match_pdict_Twobit.c:104; ACtree2_build,
src/match_pdict_ACtree2.c:792) becomes one call to bsgpu_pdict3parts_build() in
libbsgpu.so, which builds the 2-bit hash index on the GPU side of the C ABI.

    PDict(x, max_mismatch=None, tb_start=None, tb_end=None, tb_width=None,
          algorithm="ACtree2")                       R/PDict-class.R:556-573

returns a TB_PDict (one Trusted Band) or an MTB_PDict (max_mismatch >= 1: k+1 bands).
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib
from ._lib import NA_INTEGER
from .xstring import DNAStringSet


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- IRanges pieces the constructor relies on (un-vendored in the reference) ------------

def solve_user_sew(refwidths, start, end, width):
    """IRanges::solveUserSEW(refwidths, start, end, width) for scalar start/end/width
    (None = NA).  Negative start/end are relative to the end of each element
    (man/PDict-class.Rd:225-229).  Vectorised; returns 1-based starts and widths."""
    rw = np.asarray(refwidths, dtype=np.int64)

    def fail(rows, msg):
        raise ValueError(f"solving row {int(np.flatnonzero(rows)[0]) + 1}: {msg}")

    s = e = None
    if start is not None:
        s = np.full(rw.shape, int(start), dtype=np.int64)
        if start < 0:
            s += rw + 1
        bad = (s < 1) | (s > rw + 1)
        if bad.any():
            fail(bad, "'allow.nonnarrowing' is FALSE and the supplied start "
                      f"({start}) is < 1 or > refwidth + 1")
    if end is not None:
        e = np.full(rw.shape, int(end), dtype=np.int64)
        if end < 0:
            e += rw + 1
        bad = (e < 0) | (e > rw)
        if bad.any():
            fail(bad, "'allow.nonnarrowing' is FALSE and the supplied end "
                      f"({end}) is < 0 or > refwidth")
    if width is None:
        if s is None:
            s = np.ones(rw.shape, dtype=np.int64)
        if e is None:
            e = rw.copy()
        w = e - s + 1
        if (w < 0).any():
            fail(w < 0, "the supplied start/end lead to a negative width")
    else:
        if width < 0:
            fail(np.ones(rw.shape, bool), "negative values are not allowed in 'width'")
        if (s is None) == (e is None):
            fail(np.ones(rw.shape, bool),
                 "either the supplied start or the supplied end (but not both) must be NA "
                 "when the supplied width is not NA")
        w = np.full(rw.shape, int(width), dtype=np.int64)
        if s is None:
            s = e - w + 1
        else:
            e = s + w - 1
        bad = (s < 1) | (e > rw)
        if bad.any():
            fail(bad, "'allow.nonnarrowing' is FALSE and the supplied start/end/width "
                      "lead to a non-narrowing")
    return s, w


def threebands(x, start=None, end=None, width=None):
    """IRanges::threebands(): (left, middle, right) of every element."""
    s, w = solve_user_sew(x.widths, start, end, width)
    zero = np.zeros(len(x), dtype=np.int64)
    return (x.narrow(zero, s - 1), x.narrow(s - 1, w),
            x.narrow(s - 1 + w, x.widths.astype(np.int64) - (s - 1 + w)))


class Dups:
    """IRanges' Dups / H2LGrouping: high2low[i] = 1-based index of the first earlier
    identical element, or NA."""

    def __init__(self, high2low):
        self.high2low = np.ascontiguousarray(high2low, dtype=np.int32)
        n = self.high2low.size
        is_dup = self.high2low != NA_INTEGER
        self._rep = np.where(is_dup, self.high2low, np.arange(1, n + 1, dtype=np.int32))
        # low2high in CSR form: for every low, its highs in ascending order
        highs = np.flatnonzero(is_dup)
        order = np.argsort(self.high2low[highs], kind="stable")
        self._l2h_vals = (highs[order] + 1).astype(np.int32)
        cnt = np.bincount(self.high2low[highs] - 1, minlength=n)
        self._l2h_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(cnt, out=self._l2h_off[1:])

        # which(duplicated(dups0)) and their representatives, used by every match call
        # (R/matchPDict.R:410,441-443): computed once
        self.which_duplicated = (highs + 1).astype(np.int64)
        self.rep_of_duplicated = self._rep[highs].astype(np.int64)

    def __len__(self):
        return int(self.high2low.size)

    def duplicated(self):
        return self.high2low != NA_INTEGER

    def low2high(self, i):
        """1-based i -> ascending 1-based highs, or None."""
        a, b = self._l2h_off[i - 1], self._l2h_off[i]
        return None if a == b else self._l2h_vals[a:b]

    def togroup(self, j=None):
        return self._rep if j is None else self._rep[np.asarray(j) - 1]

    def togrouplength(self, j=None):
        size = np.bincount(self._rep, minlength=len(self) + 1)
        out = size[self._rep]
        return out if j is None else out[np.asarray(j) - 1]

    def members(self, i):
        """sort(c(i, unlist(low2high[i]))) for group representatives i (1-based)."""
        i = np.asarray(i, dtype=np.int64)
        if i.size == 0:
            return np.zeros(0, dtype=np.int32)
        lens = self._l2h_off[i] - self._l2h_off[i - 1]
        starts = np.repeat(self._l2h_off[i - 1], lens)
        within = np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens)
        highs = self._l2h_vals[starts + within]
        return np.sort(np.concatenate([i.astype(np.int32), highs]))


# ---- PDict3Parts ------------------------------------------------------------------------

class PDict3Parts:
    """head / Trusted Band / tail of a dictionary with its device index
    (R/PDict-class.R:174-233).  Owns a bsgpu_pdict3parts handle."""

    def __init__(self, head, tb, tail, algo, pp_exclude=None):
        L = _lib.lib()
        self.head, self.tb, self.tail, self.algo = head, tb, tail, algo
        M = len(tb)
        self.high2low = np.empty(M, dtype=np.int32)
        has_head = head is not None and bool((head.widths != 0).any())
        has_tail = tail is not None and bool((tail.widths != 0).any())
        excl = None if pp_exclude is None else np.ascontiguousarray(pp_exclude, dtype=np.int32)
        handle = C.c_void_p()
        _lib.check(L.bsgpu_pdict3parts_build(
            _lib.ALGO[algo], M, _ptr(tb.concat), _ptr(tb.widths),
            _ptr(head.concat) if has_head else None, _ptr(head.widths) if has_head else None,
            _ptr(tail.concat) if has_tail else None, _ptr(tail.widths) if has_tail else None,
            _ptr(excl), _ptr(self.high2low), C.byref(handle)))
        self._h = handle
        self.exclude_dups0 = pp_exclude is not None
        self.dups = Dups(self.high2low)

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("PDict3Parts has been freed")
        return self._h

    def blank_dups(self):
        """pptb@dups <- Dups(rep(NA, length)); exclude_dups0 <- TRUE  (R/PDict-class.R:224-226)."""
        _lib.check(_lib.lib().bsgpu_pdict3parts_blank_dups(self.handle))
        self.high2low = np.full(len(self.tb), NA_INTEGER, dtype=np.int32)
        self.dups = Dups(self.high2low)
        self.exclude_dups0 = True

    def __len__(self):
        return len(self.tb)

    def tb_width(self):
        return int(self.tb.widths[0]) if len(self.tb) else 0

    def head_or_none(self):          # R/PDict-class.R:184-191
        return None if (self.head.widths == 0).all() else self.head

    def tail_or_none(self):          # R/PDict-class.R:197-204
        return None if (self.tail.widths == 0).all() else self.tail

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().bsgpu_pdict3parts_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pdict3parts(x, tb_start, tb_end, tb_width, algo, pptb0):
    """.PDict3Parts(), R/PDict-class.R:210-233.  `pptb0` is the whole-pattern PDict3Parts
    built for dups0 (or None)."""
    head, tb, tail = threebands(x, tb_start, tb_end, tb_width)
    if pptb0 is None:
        return PDict3Parts(head, tb, tail, algo, None)
    use_pptb0 = algo == pptb0.algo and not head.widths.any() and not tail.widths.any()
    if use_pptb0:
        pptb0.blank_dups()
        return pptb0
    return PDict3Parts(head, tb, tail, algo, pptb0.high2low0)


# ---- PDict ------------------------------------------------------------------------------------

class PDict:
    """Common part of TB_PDict / MTB_PDict (R/PDict-class.R:246-330)."""

    def __init__(self, dict0):
        self.dict0 = dict0
        self.constant_width = dict0.is_constant_width()
        self.dups0 = None

    def __len__(self):
        return len(self.dict0)

    def width(self):
        return self.dict0.widths

    def names(self):
        return self.dict0.names

    def dups(self):                  # R/PDict-class.R:269-271
        return None if self.dups0 is None or len(self.dups0) == 0 else self.dups0

    def _need_dups(self):
        if self.dups() is None:
            raise ValueError("duplicates information not available for this object")
        return self.dups0

    def duplicated(self):
        return self._need_dups().duplicated()

    def togrouplength(self, j=None):
        return self._need_dups().togrouplength(j)

    patternFrequency = togrouplength

    def __getitem__(self, i):
        return self.dict0[i]


class TB_PDict(PDict):
    def __init__(self, x, tb_start, tb_end, tb_width, algo):
        """.TB_PDict(), R/PDict-class.R:392-407."""
        super().__init__(x)
        pptb0 = None
        if self.constant_width and x.has_only_base_letters():
            empty = DNAStringSet(concat=np.zeros(0, np.uint8), widths=np.zeros(len(x), np.int32))
            pptb0 = PDict3Parts(empty, x, empty, "ACtree2", None)
            pptb0.high2low0 = pptb0.high2low.copy()
        self.threeparts = _pdict3parts(x, tb_start, tb_end, tb_width, algo, pptb0)
        if pptb0 is not None:
            self.dups0 = Dups(pptb0.high2low0)
            if pptb0 is not self.threeparts:
                pptb0.close()        # only its high2low was needed

    def parts(self):
        return [self.threeparts]

    def head(self):
        return self.threeparts.head_or_none()

    def tb(self):
        return self.threeparts.tb

    def tb_width(self):
        return self.threeparts.tb_width()

    def tail(self):
        return self.threeparts.tail_or_none()


class MTB_PDict(PDict):
    def __init__(self, x, max_mismatch, algo):
        """.MTB_PDict(), R/PDict-class.R:447-491."""
        super().__init__(x)
        min_tbw = 3
        min_width = int(x.widths.min())
        if min_width < 2 * min_tbw:
            raise ValueError("'max.mismatch >= 1' is supported only if the width "
                             f"of dictionary 'x' is >= {2 * min_tbw}")
        ntb = max_mismatch + 1
        tbw0 = min_width // ntb
        if tbw0 < min_tbw:
            raise ValueError(f"'max.mismatch' must be <= {min_width // min_tbw - 1} "
                             "given the width of dictionary 'x'")
        all_tbw = [tbw0] * (ntb - min_width % ntb) + [tbw0 + 1] * (min_width % ntb)
        r1 = len(x) * sum(1.0 / 4 ** t for t in all_tbw)
        if r1 > 10:
            warnings.warn("given the characteristics of dictionary 'x', this value of "
                          "'max.mismatch' will\n  give poor performance when you call "
                          "matchPDict() on this MTB_PDict object\n  (it will of course "
                          "depend ultimately on the length of the subject)")
        all_headw = np.concatenate([[0], np.cumsum(all_tbw)]).tolist()
        pptb0 = None
        if self.constant_width:
            empty = DNAStringSet(concat=np.zeros(0, np.uint8), widths=np.zeros(len(x), np.int32))
            pptb0 = PDict3Parts(empty, x, empty, "ACtree2", None)
            pptb0.high2low0 = pptb0.high2low.copy()
        self.threeparts_list = [
            _pdict3parts(x, all_headw[i] + 1, all_headw[i + 1], None, algo, pptb0)
            for i in range(ntb)]
        if pptb0 is not None:
            self.dups0 = Dups(pptb0.high2low0)
            pptb0.close()

    def parts(self):
        return self.threeparts_list

    def as_list(self):
        return self.threeparts_list


def make_pdict(x, max_mismatch=None, tb_start=None, tb_end=None, tb_width=None,
               algorithm="ACtree2", skip_invalid_patterns=False):
    """.PDict(), R/PDict-class.R:510-554."""
    if not isinstance(x, DNAStringSet):
        x = DNAStringSet(x)
    if len(x) == 0:
        raise ValueError("'x' must contain at least one pattern")
    if x.names is not None:
        if any(n is None or n == "" for n in x.names):
            raise ValueError("'x' has invalid names")
        if len(set(x.names)) != len(x.names):
            raise ValueError("'x' has duplicated names")

    def single_or_na(v, what):
        if v is None:
            return None
        if isinstance(v, bool) or not isinstance(v, (int, float, np.integer, np.floating)):
            raise ValueError(f"'{what}' must be a single integer or 'NA'")
        return int(v)

    max_mismatch = single_or_na(max_mismatch, "max.mismatch")
    tb_start = single_or_na(tb_start, "tb.start")
    tb_end = single_or_na(tb_end, "tb.end")
    tb_width = single_or_na(tb_width, "tb.width")
    if not isinstance(algorithm, str):
        raise ValueError("'algorithm' must be a character vector")
    if algorithm == "ACtree":
        warnings.warn("support for ACtree preprocessing algo has been dropped, "
                      "using ACtree2 algo")
        algorithm = "ACtree2"
    if algorithm not in _lib.ALGO:
        raise ValueError(f"{algorithm}: unsupported Trusted Band type in 'pdict'")
    if skip_invalid_patterns is not False:
        raise ValueError("'skip.invalid.patterns' must be FALSE for now, sorry")
    is_default_tb = tb_start is None and tb_end is None and tb_width is None
    if max_mismatch is not None and not is_default_tb:
        raise ValueError("'tb.start', 'tb.end' and 'tb.width' must be NAs "
                         "when 'max.mismatch' is not NA")
    if max_mismatch is None or max_mismatch == 0:
        return TB_PDict(x, tb_start, tb_end, tb_width, algorithm)
    if max_mismatch < 0:
        raise ValueError("'max.mismatch' must be 'NA' or >= 0")
    return MTB_PDict(x, max_mismatch, algorithm)
