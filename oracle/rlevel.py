"""R-level restatement of Biostrings' PDict / matchPDict family, backend-agnostic
(TEST INFRASTRUCTURE ONLY).

The reference's R layer cannot run here (no R), so the logic that sits between the user
API and the ``.Call`` boundary is restated in plain Python, one function per R function,
each citing the lines it follows.  The C level underneath is pluggable:

    backend="ref"   the reference's own C, compiled in oracle/_ref  (oracle/ref.py)
    backend="port"  our C restatement                               (oracle/port.py)

Restated here:
    DNA encoding                      R/XStringCodec-class.R:114,129-139
    threebands()/solveUserSEW()       IRanges (un-vendored; semantics from the call site
                                      R/PDict-class.R:212, man/PDict-class.Rd:225-229)
    Dups / H2LGrouping helpers        IRanges (un-vendored; call sites R/matchPDict.R:410-452)
    .PDict3Parts/.TB_PDict/.MTB_PDict/.PDict      R/PDict-class.R:210-233,392-407,447-554
    .matchPDict / .match.TB_PDict / .match.MTB_PDict   R/matchPDict.R:312-377,400-453
    .vmatchPDict                      R/matchPDict.R:461-605
    generics' subject dispatch        R/matchPDict.R:612-904
    MIndex endIndex/startIndex        R/MIndex-class.R:178-193, src/MIndex_class.c:134-158
    MTB union (sort+uniq)             src/MIndex_class.c:230-263

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import numpy as np

NA = -2147483648
MATCHES_AS_WHICH, MATCHES_AS_COUNTS, MATCHES_AS_ENDS = 1, 2, 4

# ---- encoding ---------------------------------------------------------------------

DNA_CODES = {"A": 1, "C": 2, "G": 4, "T": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10,
             "K": 12, "V": 7, "H": 11, "D": 13, "B": 14, "N": 15, "-": 16, "+": 32, ".": 64}
_ENC = np.zeros(256, dtype=np.uint8)
for _k, _v in DNA_CODES.items():
    _ENC[ord(_k)] = _v
    _ENC[ord(_k.lower())] = _v


def encode(s):
    """DNAString(s): letters -> Biostrings byte codes (lower case folded)."""
    if isinstance(s, np.ndarray):
        return np.ascontiguousarray(s, dtype=np.uint8)
    raw = np.frombuffer(s.encode("ascii"), dtype=np.uint8)
    out = _ENC[raw]
    if len(out) and out.min() == 0:
        bad = chr(raw[np.argmin(out)])
        raise ValueError(f"key {ord(bad)} (char '{bad}') not in lookup table")
    return out


class SeqSet:
    """DNAStringSet: concatenated encoded bytes + widths."""

    def __init__(self, seqs=None, concat=None, widths=None, names=None):
        if seqs is not None:
            enc = [encode(s) for s in seqs]
            self.widths = np.array([len(e) for e in enc], dtype=np.int32)
            self.concat = np.concatenate(enc) if enc else np.zeros(0, np.uint8)
        else:
            self.concat = np.ascontiguousarray(concat, dtype=np.uint8)
            self.widths = np.ascontiguousarray(widths, dtype=np.int32)
        self.offsets = np.concatenate([[0], np.cumsum(self.widths, dtype=np.int64)])
        self.names = names

    def __len__(self):
        return len(self.widths)

    def __getitem__(self, i):
        return self.concat[self.offsets[i]:self.offsets[i + 1]]

    def narrow(self, start0, width):
        """Per-element sub-sequence [start0, start0+width) (0-based starts)."""
        start0 = np.asarray(start0, dtype=np.int64)
        width = np.asarray(width, dtype=np.int64)
        parts = [self.concat[self.offsets[i] + start0[i]:self.offsets[i] + start0[i] + width[i]]
                 for i in range(len(self))]
        return SeqSet(concat=np.concatenate(parts) if parts else np.zeros(0, np.uint8),
                      widths=width.astype(np.int32))

    def pair(self):
        return (self.concat, self.widths)


class Views:
    """XStringViews on a DNAString subject."""

    def __init__(self, subject, start, width):
        self.subject = encode(subject)
        self.start = np.atleast_1d(np.asarray(start, dtype=np.int32))
        self.width = np.atleast_1d(np.asarray(width, dtype=np.int32))


class Masked:
    """MaskedDNAString: subject + masked ranges (1-based start/end, inclusive)."""

    def __init__(self, subject, mask_start, mask_end):
        self.subject = encode(subject)
        self.mask_start = np.atleast_1d(np.asarray(mask_start, dtype=np.int64))
        self.mask_end = np.atleast_1d(np.asarray(mask_end, dtype=np.int64))

    def as_views(self):
        # R/MaskedXString-class.R:244-252: views = gaps of the collapsed masks
        L = len(self.subject)
        masked = np.zeros(L + 2, dtype=bool)
        for s, e in zip(self.mask_start, self.mask_end):
            masked[max(s, 1):min(e, L) + 1] = True
        un = ~masked[1:L + 1]
        d = np.diff(np.concatenate([[0], un.astype(np.int8), [0]]))
        starts = np.flatnonzero(d == 1) + 1
        ends = np.flatnonzero(d == -1)
        return Views(self.subject, starts, ends - starts + 1)


# ---- IRanges pieces (un-vendored) --------------------------------------------------

def solve_user_SEW(refwidths, start, end, width):
    """IRanges::solveUserSEW with scalar start/end/width (None = NA), negative start/end
    relative to each element's end.  Returns (start1[], width[])."""
    out_s, out_w = [], []
    for i, rw in enumerate(np.asarray(refwidths).tolist()):
        s, e, w = start, end, width
        if s is not None:
            if s < 0:
                s += rw + 1
            if s < 1 or s > rw + 1:
                raise ValueError(f"solving row {i + 1}: 'allow.nonnarrowing' is FALSE and "
                                 f"the supplied start ({start}) is < 1 or > refwidth + 1")
        if e is not None:
            if e < 0:
                e += rw + 1
            if e < 0 or e > rw:
                raise ValueError(f"solving row {i + 1}: 'allow.nonnarrowing' is FALSE and "
                                 f"the supplied end ({end}) is < 0 or > refwidth")
        if w is None:
            if s is None:
                s = 1
            if e is None:
                e = rw
            w2 = e - s + 1
            if w2 < 0:
                raise ValueError(f"solving row {i + 1}: the supplied start/end "
                                 "lead to a negative width")
        else:
            if w < 0:
                raise ValueError(f"solving row {i + 1}: negative values are not "
                                 "allowed in 'width'")
            if (s is None) == (e is None):
                raise ValueError(f"solving row {i + 1}: either the supplied start or the "
                                 "supplied end (but not both) must be NA when the "
                                 "supplied width is not NA")
            if s is None:
                s = e - w + 1
            else:
                e = s + w - 1
            if s < 1 or e > rw:
                raise ValueError(f"solving row {i + 1}: 'allow.nonnarrowing' is FALSE and "
                                 "the supplied start/end/width lead to a non-narrowing")
            w2 = w
        out_s.append(s)
        out_w.append(w2)
    return np.array(out_s, dtype=np.int64), np.array(out_w, dtype=np.int64)


def threebands(x, start, end, width):
    """IRanges::threebands on a SeqSet -> (left, middle, right)."""
    s, w = solve_user_SEW(x.widths, start, end, width)
    left = x.narrow(np.zeros(len(x), np.int64), s - 1)
    middle = x.narrow(s - 1, w)
    right = x.narrow(s - 1 + w, x.widths - (s - 1 + w))
    return left, middle, right


class Dups:
    """IRanges Dups (H2LGrouping) over a high2low vector (NA or 1-based)."""

    def __init__(self, high2low):
        self.high2low = np.asarray(high2low, dtype=np.int32)
        n = len(self.high2low)
        self.low2high = [None] * n
        for i, h in enumerate(self.high2low.tolist()):
            if h != NA:
                if self.low2high[h - 1] is None:
                    self.low2high[h - 1] = []
                self.low2high[h - 1].append(i + 1)

    def __len__(self):
        return len(self.high2low)

    def duplicated(self):
        return self.high2low != NA

    def togroup(self, j):            # 1-based in/out
        j = np.asarray(j)
        h = self.high2low[j - 1]
        return np.where(h == NA, j, h)

    def togrouplength(self):
        rep = self.togroup(np.arange(1, len(self) + 1))
        size = np.bincount(rep, minlength=len(self) + 1)
        return size[rep]

    def members(self, i):
        out = []
        for k in np.asarray(i).tolist():
            out.append(k)
            if self.low2high[k - 1] is not None:
                out.extend(self.low2high[k - 1])
        return np.array(sorted(out), dtype=np.int32)


# ---- PDict ---------------------------------------------------------------------------

def _backend_cls(backend):
    if backend == "ref":
        from . import ref
        return ref.RefPPTB
    if backend == "port":
        from . import port
        return port.PortPPTB
    raise ValueError(backend)


class PDict3Parts:
    def __init__(self, head, pptb, tail, tb):
        self.head, self.pptb, self.tail, self.tb = head, pptb, tail, tb

    def head_or_none(self):          # R/PDict-class.R:184-191
        return None if (self.head.widths == 0).all() else self.head.pair()

    def tail_or_none(self):          # R/PDict-class.R:197-204
        return None if (self.tail.widths == 0).all() else self.tail.pair()


def _pdict3parts(x, tb_start, tb_end, tb_width, algo, pptb0, PPTB):
    """.PDict3Parts, R/PDict-class.R:210-233."""
    head, tb, tail = threebands(x, tb_start, tb_end, tb_width)
    if pptb0 is None:
        pptb = PPTB(algo, tb.concat, tb.widths, None)
    else:
        use_pptb0 = algo == pptb0.algo and (head.widths == 0).all() and \
            (tail.widths == 0).all()
        if use_pptb0:
            pptb = pptb0
            pptb.blank_dups()
            pptb.dups_blanked = True
        else:
            pptb = PPTB(algo, tb.concat, tb.widths, pptb0.high2low0)
    return PDict3Parts(head, pptb, tail, tb)


class PDict:
    """PDict(x, max.mismatch, tb.start, tb.end, tb.width, algorithm): R/PDict-class.R:510-573."""

    def __init__(self, x, max_mismatch=None, tb_start=None, tb_end=None, tb_width=None,
                 algorithm="ACtree2", backend="ref", names=None):
        PPTB = _backend_cls(backend)
        if not isinstance(x, SeqSet):
            x = SeqSet(list(x), names=names)
        if len(x) == 0:
            raise ValueError("'x' must contain at least one pattern")
        if x.names is not None:
            if any(n in ("", None) for n in x.names):
                raise ValueError("'x' has invalid names")
            if len(set(x.names)) != len(x.names):
                raise ValueError("'x' has duplicated names")
        is_default_tb = tb_start is None and tb_end is None and tb_width is None
        if max_mismatch is not None and not is_default_tb:
            raise ValueError("'tb.start', 'tb.end' and 'tb.width' must be NAs "
                             "when 'max.mismatch' is not NA")
        self.dict0 = x
        self.backend = backend
        self.constant_width = bool((x.widths == x.widths[0]).all())
        self.dups0 = None
        if max_mismatch is None or max_mismatch == 0:
            self._tb_pdict(x, tb_start, tb_end, tb_width, algorithm, PPTB)
        else:
            if max_mismatch < 0:
                raise ValueError("'max.mismatch' must be 'NA' or >= 0")
            self._mtb_pdict(x, max_mismatch, algorithm, PPTB)

    def _tb_pdict(self, x, tb_start, tb_end, tb_width, algo, PPTB):
        """.TB_PDict, R/PDict-class.R:392-407."""
        only_base = bool(np.isin(x.concat, [1, 2, 4, 8]).all())
        pptb0 = None
        if self.constant_width and only_base:
            pptb0 = PPTB("ACtree2", x.concat, x.widths, None)
            pptb0.high2low0 = pptb0.high2low.copy()
        self.threeparts_list = [_pdict3parts(x, tb_start, tb_end, tb_width, algo, pptb0, PPTB)]
        self.kind = "TB_PDict"
        if pptb0 is not None:
            self.dups0 = Dups(pptb0.high2low0)

    def _mtb_pdict(self, x, max_mismatch, algo, PPTB):
        """.MTB_PDict, R/PDict-class.R:447-491."""
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
        self.warning = None
        if len(x) * sum(1.0 / 4 ** t for t in all_tbw) > 10:
            self.warning = "poor performance"
        all_headw = np.concatenate([[0], np.cumsum(all_tbw)]).tolist()
        pptb0 = None
        if self.constant_width:
            pptb0 = PPTB("ACtree2", x.concat, x.widths, None)
            pptb0.high2low0 = pptb0.high2low.copy()
        self.threeparts_list = [
            _pdict3parts(x, all_headw[i] + 1, all_headw[i + 1], None, algo, pptb0, PPTB)
            for i in range(ntb)]
        self.kind = "MTB_PDict"
        if pptb0 is not None:
            self.dups0 = Dups(pptb0.high2low0)

    def __len__(self):
        return len(self.dict0)

    def width(self):
        return self.dict0.widths

    def dups(self):                  # R/PDict-class.R:269-271
        return None if self.dups0 is None or len(self.dups0) == 0 else self.dups0

    def duplicated(self):
        if self.dups() is None:
            raise ValueError("duplicates information not available for this object")
        return self.dups0.duplicated()

    def patternFrequency(self):
        if self.dups() is None:
            raise ValueError("duplicates information not available for this object")
        return self.dups0.togrouplength()


# ---- MIndex ----------------------------------------------------------------------------

class MIndex:
    """ByPos_MIndex: width0, ends (list of int arrays or None), dups0."""

    def __init__(self, width0, ends, names=None, dups0=None):
        self.width0, self.ends, self.names, self.dups0 = width0, ends, names, dups0

    def __len__(self):
        return len(self.ends)

    def endIndex(self):              # src/MIndex_class.c:134-158 (x_width0 = NULL)
        out = list(self.ends)
        if self.dups0 is not None:
            for i, low in enumerate(self.dups0.high2low.tolist()):
                if low != NA:
                    out[i] = out[low - 1]
        return out

    def startIndex(self):            # src/MIndex_class.c:149-154
        out = []
        for i, e in enumerate(self.endIndex()):
            if self.dups0 is not None and self.dups0.high2low[i] != NA:
                # the duplicate copies the *already shifted* low element (same width)
                low = self.dups0.high2low[i] - 1
                out.append(out[low] if out[low] is None else out[low].copy())
            else:
                out.append(None if e is None else e + (1 - int(self.width0[i])))
        return out

    def elementNROWS(self):
        return np.array([0 if e is None else len(e) for e in self.endIndex()], dtype=np.int32)

    def unlist(self, use_names=True):
        """unlist(<MIndex>), R/MIndex-class.R:56-79: start / end of every match, pattern-major,
        built from startIndex() / endIndex() exactly as the R method does; names repeated per
        match when the dictionary has names.  -> (start, end, names or None)."""
        si, ei = self.startIndex(), self.endIndex()
        s = [x for x in si if x is not None]
        e = [x for x in ei if x is not None]
        ans_start = np.concatenate(s).astype(np.int32) if s else np.zeros(0, np.int32)
        ans_end = np.concatenate(e).astype(np.int32) if e else np.zeros(0, np.int32)
        names = None
        if use_names and self.names is not None:
            names = [nm for nm, x in zip(self.names, ei) for _ in range(0 if x is None else len(x))]
        return ans_start, ans_end, names

    def coverage(self, width=None):
        """coverage(<MIndex>): the IRanges method on unlist(x) -- for every subject position
        1..width the number of matches covering it (ranges are clipped to 1..width, as
        IRanges::coverage does with its default shift=0 / given width); width defaults to the
        largest end."""
        s, e, _ = self.unlist(use_names=False)
        if width is None:
            width = int(e.max()) if e.size else 0
        cov = np.zeros(max(width, 0), dtype=np.int32)
        for a, b in zip(s.tolist(), e.tolist()):        # plain loop: this is the oracle
            a, b = max(a, 1), min(b, width)
            if b >= a:
                cov[a - 1:b] += 1
        return cov


def extractAllMatches(subject, mindex):
    """R/MIndex-class.R:91-103: views (start, width, names) of all matches on the unmasked subject;
    unsafe.newXStringViews does not check the limits."""
    s, e, names = mindex.unlist()
    return s, e - s + 1, names


def mindex_combine(ends_listlist):
    """ByPos_MIndex_combine, src/MIndex_class.c:230-263."""
    n = len(ends_listlist[0])
    out = [None] * n
    for i in range(n):
        parts = [el[i] for el in ends_listlist if el[i] is not None]
        if parts:
            allv = np.concatenate(parts)
            if len(allv):
                out[i] = np.unique(allv).astype(np.int32)
    return out


# ---- matchPDict family ---------------------------------------------------------------------

def normarg_fixed(fixed):
    """normargFixed, R/lowlevel-matching.R:71-101 -> (fixedP, fixedS)."""
    if isinstance(fixed, bool):
        return (fixed, fixed)
    if isinstance(fixed, str):
        fixed = [fixed]
    fixed = list(fixed)
    if all(isinstance(f, bool) for f in fixed):
        if len(fixed) == 1:
            return (fixed[0], fixed[0])
        if len(fixed) == 2:
            return (fixed[0], fixed[1])
        raise ValueError("when an unnamed logical vector, 'fixed' fixed must be of length 1 or 2")
    if len(set(fixed)) != len(fixed) or not all(f in ("pattern", "subject") for f in fixed):
        raise ValueError("when a character vector, 'fixed' must be a subset of "
                         "'c(\"pattern\", \"subject\")' with no duplicated")
    return ("pattern" in fixed, "subject" in fixed)


def _check_full_tb(max_mismatch, fixed):
    """.checkUserArgsWhenTrustedBandIsFull, R/matchPDict.R:118-127."""
    if max_mismatch != 0:
        raise ValueError("'max.mismatch' must be 0 when there is no head and no tail")


def _match_threeparts(tp, subject, max_mismatch, min_mismatch, fixed, mode):
    """.match.PDict3Parts.XString / .XStringViews, R/matchPDict.R:149-168,194-213."""
    head, tail = tp.head_or_none(), tp.tail_or_none()
    if head is None and tail is None:
        _check_full_tb(max_mismatch, fixed)
    if isinstance(subject, Views):
        return tp.pptb.match_views(head, tail, subject.subject, subject.start, subject.width,
                                   max_mismatch, min_mismatch, fixed, mode)
    return tp.pptb.match_xstring(head, tail, subject, max_mismatch, min_mismatch, fixed, mode)


def _check_max_mismatch(max_mismatch, ntb):
    """.checkMaxMismatch, R/matchPDict.R:129-141 (the < case only prints a WARNING)."""
    if max_mismatch > ntb - 1:
        raise ValueError("cannot use vmatchPDict()/vcountPDict()/vwhichPDict() with an "
                         "MTB_PDict object that was preprocessed to be used with "
                         f"'max.mismatch' <= {ntb - 1}")


def _matchPDict(pdict, subject, max_mismatch, min_mismatch, fixed, mode):
    """.matchPDict, R/matchPDict.R:400-453 (PDict branch)."""
    if max_mismatch < 0:
        raise ValueError("'max.mismatch' must be a non-negative integer")
    if min_mismatch < 0:
        raise ValueError("'min.mismatch' must be a non-negative integer")
    if min_mismatch > max_mismatch:
        raise ValueError("'min.mismatch' must be <= 'max.mismatch'")
    fixed = normarg_fixed(fixed)
    if isinstance(subject, Masked):
        subject = subject.as_views()
    elif not isinstance(subject, Views):
        subject = encode(subject)
    dups0 = pdict.dups()
    which_pp_excluded = None
    if dups0 is not None:
        which_pp_excluded = np.flatnonzero(dups0.duplicated()) + 1
    if pdict.kind == "TB_PDict":        # .match.TB_PDict, R/matchPDict.R:312-332
        ans = _match_threeparts(pdict.threeparts_list[0], subject, max_mismatch,
                                min_mismatch, fixed, mode)
        if mode == MATCHES_AS_ENDS:
            ans = MIndex(pdict.width(), ans, pdict.dict0.names)
    else:                               # .match.MTB_PDict, R/matchPDict.R:335-377
        ntb = len(pdict.threeparts_list)
        _check_max_mismatch(max_mismatch, ntb)
        mode2 = MATCHES_AS_ENDS if mode == MATCHES_AS_COUNTS else mode
        compons = [_match_threeparts(tp, subject, max_mismatch, min_mismatch, fixed, mode2)
                   for tp in pdict.threeparts_list]
        if mode == MATCHES_AS_WHICH:
            ans = np.unique(np.concatenate(compons)).astype(np.int32)
        else:
            ends = mindex_combine(compons)
            if mode == MATCHES_AS_COUNTS:
                ans = np.array([0 if e is None else len(e) for e in ends], dtype=np.int32)
            else:
                ans = MIndex(pdict.width(), ends, pdict.dict0.names)
    if which_pp_excluded is None:
        return ans
    if mode == MATCHES_AS_WHICH:
        return dups0.members(ans)
    if mode == MATCHES_AS_COUNTS:
        ans = ans.copy()
        ans[which_pp_excluded - 1] = ans[dups0.togroup(which_pp_excluded) - 1]
        return ans
    ans.dups0 = dups0
    return ans


# ---- non-preprocessed dictionaries: 'pdict' is a plain SeqSet ------------------------------------
# R/matchPDict.R:170-191,215-236,262-285,379-398,502-511 over src/match_pdict.c:174,267,519

def _valid_algos(widths, max_mismatch, min_mismatch, with_indels, fixed):
    """.valid.algos, R/match-utils.R:36-71."""
    if len(widths) and int(np.min(widths)) == 0:
        raise ValueError("empty patterns are not supported")
    if len(widths) and int(np.max(widths)) > 20000:
        raise ValueError("patterns with more than 20000 letters are not supported")
    if max_mismatch != 0 and with_indels:
        if min_mismatch != 0:
            raise ValueError("'min.mismatch' must be 0 when 'with.indels' is TRUE")
        return ["indels"]
    wmax = int(np.max(widths)) if len(widths) else 0
    algos = []
    if max_mismatch == 0 and all(fixed):
        algos.append("boyer-moore")
        if wmax <= 64:
            algos.append("shift-or")
        algos.append("naive-exact")
    elif min_mismatch == 0 and fixed[0] == fixed[1] and wmax <= 64:
        algos.append("shift-or")
    return algos + ["naive-inexact"]


def select_algo(algo, widths, max_mismatch, min_mismatch, with_indels, fixed):
    """selectAlgo, R/match-utils.R:73-85."""
    algos = _valid_algos(widths, max_mismatch, min_mismatch, with_indels, fixed)
    if algo == "auto":
        return algos[0]
    if algo not in algos:
        raise ValueError("valid algos for your string matching problem (best suited first): " +
                         ", ".join(f'"{a}"' for a in algos))
    return algo


def _letters_match(p, s, fixedP, fixedS):
    """src/lowlevel_matching.c:26-56 on the raw code bytes."""
    if fixedP:
        return p == s if fixedS else (p & ~s & 0xFF) == 0
    return (~p & s & 0xFF) == 0 if fixedS else (p & s) != 0


def naive_match_ends(P, S, max_nmis, min_nmis, fixedP, fixedS):
    """Independent restatement of match_naive_inexact (src/match_pattern.c:53-76) over
    _nmismatch_at_Pshift (src/lowlevel_matching.c:66-89): 1-based ends of the matches of
    pattern P in subject S, out-of-limits letters counting as mismatches.  Pure Python:
    small cases only."""
    plen, slen = len(P), len(S)
    if plen <= 0:
        raise ValueError("empty pattern")
    if max_nmis < plen - slen or min_nmis > plen:
        return []
    min_pshift = 1 - plen if plen <= max_nmis else -max_nmis
    out = []
    for pshift in range(min_pshift, slen - plen - min_pshift + 1):
        nmis = 0
        for i in range(plen):
            j = pshift + i
            if 0 <= j < slen and _letters_match(int(P[i]), int(S[j]), fixedP, fixedS):
                continue
            nmis += 1
            if nmis > max_nmis:
                break
        if min_nmis <= nmis <= max_nmis:
            out.append(pshift + plen)
    return out


def _plain_match(pset, subject, max_mismatch, min_mismatch, with_indels, fixed, algo, mode, backend):
    """One subject (encoded bytes) or Views against a plain SeqSet; returns what the .Call
    returns: counts int32[M] / which int32[] / list of ends (None for no match)."""
    if backend == "ref":
        from . import ref
        if isinstance(subject, Views):
            return ref.match_set_views(pset.concat, pset.widths, subject.subject, subject.start,
                                       subject.width, max_mismatch, min_mismatch, with_indels,
                                       fixed, algo, mode)
        return ref.match_set_xstring(pset.concat, pset.widths, subject, max_mismatch,
                                     min_mismatch, with_indels, fixed, algo, mode)
    if with_indels:
        raise NotImplementedError("indels are not restated")
    ends = []
    for i in range(len(pset)):
        if isinstance(subject, Views):
            e = []
            L = len(subject.subject)
            for st, w in zip(subject.start.tolist(), subject.width.tolist()):
                if st - 1 < 0 or st - 1 + w > L:
                    raise ValueError("'subject' has \"out of limits\" views")
                e += [x + st - 1 for x in naive_match_ends(pset[i], subject.subject[st - 1:st - 1 + w],
                                                           max_mismatch, min_mismatch, *fixed)]
        else:
            e = naive_match_ends(pset[i], subject, max_mismatch, min_mismatch, *fixed)
        ends.append(np.array(e, dtype=np.int32) if e else None)
    if mode == MATCHES_AS_ENDS:
        return ends
    counts = np.array([0 if e is None else len(e) for e in ends], dtype=np.int32)
    if mode == MATCHES_AS_COUNTS:
        return counts
    return (np.flatnonzero(counts) + 1).astype(np.int32)


def _plain_matchPDict(pset, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm,
                      mode, backend):
    """.matchPDict -> .match.XStringSet, R/matchPDict.R:400-453,380-398."""
    if min_mismatch > max_mismatch:
        raise ValueError("'min.mismatch' must be <= 'max.mismatch'")
    fixed = normarg_fixed(fixed)
    if isinstance(subject, Masked):
        subject = subject.as_views()
    elif not isinstance(subject, Views):
        subject = encode(subject)
    if with_indels and mode not in (MATCHES_AS_WHICH, MATCHES_AS_COUNTS):
        raise ValueError("at the moment, within the matchPDict family, only countPDict(), "
                         "whichPDict(), vcountPDict() and vwhichPDict() support indels")
    algo = select_algo(algorithm, pset.widths, max_mismatch, min_mismatch, with_indels, fixed)
    ans = _plain_match(pset, subject, max_mismatch, min_mismatch, with_indels, fixed, algo, mode,
                       backend)
    if mode == MATCHES_AS_ENDS:
        return MIndex(pset.widths, ans, pset.names)
    return ans


def _plain_vmatchPDict(pset, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm,
                       collapse, weight, mode, backend):
    """.vmatchPDict -> .vmatch.XStringSet, R/matchPDict.R:514-605,502-511."""
    if isinstance(subject, Views):
        subject = _views_to_set(subject)
    elif not isinstance(subject, SeqSet):
        if isinstance(subject, (list, tuple)):
            subject = SeqSet(list(subject))
        else:
            what = {MATCHES_AS_WHICH: "whichPDict", MATCHES_AS_COUNTS: "countPDict"}[mode]
            raise ValueError(f"please use {what}() when 'subject' is an XString or "
                             "MaskedXString object (single sequence)")
    if min_mismatch > max_mismatch:
        raise ValueError("'min.mismatch' must be <= 'max.mismatch'")
    fixed = normarg_fixed(fixed)
    M, N = len(pset), len(subject)
    if mode == MATCHES_AS_COUNTS:
        collapse = 0 if collapse is False else int(collapse)
        if collapse not in (0, 1, 2):
            raise ValueError("'collapse' must be FALSE, 1 or 2")
        weight = np.atleast_1d(np.asarray(weight))
        if collapse == 1:
            weight = np.resize(weight, N)
        elif collapse == 2:
            weight = np.resize(weight, M)
    else:
        collapse, weight = 0, np.array([1])
    algo = select_algo(algorithm, pset.widths, max_mismatch, min_mismatch, with_indels, fixed)
    if backend == "ref":
        from . import ref
        ans = ref.vmatch_set_set(pset.concat, pset.widths, subject.concat, subject.widths,
                                 max_mismatch, min_mismatch, with_indels, fixed, algo, collapse,
                                 weight, mode)
        if mode == MATCHES_AS_COUNTS and collapse == 0 and getattr(ans, "ndim", 0) == 1:
            ans = ans.reshape((M, N), order="F")
        if mode == MATCHES_AS_WHICH:
            ans = [np.zeros(0, np.int32) if a is None else a for a in ans]
        return ans
    cnt = np.zeros((M, N), dtype=np.int64)
    for i in range(M):
        for j in range(N):
            cnt[i, j] = len(naive_match_ends(pset[i], subject[j], max_mismatch, min_mismatch, *fixed))
    if mode == MATCHES_AS_WHICH:
        return [(np.flatnonzero(cnt[:, j]) + 1).astype(np.int32) for j in range(N)]
    if collapse == 0:
        return cnt.astype(np.int32)
    # update_vcount_collapsed_ans, src/match_pdict.c:85-127, i outer / j inner
    out = np.zeros(M if collapse == 1 else N, dtype=weight.dtype if weight.dtype.kind == "f" else np.int32)
    for i in range(M):
        for j in range(N):
            if collapse == 1:
                out[i] += cnt[i, j] * weight[j]
            else:
                out[j] += cnt[i, j] * weight[i]
    return out


def _plain_backend(pdict, backend):
    if backend is not None:
        return backend
    from . import ref
    return "ref" if ref.available() else "port"


def _is_set(subject):
    return isinstance(subject, (SeqSet, list, tuple))


def _dispatch(pdict, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm, mode,
              backend):
    if isinstance(pdict, SeqSet):
        return _plain_matchPDict(pdict, subject, max_mismatch, min_mismatch, with_indels, fixed,
                                 algorithm, mode, _plain_backend(pdict, backend))
    if with_indels:
        raise ValueError("at the moment, matchPDict() and family only support indels "
                         "on a non-preprocessed pattern dictionary, sorry")
    return _matchPDict(pdict, subject, max_mismatch, min_mismatch, fixed, mode)


def matchPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
               algorithm="auto", backend=None):
    if _is_set(subject):             # R/matchPDict.R:631-636
        raise ValueError("please use vmatchPDict() when 'subject' is an XStringSet object "
                         "(multiple sequence)")
    return _dispatch(pdict, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm,
                     MATCHES_AS_ENDS, backend)


def countPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
               algorithm="auto", backend=None):
    if _is_set(subject):
        raise ValueError("please use vcountPDict() when 'subject' is an XStringSet object "
                         "(multiple sequence)")
    return _dispatch(pdict, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm,
                     MATCHES_AS_COUNTS, backend)


def whichPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
               algorithm="auto", backend=None):
    if _is_set(subject):
        raise ValueError("please use vwhichPDict() when 'subject' is an XStringSet object "
                         "(multiple sequence)")
    return _dispatch(pdict, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm,
                     MATCHES_AS_WHICH, backend)


def _views_to_set(v):
    """as(XStringViews, "XStringSet"): out-of-limits views are trimmed silently
    (R/matchPDict.R:843,891; R/XStringViews-class.R:94-109)."""
    L = len(v.subject)
    parts = []
    for s, w in zip(v.start.tolist(), v.width.tolist()):
        a, b = max(s, 1), min(s + w - 1, L)
        parts.append(v.subject[a - 1:b] if b >= a else v.subject[:0])
    return SeqSet(concat=np.concatenate(parts) if parts else np.zeros(0, np.uint8),
                  widths=[len(p) for p in parts])


def _vmatchPDict(pdict, subject, max_mismatch, min_mismatch, fixed, collapse, weight, mode):
    """.vmatchPDict, R/matchPDict.R:514-605 (PDict branch)."""
    if isinstance(subject, Views):
        subject = _views_to_set(subject)
    elif not isinstance(subject, SeqSet):
        if isinstance(subject, (list, tuple)):
            subject = SeqSet(list(subject))
        else:
            what = {MATCHES_AS_WHICH: "whichPDict", MATCHES_AS_COUNTS: "countPDict"}[mode]
            raise ValueError(f"please use {what}() when 'subject' is an XString or "
                             "MaskedXString object (single sequence)")
    if min_mismatch > max_mismatch:
        raise ValueError("'min.mismatch' must be <= 'max.mismatch'")
    fixed = normarg_fixed(fixed)
    dups0 = pdict.dups()
    which_pp_excluded = None
    if dups0 is not None:
        which_pp_excluded = np.flatnonzero(dups0.duplicated()) + 1
    if mode == MATCHES_AS_COUNTS:
        collapse = 0 if collapse is False else int(collapse)
        if collapse not in (0, 1, 2):
            raise ValueError("'collapse' must be FALSE, 1 or 2")
        weight = np.atleast_1d(np.asarray(weight))
        if collapse == 1:
            weight = np.resize(weight, len(subject))
        elif collapse == 2:
            weight = np.resize(weight, len(pdict)).copy()
            if which_pp_excluded is not None:
                # R/matchPDict.R:547-558: fold the duplicates' weights into the rep
                w0 = weight.copy()
                for i, l2h in enumerate(dups0.low2high):
                    if l2h is not None:
                        weight[i] = w0[i] + w0[np.array(l2h) - 1].sum()
    if pdict.kind == "TB_PDict":
        tp = pdict.threeparts_list[0]
        head, tail = tp.head_or_none(), tp.tail_or_none()
        if head is None and tail is None:
            _check_full_tb(max_mismatch, fixed)
        ans = tp.pptb.vmatch_set(head, tail, subject.concat, subject.widths, max_mismatch,
                                 min_mismatch, fixed, collapse if mode == MATCHES_AS_COUNTS else 0,
                                 weight if mode == MATCHES_AS_COUNTS else np.array([1]), mode)
        if mode == MATCHES_AS_COUNTS and collapse == 0 and getattr(ans, "ndim", 0) == 1:
            ans = ans.reshape((len(pdict), len(subject)), order="F")   # allocMatrix(M, N)
    else:                               # .vmatch.MTB_PDict, R/matchPDict.R:472-504
        _check_max_mismatch(max_mismatch, len(pdict.threeparts_list))
        if mode != MATCHES_AS_WHICH:
            raise ValueError("MTB_PDict objects are not supported yet, sorry")
        compons = [tp.pptb.vmatch_set(tp.head_or_none(), tp.tail_or_none(), subject.concat,
                                      subject.widths, max_mismatch, min_mismatch, fixed, 0,
                                      np.array([1]), mode)
                   for tp in pdict.threeparts_list]
        ans = [np.unique(np.concatenate([c[j] for c in compons])).astype(np.int32)
               for j in range(len(subject))]
    if which_pp_excluded is None:
        return ans
    if mode == MATCHES_AS_WHICH:
        return [dups0.members(a) for a in ans]
    if collapse == 0:
        ans = ans.copy()
        ans[which_pp_excluded - 1, :] = ans[dups0.togroup(which_pp_excluded) - 1, :]
    elif collapse == 1:
        ans = ans.copy()
        ans[which_pp_excluded - 1] = ans[dups0.togroup(which_pp_excluded) - 1]
    return ans


def vcountPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
                algorithm="auto", collapse=False, weight=1, backend=None):
    if isinstance(pdict, SeqSet):
        return _plain_vmatchPDict(pdict, subject, max_mismatch, min_mismatch, with_indels, fixed,
                                  algorithm, collapse, weight, MATCHES_AS_COUNTS,
                                  _plain_backend(pdict, backend))
    return _vmatchPDict(pdict, subject, max_mismatch, min_mismatch, fixed, collapse,
                        weight, MATCHES_AS_COUNTS)


def vwhichPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
                algorithm="auto", backend=None):
    if isinstance(pdict, SeqSet):
        return _plain_vmatchPDict(pdict, subject, max_mismatch, min_mismatch, with_indels, fixed,
                                  algorithm, False, 1, MATCHES_AS_WHICH,
                                  _plain_backend(pdict, backend))
    return _vmatchPDict(pdict, subject, max_mismatch, min_mismatch, fixed, False, 1,
                        MATCHES_AS_WHICH)


def vmatchPDict(pdict, subject, **kw):
    if not _is_set(subject) and not isinstance(subject, Views):
        raise ValueError("please use matchPDict() when 'subject' is an XString or "
                         "MaskedXString object (single sequence)")
    raise NotImplementedError("vmatchPDict() is not ready yet, sorry")
