"""matchPDict / countPDict / whichPDict / vcountPDict / vwhichPDict on the GPU.

Mirror of R/matchPDict.R (argument normalisation, subject dispatch, duplicate
bookkeeping) over the C ABI of libbsgpu.so.  What the reference does with one .Call per
Trusted-Band component followed by ByPos_MIndex_combine (R/matchPDict.R:335-377) is one
bsgpu_match() call here: the components are merged on the device.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import MATCHES_AS_COUNTS, MATCHES_AS_ENDS, MATCHES_AS_WHICH, NA_INTEGER
from .mindex import ByPos_MIndex
from .pdict import MTB_PDict, PDict, TB_PDict
from .xstring import DNAString, DNAStringSet, MaskedDNAString, XStringViews, encode_dna


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- argument normalisation (R/lowlevel-matching.R:41-112) -----------------------------------

def normarg_max_mismatch(v):
    if isinstance(v, bool) or not isinstance(v, (int, float, np.integer, np.floating)):
        raise ValueError("'max.mismatch' must be a single integer")
    v = int(v)
    if v < 0:
        raise ValueError("'max.mismatch' must be a non-negative integer")
    return v


def normarg_min_mismatch(v, max_mismatch):
    if isinstance(v, bool) or not isinstance(v, (int, float, np.integer, np.floating)):
        raise ValueError("'min.mismatch' must be a single integer")
    v = int(v)
    if v < 0:
        raise ValueError("'min.mismatch' must be a non-negative integer")
    if v > max_mismatch:
        raise ValueError("'min.mismatch' must be <= 'max.mismatch'")
    return v


def normarg_fixed(fixed):
    """-> (fixedP, fixedS).  TRUE/FALSE, a pair, "pattern", "subject", or a dict with
    keys pattern/subject (R's named logical vector)."""
    if isinstance(fixed, (bool, np.bool_)):
        return bool(fixed), bool(fixed)
    if isinstance(fixed, dict):
        if set(fixed) != {"pattern", "subject"}:
            raise ValueError("'fixed' names must be \"pattern\" and \"subject\"")
        return bool(fixed["pattern"]), bool(fixed["subject"])
    if isinstance(fixed, str):
        fixed = [fixed]
    fixed = list(fixed)
    if all(isinstance(f, (bool, np.bool_)) for f in fixed):
        if len(fixed) not in (1, 2):
            raise ValueError("when an unnamed logical vector, 'fixed' fixed must be of length 1 or 2")
        return bool(fixed[0]), bool(fixed[-1])
    if not all(isinstance(f, str) for f in fixed):
        raise ValueError("'fixed' not a logical or character vector")
    if len(set(fixed)) != len(fixed) or any(f not in ("pattern", "subject") for f in fixed):
        raise ValueError("when a character vector, 'fixed' must be a subset of "
                         "'c(\"pattern\", \"subject\")' with no duplicated")
    return "pattern" in fixed, "subject" in fixed


def recycle_numeric_arg(arg, argname, length_out):
    """recycleNumericArg() of S4Vectors as R/matchPDict.R:541-560 uses it for 'weight' (the package
    is not vendored in the reference tree; texts as in S4Vectors/R/utils.R): a numeric vector
    without NAs, neither empty nor longer than what it is recycled to; recycling to a length that
    is not a multiple of its own warns the way R's `ans[] <- x` does."""
    import warnings
    a = np.asarray(arg)
    if isinstance(arg, (bool, np.bool_)) or a.dtype.kind not in "iuf":
        raise ValueError(f"'{argname}' must be a numeric vector")
    a = np.atleast_1d(a)
    if a.dtype.kind != "f":
        a = a.astype(np.int32)
    if length_out == 0:
        if a.size > 1:
            raise ValueError(f"invalid length for '{argname}'")
    elif a.size == 0:
        raise ValueError(f"'{argname}' has no elements")
    elif a.size > length_out:
        raise ValueError(f"'{argname}' is longer than the object it applies to")
    if a.dtype.kind == "f" and np.isnan(a).any():
        raise ValueError(f"'{argname}' contains NAs")
    if a.size and length_out % a.size != 0:
        warnings.warn("number of items to replace is not a multiple of replacement length")
    return np.resize(a, length_out) if a.size else a


def check_weight_without_collapse(weight):
    """R/matchPDict.R:559-561: `if (!identical(weight, 1L)) warning(...)`."""
    import warnings
    if not (isinstance(weight, (int, np.integer)) and not isinstance(weight, bool) and int(weight) == 1):
        warnings.warn("'weight' is ignored when 'collapse=FALSE'")


def normarg_collapse(collapse):
    if collapse is False:
        return 0
    if isinstance(collapse, bool) or collapse not in (0, 1, 2):
        raise ValueError("'collapse' must be FALSE, 1 or 2")
    return int(collapse)


# ---- device subjects ---------------------------------------------------------------------------

class DeviceSubject:
    """A subject resident on the GPU (bytes + packed planes); reusable across calls."""

    def __init__(self, handle, kind, length):
        self._h, self.kind, self.length = handle, kind, length

    @classmethod
    def from_xstring(cls, s):
        codes = encode_dna(s)
        h = C.c_void_p()
        _lib.check(_lib.lib().bsgpu_subject_xstring(_ptr(codes), codes.size, C.byref(h)))
        return cls(h, "xstring", int(codes.size))

    @classmethod
    def from_views(cls, v):
        codes = v.subject.codes
        st = np.ascontiguousarray(v.start, dtype=np.int32)
        wd = np.ascontiguousarray(v.width, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(_lib.lib().bsgpu_subject_views(_ptr(codes), codes.size, _ptr(st), _ptr(wd),
                                                  st.size, C.byref(h)))
        return cls(h, "views", int(st.size))

    @classmethod
    def from_set(cls, s):
        h = C.c_void_p()
        _lib.check(_lib.lib().bsgpu_subject_set(_ptr(s.concat), _ptr(s.widths), len(s), C.byref(h)))
        return cls(h, "set", len(s))

    @classmethod
    def from_chunk(cls, chunk, chunk_start, subject_len, report_lo, report_hi):
        codes = encode_dna(chunk)
        h = C.c_void_p()
        _lib.check(_lib.lib().bsgpu_subject_chunk(_ptr(codes), codes.size, chunk_start,
                                                  subject_len, report_lo, report_hi, C.byref(h)))
        return cls(h, "chunk", int(codes.size))

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("DeviceSubject has been freed")
        return self._h

    def nletters(self):
        return int(_lib.lib().bsgpu_subject_nletters(self.handle))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().bsgpu_subject_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _parts_array(parts):
    arr = (C.c_void_p * len(parts))(*[p.handle for p in parts])
    return arr


def _as_array(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class _OwnedResult:
    """Keeps a bsgpu_result alive while numpy arrays view its (pinned, recycled) blocks: the big
    vectors of a result -- counts, CSR offsets, ends -- are handed out without a copy (a fresh 8 MB
    array costs more in page faults than the whole device side of a matchPDict call) and the
    blocks go back to the library when the last view dies."""

    def __init__(self, res):
        self.res, self.views = res, 0

    def view(self, ptr, n, dtype):
        import weakref
        if n == 0:
            return np.zeros(0, dtype=dtype)
        arr = np.ctypeslib.as_array(ptr, shape=(n,))
        assert arr.dtype == dtype
        self.views += 1
        weakref.finalize(arr, self._release).atexit = False
        return arr

    def _release(self):
        self.views -= 1
        if self.views == 0:
            _lib.lib().bsgpu_result_free(C.byref(self.res))

    def done(self):
        if self.views == 0:
            _lib.lib().bsgpu_result_free(C.byref(self.res))


def _take_result(res, mode, vmatch_collapse=None, n_subjects=None):
    own = _OwnedResult(res)
    try:
        if mode == MATCHES_AS_COUNTS:
            if res.dcounts:
                return _as_array(res.dcounts, res.n_counts, np.float64)
            return own.view(res.counts, res.n_counts, np.int32)
        if mode == MATCHES_AS_WHICH and n_subjects is None:
            return _as_array(res.which, res.n_which, np.int32)
        offsets = own.view(res.ends_offsets, res.n_rows + 1, np.int64)
        ends = own.view(res.ends, int(offsets[-1]) if res.n_rows else 0, np.int32)
        return offsets, ends
    finally:
        own.done()


def match_parts(parts, dsubj, max_mismatch, min_mismatch, fixed, mode):
    """bsgpu_match() on PDict3Parts objects and a DeviceSubject (the .Call level)."""
    res = _lib.Result()
    _lib.check(_lib.lib().bsgpu_match(_parts_array(parts), len(parts), dsubj.handle,
                                      max_mismatch, min_mismatch, int(fixed[0]), int(fixed[1]),
                                      mode, C.byref(res)))
    return _take_result(res, mode)


def count_parts_host(parts, letters, max_mismatch, min_mismatch, fixed, chunk_start=0, subject_len=None,
                     report=None, out=None, d_counts=None, stream=None):
    """bsgpu_count_host(): countPDict on letters that live in host memory (a whole subject, or
    one overlap chunk of it), the upload pipelined with the scan.  out: int32[M] host array to
    fill (allocated when missing); d_counts: device pointer of an int32[M] vector to ADD the
    counts to (then nothing is copied back unless `out` is given); stream: CUDA stream handle."""
    letters = np.ascontiguousarray(letters, dtype=np.uint8)
    n = int(letters.size)
    subject_len = n if subject_len is None else int(subject_len)
    lo, hi = (chunk_start, chunk_start + n) if report is None else report
    M = _lib.lib().bsgpu_pdict3parts_length(parts[0].handle)
    if out is None and d_counts is None:
        out = np.empty(M, dtype=np.int32)
    _lib.check(_lib.lib().bsgpu_count_host(_parts_array(parts), len(parts), _ptr(letters), n, int(chunk_start),
                                           subject_len, int(lo), int(hi), max_mismatch, min_mismatch,
                                           int(fixed[0]), int(fixed[1]), _ptr(out),
                                           None if d_counts is None else C.c_void_p(d_counts),
                                           None if stream is None else C.c_void_p(stream)))
    return out


def vmatch_parts(parts, dsubj, max_mismatch, min_mismatch, fixed, collapse, weight, mode):
    res = _lib.Result()
    w = None
    is_double = 0
    if mode == MATCHES_AS_COUNTS and collapse != 0:
        w = np.asarray(weight)
        is_double = int(w.dtype.kind == "f")
        w = np.ascontiguousarray(w, dtype=np.float64 if is_double else np.int32)
    _lib.check(_lib.lib().bsgpu_vmatch(_parts_array(parts), len(parts), dsubj.handle,
                                       max_mismatch, min_mismatch, int(fixed[0]), int(fixed[1]),
                                       collapse, is_double, _ptr(w), 0 if w is None else w.size,
                                       mode, C.byref(res)))
    return _take_result(res, mode, n_subjects=dsubj.length)


# ---- non-preprocessed dictionaries: 'pdict' is a plain DNAStringSet ------------------------------
# (R/matchPDict.R:170-191,215-236,262-285,379-398,502-511; R/match-utils.R:21-85)

ALL_ALGOS = ("auto", "naive-exact", "naive-inexact", "boyer-moore", "shift-or", "indels",
             "gregexpr", "gregexpr2")
CLONGINT_NBITS = 64


def normarg_algorithm(algorithm):
    if not isinstance(algorithm, str):
        raise ValueError("'algorithm' must be a single string")
    hits = [a for a in ALL_ALGOS if a.startswith(algorithm)]
    if algorithm in ALL_ALGOS:
        return algorithm
    if len(hits) != 1:
        raise ValueError("'arg' should be one of " + ", ".join(f"'{a}'" for a in ALL_ALGOS))
    return hits[0]


def normarg_with_indels(with_indels):
    if not isinstance(with_indels, (bool, np.bool_)):
        raise ValueError("'with.indels' must be TRUE or FALSE")
    return bool(with_indels)


def valid_algos(widths, max_mismatch, min_mismatch, with_indels, fixed):
    """.valid.algos(): the algorithms the reference would accept, best suited first.  They all
    report the same matches; the GPU path runs one kernel whatever the choice, but the choice
    is validated as the reference validates it."""
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
        if wmax <= CLONGINT_NBITS:
            algos.append("shift-or")
        algos.append("naive-exact")
    elif min_mismatch == 0 and fixed[0] == fixed[1] and wmax <= CLONGINT_NBITS:
        algos.append("shift-or")
    return algos + ["naive-inexact"]


def select_algo(algo, widths, max_mismatch, min_mismatch, with_indels, fixed):
    algos = valid_algos(widths, max_mismatch, min_mismatch, with_indels, fixed)
    if algo == "auto":
        return algos[0]
    if algo not in algos:
        raise ValueError("valid algos for your string matching problem (best suited first): " +
                         ", ".join(f'"{a}"' for a in algos))
    return algo


class PlainPatterns:
    """Device handle of a non-preprocessed dictionary (bsgpu_patterns_build)."""

    def __init__(self, dna_set):
        self._h = C.c_void_p()
        concat = np.ascontiguousarray(dna_set.concat, dtype=np.uint8)
        widths = np.ascontiguousarray(dna_set.widths, dtype=np.int32)
        _lib.check(_lib.lib().bsgpu_patterns_build(_ptr(concat), _ptr(widths), len(widths),
                                                   C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().bsgpu_pdict3parts_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _plain_patterns_of(pset):
    """The device handle of a non-preprocessed dictionary, kept on the DNAStringSet object (immutable,
    like its R counterpart): the index the library builds behind a large base-only dictionary on first
    use (plain_auto_index, csrc/bsgpu.cu) then serves every later call on the same object."""
    pat = getattr(pset, "_bsgpu_plain", None)
    if pat is None or not getattr(pat, "_h", None):
        pat = PlainPatterns(pset)
        try:
            pset._bsgpu_plain = pat
        except AttributeError:          # an object without a __dict__: one handle per call
            pass
    return pat


def _as_plain_set(pdict):
    if isinstance(pdict, DNAStringSet):
        return pdict
    if isinstance(pdict, (list, tuple)):
        return DNAStringSet(pdict)
    return None


def _check_plain_args(pset, max_mismatch, min_mismatch, with_indels, fixed, algorithm, mode):
    with_indels = normarg_with_indels(with_indels)
    if with_indels and mode not in (MATCHES_AS_WHICH, MATCHES_AS_COUNTS):
        raise ValueError("at the moment, within the matchPDict family, only countPDict(), "
                         "whichPDict(), vcountPDict() and vwhichPDict() support indels")
    algo = select_algo(normarg_algorithm(algorithm), np.asarray(pset.widths), max_mismatch,
                       min_mismatch, with_indels, fixed)
    return algo


def _count_indels(pat, dsubj, max_mismatch, fixed, per_sequence):
    """with.indels=TRUE: bsgpu_count_indels() -> int32[M] or [M, N] (R/matchPDict.R:171-191)."""
    M = _lib.lib().bsgpu_pdict3parts_length(pat.handle)
    N = dsubj.length if per_sequence else 1
    out = np.zeros(max(M * N, 1), dtype=np.int32)
    _lib.check(_lib.lib().bsgpu_count_indels(pat.handle, dsubj.handle, max_mismatch, int(fixed[0]),
                                             int(fixed[1]), int(per_sequence), _ptr(out)))
    out = out[:M * N]
    return out.reshape((M, N), order="F") if per_sequence else out


# ---- .checkUserArgsWhenTrustedBandIsFull / .checkMaxMismatch (R/matchPDict.R:118-141) --------

def _check_args_full_tb(parts, max_mismatch, fixed):
    import warnings
    for p in parts:
        if p.head_or_none() is None and p.tail_or_none() is None:
            if max_mismatch != 0:
                raise ValueError("'max.mismatch' must be 0 when there is no head and no tail")
            if not fixed[0]:
                warnings.warn("the value you specified for 'fixed' means that IUPAC\n"
                              "extended letters in the patterns should be treated as\n"
                              "ambiguities, but this has no effect because the patterns\n"
                              "don't contain such letters")


def _check_max_mismatch(max_mismatch, ntb):
    if max_mismatch > ntb - 1:
        raise ValueError("cannot use vmatchPDict()/vcountPDict()/vwhichPDict() with an "
                         "MTB_PDict object that was preprocessed to be used with "
                         f"'max.mismatch' <= {ntb - 1}")
    if max_mismatch < ntb - 1:
        print(f"WARNING: using 'max.mismatch={max_mismatch}' with an MTB_PDict object that "
              f"was preprocessed for allowing up to {ntb - 1} mismatching letter(s) per "
              "match is not optimal")


# ---- .matchPDict (R/matchPDict.R:400-453) -------------------------------------------------------

def _subject_to_device(subject):
    """XString / XStringViews / MaskedXString dispatch of the generics
    (R/matchPDict.R:612-760)."""
    if isinstance(subject, DeviceSubject):
        if subject.kind == "set":
            raise ValueError("please use vmatchPDict() when 'subject' is an XStringSet object "
                             "(multiple sequence)")
        return subject, False
    if isinstance(subject, (DNAStringSet, list, tuple)):
        raise ValueError("please use vmatchPDict() when 'subject' is an XStringSet object "
                         "(multiple sequence)")
    if isinstance(subject, MaskedDNAString):
        subject = subject.as_views()
    if isinstance(subject, XStringViews):
        return DeviceSubject.from_views(subject), True
    return DeviceSubject.from_xstring(subject), True


def _match_plain(pset, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm, mode,
                 what):
    """.match.XStringSet (R/matchPDict.R:380-398): no duplicates handling, no Trusted Band."""
    try:
        dsubj, owned = _subject_to_device(subject)
    except ValueError as e:
        raise ValueError(str(e).replace("vmatchPDict", "v" + what)) from None
    pat = None
    try:
        max_mismatch = normarg_max_mismatch(max_mismatch)
        min_mismatch = normarg_min_mismatch(min_mismatch, max_mismatch)
        fixed = normarg_fixed(fixed)
        algo = _check_plain_args(pset, max_mismatch, min_mismatch, with_indels, fixed, algorithm, mode)
        pat = _plain_patterns_of(pset)
        if algo == "indels":
            ans = _count_indels(pat, dsubj, max_mismatch, fixed, False)
            if mode == MATCHES_AS_WHICH:
                ans = (np.flatnonzero(ans) + 1).astype(np.int32)
        else:
            ans = match_parts([pat], dsubj, max_mismatch, min_mismatch, fixed, mode)
    finally:
        if pat is not None and getattr(pset, "_bsgpu_plain", None) is not pat:
            pat.close()
        if owned:
            dsubj.close()
    if mode == MATCHES_AS_ENDS:
        offsets, ends = ans
        return ByPos_MIndex(np.asarray(pset.widths, dtype=np.int32), offsets, ends,
                            names=getattr(pset, "names", None))
    return ans


def _match_pdict(pdict, subject, max_mismatch, min_mismatch, fixed, mode, what,
                 with_indels=False, algorithm="auto"):
    if not isinstance(pdict, PDict):
        pset = _as_plain_set(pdict)
        if pset is None:
            raise ValueError("'pdict' must be a PDict or XStringSet object")
        return _match_plain(pset, subject, max_mismatch, min_mismatch, with_indels, fixed,
                            algorithm, mode, what)
    host_letters = None
    if mode in (MATCHES_AS_COUNTS, MATCHES_AS_WHICH) and \
            isinstance(subject, (DNAString, str, bytes, np.ndarray)):
        # a single sequence in host memory: upload and scan pipelined (bsgpu_count_host)
        host_letters = encode_dna(subject)
        dsubj, owned = None, False
    else:
        try:
            dsubj, owned = _subject_to_device(subject)
        except ValueError as e:
            raise ValueError(str(e).replace("vmatchPDict", "v" + what)) from None
    try:
        dups0 = pdict.dups()
        max_mismatch = normarg_max_mismatch(max_mismatch)
        min_mismatch = normarg_min_mismatch(min_mismatch, max_mismatch)
        fixed = normarg_fixed(fixed)
        if normarg_with_indels(with_indels):
            raise ValueError("at the moment, matchPDict() and family only support indels "
                             "on a non-preprocessed pattern dictionary, sorry")
        if algorithm != "auto":
            import warnings
            warnings.warn("'algorithm' is ignored when 'pdict' is a PDict object")
        parts = pdict.parts()
        _check_args_full_tb(parts, max_mismatch, fixed)
        if isinstance(pdict, MTB_PDict):
            _check_max_mismatch(max_mismatch, len(parts))
        if host_letters is not None:
            ans = count_parts_host(parts, host_letters, max_mismatch, min_mismatch, fixed)
            if mode == MATCHES_AS_WHICH:
                ans = (np.flatnonzero(ans) + 1).astype(np.int32)    # sort(ids)+1, src/match_reporting.c:150-161
        else:
            ans = match_parts(parts, dsubj, max_mismatch, min_mismatch, fixed, mode)
    finally:
        if owned:
            dsubj.close()
    if mode == MATCHES_AS_ENDS:
        offsets, ends = ans
        return ByPos_MIndex(pdict.width(), offsets, ends, names=pdict.names(), dups0=dups0)
    if dups0 is None:
        return ans
    if mode == MATCHES_AS_WHICH:
        return dups0.members(ans)                     # members(dups0, ans)
    # ans[which_pp_excluded] <- ans[togroup(dups0, which_pp_excluded)]
    ans[dups0.which_duplicated - 1] = ans[dups0.rep_of_duplicated - 1]
    return ans


def matchPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
               algorithm="auto"):
    """-> ByPos_MIndex (R/matchPDict.R:612-658).  `pdict`: a PDict, or a plain DNAStringSet
    (non-preprocessed dictionary: any letters, any widths, mismatches anywhere)."""
    return _match_pdict(pdict, subject, max_mismatch, min_mismatch, fixed, MATCHES_AS_ENDS,
                        "matchPDict", with_indels, algorithm)


def countPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
               algorithm="auto"):
    """-> int32[length(pdict)] (R/matchPDict.R:664-709)."""
    return _match_pdict(pdict, subject, max_mismatch, min_mismatch, fixed, MATCHES_AS_COUNTS,
                        "countPDict", with_indels, algorithm)


def whichPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
               algorithm="auto"):
    """-> ascending 1-based indices of the patterns with a match (R/matchPDict.R:716-761)."""
    return _match_pdict(pdict, subject, max_mismatch, min_mismatch, fixed, MATCHES_AS_WHICH,
                        "whichPDict", with_indels, algorithm)


# ---- .vmatchPDict (R/matchPDict.R:514-605) ---------------------------------------------------------

def _vsubject_to_device(subject, what):
    if isinstance(subject, DeviceSubject):
        if subject.kind != "set":
            raise ValueError(f"please use {what}() when 'subject' is an XString or "
                             "MaskedXString object (single sequence)")
        return subject, False
    if isinstance(subject, XStringViews):
        subject = subject.to_set()
    elif isinstance(subject, (list, tuple)):
        subject = DNAStringSet(subject)
    if not isinstance(subject, DNAStringSet):
        raise ValueError(f"please use {what}() when 'subject' is an XString or "
                         "MaskedXString object (single sequence)")
    return DeviceSubject.from_set(subject), True


def _vmatch_plain(pset, subject, max_mismatch, min_mismatch, with_indels, fixed, algorithm,
                  collapse, weight, mode, what):
    """.vmatch.XStringSet (R/matchPDict.R:502-511) behind .vmatchPDict (:514-605)."""
    dsubj, owned = _vsubject_to_device(subject, what)
    pat = None
    try:
        N, M = dsubj.length, len(pset)
        max_mismatch = normarg_max_mismatch(max_mismatch)
        min_mismatch = normarg_min_mismatch(min_mismatch, max_mismatch)
        fixed = normarg_fixed(fixed)
        if mode == MATCHES_AS_COUNTS:
            collapse = normarg_collapse(collapse)
            if collapse == 1:
                weight = recycle_numeric_arg(weight, "weight", N)
            elif collapse == 2:
                weight = recycle_numeric_arg(weight, "weight", M)
            else:
                check_weight_without_collapse(weight)
                weight = np.ones(1, dtype=np.int32)
        else:
            collapse = 0
        algo = _check_plain_args(pset, max_mismatch, min_mismatch, with_indels, fixed, algorithm, mode)
        pat = _plain_patterns_of(pset)
        if algo == "indels":
            mat = _count_indels(pat, dsubj, max_mismatch, fixed, True)
            if mode == MATCHES_AS_WHICH:
                return [(np.flatnonzero(mat[:, j]) + 1).astype(np.int32) for j in range(N)]
            if collapse == 0:
                return mat
            # src/match_pdict.c:85-127: subjects outer, patterns inner
            ans = np.zeros(M if collapse == 1 else N, dtype=np.float64 if weight.dtype.kind == "f" else np.int32)
            for j in range(N):
                if collapse == 1:
                    ans += mat[:, j] * weight[j]
                else:
                    ans[j] = sum((mat[i, j] * weight[i] for i in range(M)), ans.dtype.type(0))
            return ans
        ans = vmatch_parts([pat], dsubj, max_mismatch, min_mismatch, fixed, collapse, weight, mode)
    finally:
        if pat is not None and getattr(pset, "_bsgpu_plain", None) is not pat:
            pat.close()
        if owned:
            dsubj.close()
    if mode == MATCHES_AS_WHICH:
        offsets, ids = ans
        return [ids[offsets[j]:offsets[j + 1]] for j in range(N)]
    if collapse == 0:
        ans = ans.reshape((M, N), order="F")
    return ans


def _vmatch_pdict(pdict, subject, max_mismatch, min_mismatch, fixed, collapse, weight, mode, what,
                  with_indels=False, algorithm="auto"):
    if not isinstance(pdict, PDict):
        pset = _as_plain_set(pdict)
        if pset is None:
            raise ValueError("'pdict' must be a PDict or XStringSet object")
        return _vmatch_plain(pset, subject, max_mismatch, min_mismatch, with_indels, fixed,
                             algorithm, collapse, weight, mode, what)
    if normarg_with_indels(with_indels):
        raise ValueError("at the moment, matchPDict() and family only support indels "
                         "on a non-preprocessed pattern dictionary, sorry")
    dsubj, owned = _vsubject_to_device(subject, what)
    try:
        N, M = dsubj.length, len(pdict)
        dups0 = pdict.dups()
        which_pp_excluded = None if dups0 is None else np.flatnonzero(dups0.duplicated()) + 1
        max_mismatch = normarg_max_mismatch(max_mismatch)
        min_mismatch = normarg_min_mismatch(min_mismatch, max_mismatch)
        fixed = normarg_fixed(fixed)
        if mode == MATCHES_AS_COUNTS:
            collapse = normarg_collapse(collapse)
            if collapse == 0:
                check_weight_without_collapse(weight)
                weight = np.ones(1, dtype=np.int32)
            elif collapse == 1:
                weight = recycle_numeric_arg(weight, "weight", N)
            elif collapse == 2:
                weight = recycle_numeric_arg(weight, "weight", M).copy()
                if which_pp_excluded is not None and which_pp_excluded.size:
                    # R/matchPDict.R:547-558: duplicates' weights are folded into the
                    # representative's before the C call
                    w0 = weight.copy()
                    for i in np.unique(dups0.togroup(which_pp_excluded)).tolist():
                        # weight[i] + sum(weight[low2high[[i]]]); R's sum() accumulates
                        # sequentially in long double
                        acc = np.longdouble(0) if w0.dtype.kind == "f" else 0
                        for h in dups0.low2high(i).tolist():
                            acc = acc + (np.longdouble(w0[h - 1]) if w0.dtype.kind == "f"
                                         else int(w0[h - 1]))
                        weight[i - 1] = w0[i - 1] + (np.float64(acc) if w0.dtype.kind == "f"
                                                     else acc)
        else:
            collapse = 0
        parts = pdict.parts()
        _check_args_full_tb(parts, max_mismatch, fixed)
        if isinstance(pdict, MTB_PDict):
            _check_max_mismatch(max_mismatch, len(parts))
            if mode != MATCHES_AS_WHICH:
                raise ValueError("MTB_PDict objects are not supported yet, sorry")
        ans = vmatch_parts(parts, dsubj, max_mismatch, min_mismatch, fixed, collapse, weight, mode)
    finally:
        if owned:
            dsubj.close()
    if mode == MATCHES_AS_WHICH:
        offsets, ids = ans
        out = [ids[offsets[j]:offsets[j + 1]] for j in range(N)]
        if dups0 is not None:
            out = [dups0.members(o) for o in out]     # vmembers(dups0, ans)
        return out
    if collapse == 0:
        ans = ans.reshape((M, N), order="F")
        if which_pp_excluded is not None:
            ans[which_pp_excluded - 1, :] = ans[dups0.togroup(which_pp_excluded) - 1, :]
    elif collapse == 1 and which_pp_excluded is not None:
        ans[which_pp_excluded - 1] = ans[dups0.togroup(which_pp_excluded) - 1]
    return ans


def vcountPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
                algorithm="auto", collapse=False, weight=1):
    """-> int matrix [length(pdict), length(subject)], or a collapsed vector
    (R/matchPDict.R:819-861)."""
    return _vmatch_pdict(pdict, subject, max_mismatch, min_mismatch, fixed, collapse, weight,
                         MATCHES_AS_COUNTS, "countPDict", with_indels, algorithm)


def vwhichPDict(pdict, subject, max_mismatch=0, min_mismatch=0, with_indels=False, fixed=True,
                algorithm="auto"):
    """-> list (one per subject) of ascending 1-based pattern indices
    (R/matchPDict.R:867-904)."""
    return _vmatch_pdict(pdict, subject, max_mismatch, min_mismatch, fixed, False, 1,
                         MATCHES_AS_WHICH, "whichPDict", with_indels, algorithm)


def vmatchPDict(pdict, subject, max_mismatch=0, min_mismatch=0, fixed=True):
    """R/matchPDict.R:778-809."""
    if not isinstance(subject, (DNAStringSet, XStringViews, list, tuple)):
        raise ValueError("please use matchPDict() when 'subject' is an XString or "
                         "MaskedXString object (single sequence)")
    raise NotImplementedError("vmatchPDict() is not ready yet, sorry")
