"""ctypes binding of oracle/_ref/libbiostrings_ref.so  (TEST INFRASTRUCTURE ONLY).

The shared object is the *reference's own* hot-path C (match_pdict*.c & friends under
/root/reference/src), compiled unmodified against the mini-R stand-in in oracle/shim/ by
oracle/Makefile.  This module exposes the reference's ``.Call`` entry points at the level
R calls them (src/R_init_Biostrings.c:27-147):

    build_Twobit / ACtree2_build            -> RefPPTB(...)
    match_PDict3Parts_XString               -> RefPPTB.match_xstring
    match_PDict3Parts_XStringViews          -> RefPPTB.match_views
    vmatch_PDict3Parts_XStringSet           -> RefPPTB.vmatch_set
    ByPos_MIndex_combine / _endIndex        -> mindex_combine / mindex_endindex
    XString_oligo_frequency / XStringSet_oligo_frequency (src/letter_frequency.c:634,667)
                                            -> oligo_frequency_xstring / oligo_frequency_set

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libbiostrings_ref.so")

NA_INTEGER = -2147483648
MATCHES_AS_WHICH, MATCHES_AS_COUNTS, MATCHES_AS_ENDS = 1, 2, 4
_INTSXP, _REALSXP, _VECSXP, _NILSXP = 13, 14, 19, 0

_SO_DISPATCH = os.path.join(_HERE, "_ref", "libbiostrings_gpu_dispatch.so")
_libs = {}
_which = "ref"          # which library lib() returns: "ref" | "dispatch"


class RefError(RuntimeError):
    """An Rf_error() raised inside the C code under the .Call boundary (message = R's text)."""


def available(which="ref"):
    return os.path.exists(_SO if which == "ref" else _SO_DISPATCH)


def use(which):
    """Select the library behind this module: "ref" = the reference's own C (default),
    "dispatch" = the product's R-facing C dispatch over libbsgpu (needs a GPU).  Both export
    the same .Call entry points and are driven through the same harness."""
    global _which
    assert which in ("ref", "dispatch")
    _which = which


def lib():
    which = _which
    if which not in _libs:
        path = _SO if which == "ref" else _SO_DISPATCH
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not built: run `make -C oracle` (the reference needs /root/reference)")
        L = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_last_warning.restype = C.c_char_p
        L.ref_pptb_build.restype = C.c_void_p
        L.ref_pptb_build.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
        L.ref_pptb_blank_dups.argtypes = [C.c_void_p]
        L.ref_pptb_free.argtypes = [C.c_void_p]
        L.ref_pptb_detach_device.argtypes = [C.c_void_p]
        L.ref_actree2_nnodes.argtypes = [C.c_void_p]
        L.ref_actree2_has_all_flinks.argtypes = [C.c_void_p]
        L.ref_actree2_compute_all_flinks.argtypes = [C.c_void_p]
        L.ref_match_xstring.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + \
            [C.c_void_p, C.c_int] + [C.c_int] * 5
        L.ref_match_views.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + \
            [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 5
        L.ref_vmatch_set.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + \
            [C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 4 + \
            [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.ref_mindex_endindex.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        if hasattr(L, "ref_match_set_xstring"):
            L.ref_match_set_xstring.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + \
                [C.c_int] * 5 + [C.c_char_p, C.c_int]
            L.ref_match_set_views.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 5 + \
                [C.c_char_p, C.c_int]
            L.ref_vmatch_set_set.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int] + [C.c_int] * 5 + \
                [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        if hasattr(L, "ref_oligo_frequency_xstring"):
            L.ref_oligo_frequency_xstring.argtypes = [C.c_void_p] + [C.c_int] * 7
            L.ref_oligo_frequency_set.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_char_p]
            L.ref_result_dim.argtypes = [C.c_void_p]
            L.ref_result_labels.restype = C.c_long
            L.ref_result_labels.argtypes = [C.c_int, C.c_char_p, C.c_long]
            L.ref_result_list_flat_double.argtypes = [C.c_void_p]
        L.ref_result_length.restype = C.c_long
        L.ref_result_list_lengths.restype = C.c_long
        L.ref_result_list_lengths.argtypes = [C.c_void_p]
        L.ref_result_copy_int.argtypes = [C.c_void_p]
        L.ref_result_copy_double.argtypes = [C.c_void_p]
        L.ref_result_list_flat.argtypes = [C.c_void_p]
        _libs[which] = L
    return _libs[which]


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc, L=None):
    L = L or lib()
    if rc != 0:
        msg = L.ref_last_error().decode("latin1")
        L.ref_release()
        raise RefError(msg)


def last_warning():
    return lib().ref_last_warning().decode("latin1")


def _fetch_result(release=True, L=None):
    """Current result -> numpy: int/double vector, or list of (int array | None)."""
    L = L or lib()
    t, n = L.ref_result_type(), L.ref_result_length()
    if t == _INTSXP:
        out = np.empty(n, dtype=np.int32)
        L.ref_result_copy_int(_ptr(out))
    elif t == _REALSXP:
        out = np.empty(n, dtype=np.float64)
        L.ref_result_copy_double(_ptr(out))
    elif t == _VECSXP:
        lens = np.empty(n, dtype=np.int32)
        total = L.ref_result_list_lengths(_ptr(lens))
        if hasattr(L, "ref_result_list_elt_type") and L.ref_result_list_elt_type() == _REALSXP:
            flat = np.empty(total, dtype=np.float64)
            L.ref_result_list_flat_double(_ptr(flat))
        else:
            flat = np.empty(total, dtype=np.int32)
            L.ref_result_list_flat(_ptr(flat))
        out, off = [], 0
        for k in lens.tolist():
            if k < 0:
                out.append(None)
            else:
                out.append(flat[off:off + k].copy())
                off += k
    elif t == _NILSXP:
        out = None
    else:
        raise RuntimeError(f"unexpected result type {t}")
    if release:
        L.ref_release()
    return out


class RefPPTB:
    """A PreprocessedTB ("Twobit" | "ACtree2") built by the reference C.

    tb: (concat uint8 array, widths int32 array) of the Trusted Band (encoded bytes);
    pp_exclude: None or int32[M] with NA_INTEGER for "keep" (R/PDict-class.R:229).
    """

    def __init__(self, algo, tb_concat, tb_widths, pp_exclude=None):
        self.algo = algo
        self._lib = lib()                  # the library this handle belongs to
        self._tb = _u8(tb_concat)          # must outlive the handle
        self._tbw = _i32(tb_widths)
        self.M = len(self._tbw)
        excl = None if pp_exclude is None else _i32(pp_exclude)
        self.high2low = np.empty(self.M, dtype=np.int32)
        self._h = self._lib.ref_pptb_build(algo.encode(), self.M, _ptr(self._tb),
                                       _ptr(self._tbw), _ptr(excl), _ptr(self.high2low))
        if not self._h:
            raise RefError(self._lib.ref_last_error().decode("latin1"))

    def blank_dups(self):
        _check(self._lib.ref_pptb_blank_dups(self._h), self._lib)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ref_pptb_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nnodes(self):
        return self._lib.ref_actree2_nnodes(self._h)

    def has_all_flinks(self):
        return bool(self._lib.ref_actree2_has_all_flinks(self._h))

    @staticmethod
    def _flank(f):
        if f is None:
            return None, None
        return _u8(f[0]), _i32(f[1])

    def match_xstring(self, head, tail, subject, max_mismatch, min_mismatch, fixed,
                      mode, keep=False):
        hc, hw = self._flank(head)
        tc, tw = self._flank(tail)
        s = _u8(subject)
        _check(self._lib.ref_match_xstring(self._h, self.M, _ptr(hc), _ptr(hw), _ptr(tc),
                                       _ptr(tw), _ptr(s), len(s), max_mismatch,
                                       min_mismatch, int(fixed[0]), int(fixed[1]), mode), self._lib)
        return _fetch_result(release=not keep, L=self._lib)

    def match_views(self, head, tail, subject, views_start, views_width, max_mismatch,
                    min_mismatch, fixed, mode, keep=False):
        hc, hw = self._flank(head)
        tc, tw = self._flank(tail)
        s, vs, vw = _u8(subject), _i32(views_start), _i32(views_width)
        _check(self._lib.ref_match_views(self._h, self.M, _ptr(hc), _ptr(hw), _ptr(tc),
                                     _ptr(tw), _ptr(s), len(s), _ptr(vs), _ptr(vw),
                                     len(vs), max_mismatch, min_mismatch,
                                     int(fixed[0]), int(fixed[1]), mode), self._lib)
        return _fetch_result(release=not keep, L=self._lib)

    def vmatch_set(self, head, tail, subj_concat, subj_widths, max_mismatch,
                   min_mismatch, fixed, collapse, weight, mode):
        hc, hw = self._flank(head)
        tc, tw = self._flank(tail)
        sc, sw = _u8(subj_concat), _i32(subj_widths)
        w = np.asarray(weight)
        is_double = w.dtype.kind == "f"
        w = np.ascontiguousarray(w, dtype=np.float64 if is_double else np.int32)
        _check(self._lib.ref_vmatch_set(self._h, self.M, _ptr(hc), _ptr(hw), _ptr(tc),
                                    _ptr(tw), _ptr(sc), _ptr(sw), len(sw), max_mismatch,
                                    min_mismatch, int(fixed[0]), int(fixed[1]),
                                    collapse, int(is_double), _ptr(w), len(w), mode), self._lib)
        return _fetch_result(L=self._lib)


# ---- the non-preprocessed dictionary path (src/match_pdict.c:174,267,519) ---------------------

def match_set_xstring(pat_concat, pat_widths, subject, max_mismatch, min_mismatch, with_indels,
                      fixed, algo, mode):
    """match_XStringSet_XString: `pattern` is a plain XStringSet, one matchPattern per element."""
    L = lib()
    pc, pw, s = _u8(pat_concat), _i32(pat_widths), _u8(subject)
    _check(L.ref_match_set_xstring(_ptr(pc), _ptr(pw), len(pw), _ptr(s), len(s), max_mismatch,
                                   min_mismatch, int(with_indels), int(fixed[0]), int(fixed[1]),
                                   algo.encode(), mode), L)
    return _fetch_result(L=L)


def match_set_views(pat_concat, pat_widths, subject, views_start, views_width, max_mismatch,
                    min_mismatch, with_indels, fixed, algo, mode):
    L = lib()
    pc, pw, s = _u8(pat_concat), _i32(pat_widths), _u8(subject)
    vs, vw = _i32(views_start), _i32(views_width)
    _check(L.ref_match_set_views(_ptr(pc), _ptr(pw), len(pw), _ptr(s), len(s), _ptr(vs), _ptr(vw),
                                 len(vs), max_mismatch, min_mismatch, int(with_indels),
                                 int(fixed[0]), int(fixed[1]), algo.encode(), mode), L)
    return _fetch_result(L=L)


def vmatch_set_set(pat_concat, pat_widths, subj_concat, subj_widths, max_mismatch, min_mismatch,
                   with_indels, fixed, algo, collapse, weight, mode):
    L = lib()
    pc, pw, sc, sw = _u8(pat_concat), _i32(pat_widths), _u8(subj_concat), _i32(subj_widths)
    w = np.asarray(weight)
    is_double = w.dtype.kind == "f"
    w = np.ascontiguousarray(w, dtype=np.float64 if is_double else np.int32)
    _check(L.ref_vmatch_set_set(_ptr(pc), _ptr(pw), len(pw), _ptr(sc), _ptr(sw), len(sw),
                                max_mismatch, min_mismatch, int(with_indels), int(fixed[0]),
                                int(fixed[1]), algo.encode(), collapse, int(is_double), _ptr(w),
                                len(w), mode), L)
    return _fetch_result(L=L)


# ---- oligonucleotideFrequency (src/letter_frequency.c:634,667) ---------------------------------

def _fetch_labelled(L):
    """(values, dim tuple or None, [labels per dimension or None]) of the current result."""
    dim = (C.c_int * 16)()
    nd = L.ref_result_dim(dim)
    dims = tuple(dim[i] for i in range(nd)) if nd > 0 else None
    labels = []
    buf = C.create_string_buffer(1 << 26)
    for which in range(max(nd, 1)):
        k = L.ref_result_labels(which, buf, len(buf))
        if k < 0:
            raise RuntimeError("label buffer too small")
        labels.append(buf.value.decode("latin1").split("\n")[:-1] if k > 0 else None)
    return _fetch_result(L=L), dims, labels


def oligo_frequency_xstring(x, width, step=1, as_prob=False, as_array=False,
                            fast_moving_side="right", with_labels=True):
    """XString_oligo_frequency (the .Call behind oligonucleotideFrequency() on a DNAString,
    R/letterFrequency.R:564-586).  Returns (values, dim, labels)."""
    L = lib()
    xb = _u8(x)
    _check(L.ref_oligo_frequency_xstring(_ptr(xb), len(xb), int(width), int(step), int(as_prob),
                                         int(as_array), int(fast_moving_side != "right"),
                                         int(with_labels)), L)
    return _fetch_labelled(L)


def oligo_frequency_set(concat, widths, width, step=1, as_prob=False, as_array=False,
                        fast_moving_side="right", with_labels=True, simplify_as="matrix"):
    """XStringSet_oligo_frequency (R/letterFrequency.R:588-612).  "matrix": values is the
    column-major N x 4^width matrix, flat."""
    L = lib()
    cb, wb = _u8(concat), _i32(widths)
    _check(L.ref_oligo_frequency_set(_ptr(cb), _ptr(wb), len(wb), int(width), int(step),
                                     int(as_prob), int(as_array), int(fast_moving_side != "right"),
                                     int(with_labels), simplify_as.encode()), L)
    return _fetch_labelled(L)


def stash_result():
    """Keep the result of the last ``keep=True`` ENDS call for mindex_combine()."""
    if lib().ref_result_stash() < 0:
        raise RuntimeError("nothing to stash")


def mindex_combine():
    """ByPos_MIndex_combine over the stashed ENDS results (src/MIndex_class.c:230)."""
    _check(lib().ref_mindex_combine())
    return _fetch_result()


def release():
    lib().ref_release()


# ---- read_fasta_files (src/read_fasta_files.c:412) ----------------------------------------------

def dna_lkup():
    """get_seqtype_conversion_lookup("B", "DNA") (R/seqtype.R): byte -> DNA code for the letters
    of DNA_ALPHABET in both cases, NA elsewhere; as long as the largest letter + 1."""
    codes = {"A": 1, "C": 2, "G": 4, "T": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12,
             "V": 7, "H": 11, "D": 13, "B": 14, "N": 15, "-": 16, "+": 32, ".": 64}
    keys = {}
    for letter, code in codes.items():
        keys[ord(letter)] = code
        keys[ord(letter.lower())] = code
    lkup = np.full(max(keys) + 1, NA_INTEGER, dtype=np.int32)
    for k, v in keys.items():
        lkup[k] = v
    return lkup


def read_fasta(text, lkup=None, nrec=-1, skip=0, seek_first_rec=False, use_names=True):
    """read_fasta_files() on one in-memory file, or on a list of them (one record counter over
    the list) -> (widths int32[n], concat uint8, names or None, warning text).  `text`: bytes."""
    L = lib()
    L.ref_read_fasta.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_int] + [C.c_int] * 4
    L.ref_result_set_widths.restype = C.c_long
    L.ref_result_set_widths.argtypes = [C.c_void_p]
    L.ref_result_set_concat.argtypes = [C.c_void_p]
    L.ref_result_set_names.restype = C.c_long
    L.ref_result_set_names.argtypes = [C.c_char_p, C.c_long]
    lk = dna_lkup() if lkup is None else _i32(lkup)
    L.ref_clear_warning()
    if isinstance(text, (list, tuple)):
        sizes = (C.c_long * len(text))(*[len(t) for t in text])
        text = b"".join(bytes(t) for t in text)
        L.ref_read_fasta_multi.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 4
        _check(L.ref_read_fasta_multi(text, sizes, len(sizes), _ptr(lk), len(lk), nrec, skip,
                                      int(seek_first_rec), int(use_names)), L)
    else:
        text = bytes(text)
        _check(L.ref_read_fasta(text, len(text), _ptr(lk), len(lk), nrec, skip, int(seek_first_rec),
                                int(use_names)), L)
    n = L.ref_result_set_length()
    widths = np.empty(n, dtype=np.int32)
    total = L.ref_result_set_widths(_ptr(widths))
    concat = np.empty(total, dtype=np.uint8)
    L.ref_result_set_concat(_ptr(concat))
    buf = C.create_string_buffer(len(text) + 16)
    k = L.ref_result_set_names(buf, len(buf))
    names = buf.value.decode("latin1").split("\n")[:-1] if k > 0 else ([] if use_names and n == 0 else None)
    warn = L.ref_last_warning().decode("latin1")
    L.ref_release()
    return widths, concat, names, warn
