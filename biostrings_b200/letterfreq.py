"""oligonucleotideFrequency / dinucleotideFrequency / trinucleotideFrequency /
oligonucleotideTransitions / mkAllStrings on the GPU.

Mirror of R/letterFrequency.R:500-707 (argument normalisation, subject dispatch, labels)
over bsgpu_oligo_frequency() of the C ABI; the counting itself -- the rolling two-bit
signature walk of src/letter_frequency.c:183-267 -- is oligo_freq_kernel.  The identity the
reference states, countPDict(PDict(all k-mers), x) == oligonucleotideFrequency(x, k)
(R/letterFrequency.R:541-553), ties this to the dictionary path and is tested.
"""
import ctypes as C

import numpy as np

from . import _lib
from .matchpdict import DeviceSubject
from .xstring import DNA_BASES, DNAString, DNAStringSet, MaskedDNAString, XStringViews


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Freq(np.ndarray):
    """numpy array carrying R's `names` (vector) or `dimnames` (matrix / array)."""

    names = None
    dimnames = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self.names = getattr(obj, "names", None)
            self.dimnames = getattr(obj, "dimnames", None)

    def as_dict(self):
        """{label: value} of a named vector."""
        if self.names is None:
            raise ValueError("no names")
        return dict(zip(self.names, self.tolist()))


def _freq(a, names=None, dimnames=None):
    out = np.asarray(a).view(Freq)
    out.names, out.dimnames = names, dimnames
    return out


# ---- argument normalisation (R/letterFrequency.R:16-75) ---------------------------------------

def _is_single_number(v):
    return isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool) \
        and not (isinstance(v, float) and np.isnan(v))


def normarg_width(width, argname="width", min_width=1):
    if not _is_single_number(width):
        raise ValueError(f"'{argname}' must be a single integer")
    width = int(width)
    if width < min_width:
        raise ValueError(f"'{argname}' must be an integer >= {min_width}")
    return width


def normarg_step(step):
    if not _is_single_number(step):
        raise ValueError("'step' must be a single integer")
    step = int(step)
    if step <= 0:
        raise ValueError("'step' must be a positive integer")
    return step


def _true_or_false(v, what):
    if not isinstance(v, (bool, np.bool_)):
        raise ValueError(f"'{what}' must be TRUE or FALSE")
    return bool(v)


def normarg_fast_moving_side(fast_moving_side, as_array=False):
    if as_array:
        return "left"
    if not isinstance(fast_moving_side, str):
        raise ValueError("'fast.moving.side' must be a single string")
    hits = [c for c in ("left", "right") if c.startswith(fast_moving_side)] if fast_moving_side else []
    if len(hits) != 1:                              # match.arg
        raise ValueError("'arg' should be one of 'left', 'right'")
    return hits[0]


def normarg_simplify_as(simplify_as, as_array):
    if not isinstance(simplify_as, str):
        raise ValueError("'simplify.as' must be a single string")
    hits = [c for c in ("matrix", "list", "collapsed") if c.startswith(simplify_as)] if simplify_as else []
    if len(hits) != 1:
        raise ValueError("'arg' should be one of 'matrix', 'list', 'collapsed'")
    if hits[0] == "matrix" and as_array:
        raise ValueError("'as.array' cannot be TRUE when 'simplify.as' is \"matrix\"")
    return hits[0]


# ---- labels (mk_all_oligos, src/letter_frequency.c:368-412; mkAllStrings, R:500-535) -----------

def _all_oligos(alphabet, width, invert):
    """The 4^width (or len(alphabet)^width) words in table order: entry i spells i in base
    len(alphabet), first letter most significant, or least significant when `invert`."""
    k = len(alphabet)
    if width == 0:
        return [""]
    if k == 0:
        return []
    n = k ** width
    idx = np.arange(n, dtype=np.int64)
    digits = [(idx // k ** (width - 1 - j)) % k for j in range(width)]      # letter j, "right" order
    if invert:
        digits = digits[::-1]
    # letters may be longer than one character in mkAllStrings(); join per word
    if all(len(a) == 1 for a in alphabet):
        tab = np.frombuffer("".join(alphabet).encode("latin1"), dtype="S1")
        words = np.stack([tab[d] for d in digits], axis=1)
        return [w.tobytes().decode("latin1") for w in words]
    return ["".join(alphabet[int(d[i])] for d in digits) for i in range(n)]


def mkAllStrings(alphabet, width, fast_moving_side="right"):
    """R/letterFrequency.R:526-534."""
    if isinstance(alphabet, str) or not all(isinstance(a, str) for a in alphabet):
        raise ValueError("'alphabet' must be a character vector")
    width = normarg_width(width, min_width=0)
    side = normarg_fast_moving_side(fast_moving_side)
    return _all_oligos(list(alphabet), width, side != "right")


# ---- the counting call --------------------------------------------------------------------------

def oligo_frequency_device(dsubj, width, step=1, invert=False, collapsed=True, as_prob=False):
    """bsgpu_oligo_frequency() on a resident subject -> flat numpy array (column-major
    nseq x 4^width when not collapsed), int32 counts or float64 frequencies."""
    ncol = 1 << (2 * width) if 1 <= width <= 15 else 1
    nrow = 1 if collapsed else dsubj.length
    if as_prob:
        out = np.zeros(nrow * ncol, dtype=np.float64)
        counts, freq = None, out
    else:
        out = np.zeros(nrow * ncol, dtype=np.int32)
        counts, freq = out, None
    _lib.check(_lib.lib().bsgpu_oligo_frequency(dsubj.handle, int(width), int(step), int(invert),
                                                int(collapsed), int(as_prob), _ptr(counts), _ptr(freq)))
    return out


def _format(values, width, labels, invert, as_array, names=None):
    """format_oligo_freqs(), src/letter_frequency.c:414-430.  `names`: the labels when the
    caller already has them (the rows of a list share one list of names)."""
    if as_array:
        arr = values.reshape((4,) * width, order="F")
        return _freq(arr, dimnames=tuple(list(DNA_BASES) for _ in range(width)) if labels else None)
    if labels and names is None:
        names = _all_oligos(list(DNA_BASES), width, invert)
    return _freq(values, names=names if labels else None)


def oligonucleotideFrequency(x, width, step=1, as_prob=False, as_array=False,
                             fast_moving_side="right", with_labels=True, simplify_as="matrix"):
    """R/letterFrequency.R:557-637.  x: DNAString (or str), DNAStringSet (or list of str),
    XStringViews, MaskedDNAString, or a DeviceSubject.  Returns a Freq (named vector, matrix
    with column names, 4 x ... x 4 array) or, for simplify_as="list", a list of them."""
    if isinstance(x, MaskedDNAString):
        # as(x, "XStringViews") then simplify.as = "collapsed" (R:627-637)
        x, simplify_as = x.as_views(), "collapsed"
    if isinstance(x, XStringViews):
        x = x.to_set()                              # fromXStringViewsToStringSet (R:618)
    elif isinstance(x, (list, tuple)):
        x = DNAStringSet(x)
    is_set = isinstance(x, DNAStringSet) or (isinstance(x, DeviceSubject) and x.kind in ("set", "views"))
    width = normarg_width(width)
    step = normarg_step(step)
    as_prob = _true_or_false(as_prob, "as.prob")
    as_array = _true_or_false(as_array, "as.array")
    side = normarg_fast_moving_side(fast_moving_side, as_array)
    with_labels = _true_or_false(with_labels, "with.labels")
    if is_set:
        simplify_as = normarg_simplify_as(simplify_as, as_array)
    invert = side != "right"
    if isinstance(x, DeviceSubject):
        dsubj, owned = x, False
    elif is_set:
        dsubj, owned = DeviceSubject.from_set(x), True
    else:
        dsubj, owned = DeviceSubject.from_xstring(x if isinstance(x, DNAString) else DNAString(x)), True
    try:
        collapsed = not is_set or simplify_as == "collapsed"
        flat = oligo_frequency_device(dsubj, width, step, invert, collapsed, as_prob)
        nrow = dsubj.length
    finally:
        if owned:
            dsubj.close()
    if collapsed:
        return _format(flat, width, with_labels, invert, as_array)
    mat = flat.reshape((nrow, 1 << (2 * width)), order="F")
    if simplify_as == "matrix":
        # set_oligo_freqs_colnames(), src/letter_frequency.c:432-446
        names = _all_oligos(list(DNA_BASES), width, invert) if with_labels else None
        return _freq(mat, dimnames=(None, names) if with_labels else None)
    names = _all_oligos(list(DNA_BASES), width, invert) if with_labels and not as_array else None
    return [_format(np.ascontiguousarray(mat[i]), width, with_labels, invert, as_array, names)
            for i in range(nrow)]


def dinucleotideFrequency(x, step=1, as_prob=False, as_matrix=False, fast_moving_side="right",
                          with_labels=True, **kw):
    """R/letterFrequency.R:647-656."""
    return oligonucleotideFrequency(x, 2, step=step, as_prob=as_prob, as_array=as_matrix,
                                    fast_moving_side=fast_moving_side, with_labels=with_labels, **kw)


def trinucleotideFrequency(x, step=1, as_prob=False, as_array=False, fast_moving_side="right",
                           with_labels=True, **kw):
    """R/letterFrequency.R:658-667."""
    return oligonucleotideFrequency(x, 3, step=step, as_prob=as_prob, as_array=as_array,
                                    fast_moving_side=fast_moving_side, with_labels=with_labels, **kw)


def oligonucleotideTransitions(x, left=1, right=1, as_prob=False):
    """R/letterFrequency.R:669-707: counts of width left+right reshaped row by row into a
    4^left x 4^right matrix (rows = left words), optionally divided by the row sums."""
    if not _is_single_number(left):
        raise ValueError("'left' must be a single integer value")
    if int(left) < 1:
        raise ValueError("'left' must be >= 1")
    if not _is_single_number(right):
        raise ValueError("'right' must be a single integer value")
    if int(right) < 1:
        raise ValueError("'right' must be >= 1")
    as_prob = _true_or_false(as_prob, "as.prob")
    left, right = int(left), int(right)
    if isinstance(x, (DNAStringSet, XStringViews, list, tuple)):
        freqs = oligonucleotideFrequency(x, left + right, simplify_as="collapsed")
    else:
        freqs = oligonucleotideFrequency(x, left + right)
    ans = np.asarray(freqs).reshape((4 ** left, 4 ** right))
    if as_prob:
        with np.errstate(invalid="ignore", divide="ignore"):
            ans = ans / ans.sum(axis=1, keepdims=True)          # 0/0 -> NaN as in R
    return _freq(ans, dimnames=(_all_oligos(list(DNA_BASES), left, False),
                                _all_oligos(list(DNA_BASES), right, False)))
