"""Minimal sequence containers for the PDict path: just enough of DNAString /
DNAStringSet / XStringViews / MaskedDNAString to carry encoded bytes to the C ABI.

The hot path only ever *reads* sequences as encoded bytes (Chars_holder{ptr,length},
src/XStringSet_class.c:37-51), so these classes hold numpy uint8 arrays in Biostrings'
DNA encoding (R/XStringCodec-class.R:114,129-139; R/IUPAC_CODE_MAP.R:5-21).
"""
import numpy as np

DNA_BASES = ("A", "C", "G", "T")
DNA_CODES = {"A": 1, "C": 2, "G": 4, "T": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10,
             "K": 12, "V": 7, "H": 11, "D": 13, "B": 14, "N": 15, "-": 16, "+": 32, ".": 64}

_ENCODE = np.zeros(256, dtype=np.uint8)
_DECODE = np.zeros(256, dtype=np.uint8)
for _l, _c in DNA_CODES.items():
    _ENCODE[ord(_l)] = _c
    _ENCODE[ord(_l.lower())] = _c          # lower case is folded at encode time
    _DECODE[_c] = ord(_l)


def encode_dna(x):
    """str / bytes / uint8 array of codes -> contiguous uint8 array of DNA codes."""
    if isinstance(x, DNAString):
        return x.codes
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8)
    if isinstance(x, str):
        x = x.encode("ascii")
    raw = np.frombuffer(bytes(x), dtype=np.uint8)
    out = _ENCODE[raw]
    if out.size and not out.all():
        bad = int(raw[np.flatnonzero(out == 0)[0]])
        raise ValueError(f"key {bad} (char '{chr(bad)}') not in lookup table")
    return out


def decode_dna(codes):
    return _DECODE[np.asarray(codes, dtype=np.uint8)].tobytes().decode("ascii")


class DNAString:
    def __init__(self, x):
        self.codes = encode_dna(x)

    def __len__(self):
        return int(self.codes.size)

    def __str__(self):
        return decode_dna(self.codes)


class DNAStringSet:
    """Concatenated codes + per-element widths (+ optional names)."""

    def __init__(self, x=None, *, concat=None, widths=None, names=None):
        if isinstance(x, DNAStringSet):
            concat, widths, names = x.concat, x.widths, x.names if names is None else names
        elif x is not None:
            if isinstance(x, dict):
                names, x = list(x.keys()), list(x.values())
            enc = [encode_dna(s) for s in x]
            widths = np.fromiter((e.size for e in enc), dtype=np.int32, count=len(enc))
            concat = np.concatenate(enc) if enc else np.zeros(0, np.uint8)
        self.concat = np.ascontiguousarray(concat, dtype=np.uint8)
        self.widths = np.ascontiguousarray(widths, dtype=np.int32)
        self.names = None if names is None else list(names)
        self.offsets = np.zeros(len(self.widths) + 1, dtype=np.int64)
        np.cumsum(self.widths, out=self.offsets[1:])

    def __len__(self):
        return int(self.widths.size)

    def __getitem__(self, i):
        return DNAString(self.concat[self.offsets[i]:self.offsets[i + 1]])

    def width(self):
        return self.widths

    def is_constant_width(self):
        return bool(self.widths.size == 0 or (self.widths == self.widths[0]).all())

    def has_only_base_letters(self):
        c = self.concat
        return bool(((c == 1) | (c == 2) | (c == 4) | (c == 8)).all())

    def narrow(self, start0, width):
        """Element-wise sub-sequences [start0, start0 + width), vectorised."""
        start0 = np.asarray(start0, dtype=np.int64)
        width = np.asarray(width, dtype=np.int64)
        out_off = np.zeros(len(self) + 1, dtype=np.int64)
        np.cumsum(width, out=out_off[1:])
        total = int(out_off[-1])
        if total == 0:
            return DNAStringSet(concat=np.zeros(0, np.uint8), widths=width.astype(np.int32))
        # source index of every output letter
        elt = np.repeat(np.arange(len(self), dtype=np.int64), width)
        idx = np.arange(total, dtype=np.int64) - out_off[elt] + self.offsets[elt] + start0[elt]
        return DNAStringSet(concat=self.concat[idx], widths=width.astype(np.int32))


class XStringViews:
    """Views on a DNAString subject (1-based starts)."""

    def __init__(self, subject, start, width=None, end=None):
        self.subject = subject if isinstance(subject, DNAString) else DNAString(subject)
        self.start = np.atleast_1d(np.asarray(start, dtype=np.int32))
        if width is None:
            width = np.atleast_1d(np.asarray(end, dtype=np.int32)) - self.start + 1
        self.width = np.broadcast_to(np.atleast_1d(np.asarray(width, dtype=np.int32)),
                                     self.start.shape).copy()

    def __len__(self):
        return int(self.start.size)

    def to_set(self):
        """as(views, "XStringSet"): out-of-limits views are trimmed
        (R/XStringViews-class.R:94-109, used by vcountPDict/vwhichPDict,
        R/matchPDict.R:843,891)."""
        L = len(self.subject)
        a = np.maximum(self.start.astype(np.int64), 1)
        b = np.minimum(self.start.astype(np.int64) + self.width - 1, L)
        w = np.maximum(b - a + 1, 0)
        parts = [self.subject.codes[a[i] - 1:a[i] - 1 + w[i]] for i in range(len(a))]
        return DNAStringSet(concat=np.concatenate(parts) if parts else np.zeros(0, np.uint8),
                            widths=w.astype(np.int32))


class MaskedDNAString:
    """A DNAString with masked ranges (1-based, inclusive)."""

    def __init__(self, subject, mask_start, mask_end):
        self.subject = subject if isinstance(subject, DNAString) else DNAString(subject)
        self.mask_start = np.atleast_1d(np.asarray(mask_start, dtype=np.int64))
        self.mask_end = np.atleast_1d(np.asarray(mask_end, dtype=np.int64))

    def unmasked(self):
        return self.subject

    def as_views(self):
        """Views on the gaps of the collapsed masks (R/MaskedXString-class.R:244-252)."""
        L = len(self.subject)
        delta = np.zeros(L + 2, dtype=np.int64)
        s = np.clip(self.mask_start, 1, L + 1)
        e = np.clip(self.mask_end, 0, L)
        ok = e >= s
        np.add.at(delta, s[ok], 1)
        np.add.at(delta, e[ok] + 1, -1)
        unmasked = np.cumsum(delta)[1:L + 1] == 0
        d = np.diff(np.concatenate([[0], unmasked.astype(np.int8), [0]]))
        starts = np.flatnonzero(d == 1) + 1
        ends = np.flatnonzero(d == -1)
        return XStringViews(self.subject, starts, ends - starts + 1)
