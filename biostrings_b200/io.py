"""readDNAStringSet (FASTA) on the GPU: subject ingestion.

Mirror of R/XStringSet-io.R:45-55,203-260 (readDNAStringSet(format="fasta") ->
.read_fasta_files -> .Call "read_fasta_files") over bsgpu_subject_fasta() of the C ABI: the
file text is uploaded once, the line rules and the letter translation of
src/read_fasta_files.c:240-352 run on the device and the letters land in the resident subject
layout.  `read_fasta_to_device` keeps them there (a genome read once serves any number of
matchPDict / countPDict / oligonucleotideFrequency calls without touching the host again);
`readDNAStringSet` also brings them back as a DNAStringSet, like the reference returns.
"""
import ctypes as C
import gzip
import warnings

import numpy as np

from . import _lib
from ._lib import NA_INTEGER
from .matchpdict import DeviceSubject
from .xstring import DNA_CODES, DNAStringSet


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dna_lkup():
    """get_seqtype_conversion_lookup("B", "DNA") (R/seqtype.R): the DNA alphabet in both cases
    -> codes, NA elsewhere, as long as the largest letter + 1."""
    keys = {}
    for letter, code in DNA_CODES.items():
        keys[ord(letter)] = code
        keys[ord(letter.lower())] = code
    lkup = np.full(max(keys) + 1, NA_INTEGER, dtype=np.int32)
    for k, v in keys.items():
        lkup[k] = v
    return lkup


def _normarg_nrec(nrec):
    # R/XStringSet-io.R:.normarg_nrec
    if isinstance(nrec, bool) or not isinstance(nrec, (int, float, np.integer, np.floating)):
        raise ValueError("'nrec' must be a single integer value")
    nrec = int(nrec)
    return -1 if nrec < 0 else nrec


def _normarg_skip(skip):
    if isinstance(skip, bool) or not isinstance(skip, (int, float, np.integer, np.floating)):
        raise ValueError("'skip' must be a single non-negative integer value")
    skip = int(skip)
    if skip < 0:
        raise ValueError("'skip' cannot be negative")
    return skip


def _read_text(filepath):
    if isinstance(filepath, np.ndarray):            # file content already in (pinned) host memory
        return np.ascontiguousarray(filepath, dtype=np.uint8)
    if isinstance(filepath, (bytes, bytearray, memoryview)):
        return bytes(filepath)
    with open(filepath, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":                         # the reference reads gzip-ed files too
        with gzip.open(filepath, "rb") as f:
            return f.read()
    with open(filepath, "rb") as f:
        return f.read()


def _record_starts(arr):
    gt = np.flatnonzero(arr == ord(">"))
    return gt[(gt == 0) | (arr[np.maximum(gt - 1, 0)] == ord("\n"))]


def _check_preamble(arr, upto, name):
    """A line of sequence data in front of the first record is an error in the reference whatever
    `skip` says (parse_FASTA_file, src/read_fasta_files.c:320-325: dont_load == -1); empty lines
    and comment lines (';') are ignored."""
    lineno = 0
    for line in bytes(arr[:upto]).split(b"\n"):
        lineno += 1
        line = line.rstrip(b"\r")
        if line and not line.startswith(b";"):
            raise _lib.BsgpuError(-1, f'reading FASTA file {name}: ">" expected at beginning of line {lineno}')


def _cut(text, nrec, skip, seek_first_rec, name="<memory>"):
    """nrec / skip / seek.first.rec of parse_FASTA_file (src/read_fasta_files.c:276-316) applied
    to the text before it is uploaded: records are delimited by the lines that start with '>'.
    Returns (text to parse, number of records the text holds up to the cut)."""
    arr = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
    if nrec < 0 and skip == 0 and not seek_first_rec:
        return text, None
    starts = _record_starts(arr)
    if seek_first_rec:
        if starts.size == 0:
            raise ValueError("no FASTA record found")
        lo = int(starts[0])
    else:
        lo = 0
        if skip > 0:
            _check_preamble(arr, int(starts[0]) if starts.size else arr.size, name)
    if skip > 0:
        lo = int(starts[skip]) if skip < starts.size else len(text)
    hi = len(text)
    seen = int(starts.size)
    if nrec >= 0 and skip + nrec < starts.size:
        hi = int(starts[skip + nrec])
        seen = skip + nrec
    return text[lo:hi], seen


def _file_windows(records_per_file, nrec, skip):
    """(nrec, skip) to apply to every file of a list so that, together, they select the records
    [skip, skip + nrec) of the concatenated files: read_fasta_files() keeps one record counter over
    the whole list (src/read_fasta_files.c:447-458; parse_FASTA_file stops at `*recno >= skip + nrec`
    and loads nothing while `*recno < skip`)."""
    out, recno = [], 0
    for in_file in records_per_file:
        if nrec >= 0 and recno >= skip + nrec:
            f_nrec, f_skip = 0, 0
        else:
            f_skip = max(0, skip - recno)
            f_nrec = -1 if nrec < 0 else skip + nrec - max(recno, skip)
        out.append((f_nrec, f_skip))
        recno += in_file if f_nrec < 0 else min(in_file, f_skip + f_nrec)
    return out


def read_fasta_to_device(filepath, nrec=-1, skip=0, seek_first_rec=False, use_names=True, lkup=None):
    """-> (DeviceSubject of kind "set", widths int32[nrec], names or None).  `filepath`: a path
    (plain or gzip-ed), or the file content as bytes or as a uint8 numpy array (e.g. the view of
    a pinned buffer: the upload then runs at PCIe speed instead of the pageable-copy speed)."""
    nrec, skip = _normarg_nrec(nrec), _normarg_skip(skip)
    if not isinstance(seek_first_rec, (bool, np.bool_)):
        raise ValueError("'seek.first.rec' must be TRUE or FALSE")
    name = filepath if isinstance(filepath, str) else "<memory>"
    text, _ = _cut(_read_text(filepath), nrec, skip, bool(seek_first_rec), name)
    lk = dna_lkup() if lkup is None else np.ascontiguousarray(lkup, dtype=np.int32)
    buf = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
    h = C.c_void_p()
    info = _lib.FastaInfo()
    rc = _lib.lib().bsgpu_subject_fasta(_ptr(buf) if buf.size else None, buf.size, _ptr(lk), lk.size,
                                        C.byref(h), C.byref(info))
    if rc != 0:
        msg = _lib.lib().bsgpu_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise _lib.BsgpuError(rc, f"reading FASTA file {name}: {msg}")
        raise _lib.BsgpuError(rc, msg)
    try:
        n = info.nrec
        widths = np.ctypeslib.as_array(info.widths, shape=(max(n, 1),))[:n].astype(np.int32)
        names = None
        if use_names:
            offs = np.ctypeslib.as_array(info.desc_offset, shape=(max(n, 1),))[:n]
            lens = np.ctypeslib.as_array(info.desc_length, shape=(max(n, 1),))[:n]
            names = [bytes(text[int(o):int(o) + int(k)]).decode("latin1") for o, k in zip(offs, lens)]
        if info.ninvalid:
            warnings.warn(f"reading FASTA file {name}: ignored {info.ninvalid} invalid one-letter "
                          "sequence codes")
    finally:
        _lib.lib().bsgpu_fasta_info_free(C.byref(info))
    return DeviceSubject(h, "set", n), widths, names


def device_subject_from_letters(letters, lkup=None):
    """DNAString(<character>) with the encoding done on the device: `letters` (str, bytes or a
    uint8 array of ASCII codes) are uploaded as they are and translated by the library
    (R/XStringCodec-class.R:114-166; error text of XVector's lookup copy for a letter that is
    not in the DNA alphabet)."""
    if isinstance(letters, str):
        letters = letters.encode("latin1")
    buf = letters if isinstance(letters, np.ndarray) else np.frombuffer(bytes(letters), dtype=np.uint8)
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    lk = dna_lkup() if lkup is None else np.ascontiguousarray(lkup, dtype=np.int32)
    h = C.c_void_p()
    _lib.check(_lib.lib().bsgpu_subject_letters(_ptr(buf) if buf.size else None, buf.size, _ptr(lk), lk.size,
                                                C.byref(h)))
    return DeviceSubject(h, "xstring", int(buf.size))


def device_subject_element(dset, i):
    """One sequence of a resident set as a resident single-sequence subject (0-based i)."""
    h = C.c_void_p()
    _lib.check(_lib.lib().bsgpu_subject_element(dset.handle, int(i), C.byref(h)))
    return DeviceSubject(h, "xstring", dset_width(dset, i))


def dset_width(dset, i):
    n = _lib.lib().bsgpu_subject_nsequences(dset.handle)
    widths = np.zeros(max(n, 1), dtype=np.int32)
    _lib.check(_lib.lib().bsgpu_subject_download(dset.handle, None, _ptr(widths)))
    return int(widths[i])


def download_set(dset, names=None):
    """The letters of a resident set back in host memory as a DNAStringSet."""
    n = _lib.lib().bsgpu_subject_nsequences(dset.handle)
    widths = np.zeros(max(n, 1), dtype=np.int32)
    _lib.check(_lib.lib().bsgpu_subject_download(dset.handle, None, _ptr(widths)))
    widths = widths[:n]
    concat = np.zeros(max(int(widths.sum()), 1), dtype=np.uint8)
    _lib.check(_lib.lib().bsgpu_subject_download(dset.handle, _ptr(concat), None))
    return DNAStringSet(concat=concat[:int(widths.sum())], widths=widths, names=names)


def readDNAStringSet(filepath, format="fasta", nrec=-1, skip=0, seek_first_rec=False, use_names=True):
    """R/XStringSet-io.R:203-260 for format="fasta" (one file or a list of files)."""
    if format != "fasta":
        raise ValueError("only format=\"fasta\" is on this path")
    if isinstance(filepath, (list, tuple)):
        # one record counter over the whole list, as read_fasta_files() keeps it
        # (src/read_fasta_files.c:447-458, parse_FASTA_file: `*recno >= skip + nrec`): nrec and skip
        # select records of the concatenated files, seek.first.rec applies to every file
        nrec, skip = _normarg_nrec(nrec), _normarg_skip(skip)
        texts = [_read_text(f) for f in filepath]
        counts = [int(_record_starts(t if isinstance(t, np.ndarray) else np.frombuffer(t, dtype=np.uint8)).size)
                  for t in texts]
        parts = [readDNAStringSet(t, format, f_nrec, f_skip, seek_first_rec, use_names)
                 for t, (f_nrec, f_skip) in zip(texts, _file_windows(counts, nrec, skip))]
        names = sum((p.names or [] for p in parts), []) if use_names else None
        return DNAStringSet(concat=np.concatenate([p.concat for p in parts]) if parts else np.zeros(0, np.uint8),
                            widths=np.concatenate([p.widths for p in parts]) if parts else np.zeros(0, np.int32),
                            names=names)
    dset, widths, names = read_fasta_to_device(filepath, nrec, skip, seek_first_rec, use_names)
    with dset:
        return download_set(dset, names)
