"""Python restatement of the reference's FASTA parser (TEST INFRASTRUCTURE ONLY: only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this).

Follows parse_FASTA_file(), src/read_fasta_files.c:240-352 (line loop: empty lines and ';'
lines ignored, '>' opens a record, anything else is sequence data; data before the first '>'
is an error), translate() :31-50 (bytes the lookup does not know are dropped and counted),
delete_trailing_LF_or_CRLF (XVector, un-vendored: strips LF, and CR only when it precedes that
LF) and the lookup R builds for readDNAStringSet (R/seqtype.R, get_seqtype_conversion_lookup
("B", "DNA"): the DNA alphabet in both cases).  Pinned against the reference's own
read_fasta_files() compiled into oracle/_ref (tests/test_fasta_hostlogic.py), including the
reference's fixture inst/extdata/fastaEx.fa.
"""
import numpy as np

DNA_CODES = {"A": 1, "C": 2, "G": 4, "T": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12,
             "V": 7, "H": 11, "D": 13, "B": 14, "N": 15, "-": 16, "+": 32, ".": 64}


def dna_lkup16():
    """256-entry int16 lookup, -1 = not a DNA letter (what the C ABI takes)."""
    lk = np.full(256, -1, dtype=np.int16)
    for letter, code in DNA_CODES.items():
        lk[ord(letter)] = code
        lk[ord(letter.lower())] = code
    return lk


def strip_eol(line):
    if line.endswith(b"\n"):
        line = line[:-1]
        if line.endswith(b"\r"):
            line = line[:-1]
    return line


def desc_at(text, off):
    """description text starting at byte `off` (the byte after '>')."""
    end = text.find(b"\n", off)
    line = text[off:] if end < 0 else text[off:end + 1]
    return strip_eol(line).decode("latin1")


def read_fasta(text, lkup=None):
    """-> (widths int32[nrec], concat uint8 codes, names, ninvalid)."""
    lk = dna_lkup16() if lkup is None else lkup
    widths, names, chunks = [], [], []
    ninvalid = 0
    lineno = 0
    pos = 0
    n = len(text)
    while pos < n:
        end = text.find(b"\n", pos)
        raw = text[pos:] if end < 0 else text[pos:end + 1]
        pos += len(raw)
        lineno += 1
        line = strip_eol(raw)
        if len(line) == 0 or line.startswith(b";"):
            continue
        if line.startswith(b">"):
            names.append(line[1:].decode("latin1"))
            widths.append(0)
            continue
        if not widths:
            raise ValueError(f'">" expected at beginning of line {lineno}')
        codes = lk[np.frombuffer(line, dtype=np.uint8)]
        ok = codes >= 0
        ninvalid += int((~ok).sum())
        chunks.append(codes[ok].astype(np.uint8))
        widths[-1] += int(ok.sum())
    concat = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return np.array(widths, dtype=np.int32), concat, names, ninvalid
