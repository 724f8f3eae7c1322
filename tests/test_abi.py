"""The C ABI without a GPU: the shared library loads, exports every symbol that
include/bsgpu.h declares, validates arguments with the reference's messages before it
needs a device, and refuses to compute without one (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from biostrings_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "bsgpu.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsgpu_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(L, name), f"libbsgpu.so does not export {name}"
    assert sorted(_lib.SYMBOLS) == syms, "biostrings_b200/_lib.py SYMBOLS out of sync with bsgpu.h"
    assert b"sm_100a" in L.bsgpu_version()


def test_cuda_code_is_sm_100a():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def _build(algo, pats, head=None, tail=None, excl=None):
    L = _lib.lib()
    tb = np.concatenate([np.asarray(p, np.uint8) for p in pats]) if pats else np.zeros(0, np.uint8)
    w = np.array([len(p) for p in pats], dtype=np.int32)
    h2l = np.zeros(len(pats), np.int32)
    out = C.c_void_p()
    ex = None if excl is None else np.asarray(excl, np.int32)
    rc = L.bsgpu_pdict3parts_build(algo, len(pats), tb.ctypes.data_as(C.c_void_p),
                                   w.ctypes.data_as(C.c_void_p), None, None, None, None,
                                   None if ex is None else ex.ctypes.data_as(C.c_void_p),
                                   h2l.ctypes.data_as(C.c_void_p), C.byref(out))
    return rc, L.bsgpu_last_error().decode(), out


@pytest.mark.parametrize("algo,pats,msg", [
    (0, [[1, 2, 4, 8], [1, 15, 4, 8]], "non base DNA letter found in Trusted Band for pattern 2"),
    (1, [[1, 2, 4, 8], [1, 15, 4, 8]], "non-base DNA letter found in Trusted Band for pattern 2"),
    (0, [[1, 2, 4, 8], [1, 2, 4]], "element 2 in Trusted Band has a different length than first element"),
    (1, [[1, 2, 4, 8], [1, 2, 4]], "all the trusted regions must have the same length"),
    (0, [], "Trusted Band is empty"),
    (0, [[], [1]], "first element in Trusted Band is of length 0"),
    (1, [[1], []], "empty trusted region for pattern 2"),
    (1, [[1] * 15], "the width of the Trusted Band must be <= 14 when 'type=\"Twobit\"'"),
])
def test_build_errors_carry_the_reference_messages(algo, pats, msg):
    # src/match_pdict_ACtree2.c:762,803,815,821; src/match_pdict_Twobit.c:123,127,135,140
    rc, err, out = _build(algo, pats)
    assert rc == -1 and err == msg and not out.value


def test_no_device_means_no_compute():
    L = _lib.lib()
    if L.bsgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    rc, err, out = _build(0, [[1, 2, 4, 8]])
    assert rc == -2 and "no CPU fallback" in err and not out.value
    s = np.array([1, 2, 4, 8], np.uint8)
    h = C.c_void_p()
    assert L.bsgpu_subject_xstring(s.ctypes.data_as(C.c_void_p), 4, C.byref(h)) == -2
    import biostrings_b200 as bs
    with pytest.raises(bs.BsgpuError, match="no CPU fallback"):
        bs.PDict(["ACGT"])
    # the rows beside the dictionary path refuse the same way
    out = np.zeros(16, np.int32)
    assert L.bsgpu_XString_oligo_frequency(s.ctypes.data_as(C.c_void_p), 4, 2, 1, 0, 0,
                                           out.ctypes.data_as(C.c_void_p), None) == -2
    with pytest.raises(bs.BsgpuError, match="no CPU fallback"):
        bs.oligonucleotideFrequency("ACGT", 2)
    with pytest.raises(bs.BsgpuError, match="no CPU fallback"):
        bs.readDNAStringSet(b">a\nACGT\n")


def test_argument_errors_of_the_new_rows_need_no_device():
    """width / step of oligonucleotideFrequency are validated before a device is needed, with
    the reference's texts (src/utils.c:101-102; R/letterFrequency.R:16-75)."""
    L = _lib.lib()
    s = np.array([1, 2, 4, 8], np.uint8)
    out = np.zeros(16, np.int32)
    for width, step, msg in [(0, 1, "'buflength' must be >= 1 and <= 15"),
                             (16, 1, "'buflength' must be >= 1 and <= 15"),
                             (2, 0, "'step' must be a positive integer")]:
        rc = L.bsgpu_XString_oligo_frequency(s.ctypes.data_as(C.c_void_p), 4, width, step, 0, 0,
                                             out.ctypes.data_as(C.c_void_p), None)
        assert rc == -1 and msg in L.bsgpu_last_error().decode()
    import biostrings_b200 as bs
    for kw, msg in [(dict(width="a"), "'width' must be a single integer"),
                    (dict(width=0), "'width' must be an integer >= 1"),
                    (dict(width=2, step=0), "'step' must be a positive integer"),
                    (dict(width=2, as_prob=1), "'as.prob' must be TRUE or FALSE"),
                    (dict(width=2, fast_moving_side="up"), "should be one of"),
                    (dict(width=2, with_labels=None), "'with.labels' must be TRUE or FALSE")]:
        with pytest.raises(ValueError, match=msg):
            bs.oligonucleotideFrequency("ACGT", **kw)
    with pytest.raises(ValueError, match="'skip' cannot be negative"):
        bs.readDNAStringSet(b">a\nA\n", skip=-1)
    assert bs.mkAllStrings(["A", "T"], 2) == ["AA", "AT", "TA", "TT"]      # test-matchPDict.R:127
    assert bs.mkAllStrings(bs.DNA_BASES, 0) == [""]


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "biostrings_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "libbiostrings_ref" not in src, f
