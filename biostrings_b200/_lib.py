"""ctypes binding of libbsgpu.so -- the C ABI declared in include/bsgpu.h.

There is deliberately no fallback: if the shared library is missing this raises, and
every compute entry point of the library fails without a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libbsgpu.so")

NA_INTEGER = -2147483648
MATCHES_AS_NULL, MATCHES_AS_WHICH, MATCHES_AS_COUNTS, MATCHES_AS_ENDS = 0, 1, 2, 4
ALGO = {"ACtree2": 0, "Twobit": 1}

# every symbol include/bsgpu.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "bsgpu_last_error", "bsgpu_version", "bsgpu_device_count", "bsgpu_set_device",
    "bsgpu_pdict3parts_build", "bsgpu_pdict3parts_set_flanks", "bsgpu_pdict3parts_blank_dups", "bsgpu_pdict3parts_free",
    "bsgpu_pdict3parts_length", "bsgpu_pdict3parts_tb_width",
    "bsgpu_subject_xstring", "bsgpu_subject_views", "bsgpu_subject_set",
    "bsgpu_subject_chunk", "bsgpu_subject_repack", "bsgpu_subject_free", "bsgpu_subject_nletters",
    "bsgpu_result_free", "bsgpu_match", "bsgpu_vmatch", "bsgpu_count_device",
    "bsgpu_match_PDict3Parts_XString", "bsgpu_match_PDict3Parts_XStringViews",
    "bsgpu_vmatch_PDict3Parts_XStringSet", "bsgpu_mindex_combine",
    "bsgpu_last_plan", "bsgpu_launch_count", "bsgpu_probe_peaks",
    "bsgpu_kernel_timing", "bsgpu_kernel_times", "bsgpu_patterns_build",
    "bsgpu_oligo_frequency", "bsgpu_oligo_count_device", "bsgpu_XString_oligo_frequency",
    "bsgpu_XStringSet_oligo_frequency",
    "bsgpu_subject_fasta", "bsgpu_fasta_info_free", "bsgpu_subject_nsequences", "bsgpu_subject_element",
    "bsgpu_subject_download", "bsgpu_subject_letters", "bsgpu_count_indels", "bsgpu_count_host",
    "bsgpu_shared_counts_create", "bsgpu_shared_counts_open", "bsgpu_shared_counts_close",
]


class BsgpuError(RuntimeError):
    """An error returned through the C ABI; the text is the reference's error() message
    for the same condition where the reference has one."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class Result(C.Structure):
    _fields_ = [("M", C.c_int32), ("n_counts", C.c_int64), ("counts", C.POINTER(C.c_int32)),
                ("dcounts", C.POINTER(C.c_double)), ("n_which", C.c_int64),
                ("which", C.POINTER(C.c_int32)), ("n_rows", C.c_int64),
                ("ends_offsets", C.POINTER(C.c_int64)), ("ends", C.POINTER(C.c_int32))]


class FastaInfo(C.Structure):
    _fields_ = [("nrec", C.c_int32), ("ninvalid", C.c_int64), ("nletters", C.c_int64),
                ("widths", C.POINTER(C.c_int32)), ("desc_offset", C.POINTER(C.c_int64)),
                ("desc_length", C.POINTER(C.c_int32))]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -m biostrings_b200.build` "
            "(there is no CPU fallback for the PDict matching path)")
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.bsgpu_last_error.restype = C.c_char_p
    L.bsgpu_version.restype = C.c_char_p
    L.bsgpu_set_device.argtypes = [C.c_int]
    L.bsgpu_pdict3parts_build.argtypes = [C.c_int, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                          C.POINTER(vp)]
    L.bsgpu_pdict3parts_set_flanks.argtypes = [vp, vp, vp, vp, vp]
    L.bsgpu_pdict3parts_blank_dups.argtypes = [vp]
    L.bsgpu_pdict3parts_free.argtypes = [vp]
    L.bsgpu_pdict3parts_free.restype = None
    L.bsgpu_pdict3parts_length.argtypes = [vp]
    L.bsgpu_pdict3parts_tb_width.argtypes = [vp]
    L.bsgpu_subject_xstring.argtypes = [vp, i64, C.POINTER(vp)]
    L.bsgpu_subject_views.argtypes = [vp, i64, vp, vp, i32, C.POINTER(vp)]
    L.bsgpu_subject_set.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.bsgpu_subject_chunk.argtypes = [vp, i64, i64, i64, i64, i64, C.POINTER(vp)]
    L.bsgpu_subject_repack.argtypes = [vp, vp]
    L.bsgpu_subject_free.argtypes = [vp]
    L.bsgpu_subject_free.restype = None
    L.bsgpu_subject_nletters.argtypes = [vp]
    L.bsgpu_subject_nletters.restype = i64
    L.bsgpu_result_free.argtypes = [C.POINTER(Result)]
    L.bsgpu_result_free.restype = None
    L.bsgpu_match.argtypes = [C.POINTER(vp), C.c_int, vp] + [C.c_int] * 5 + [C.POINTER(Result)]
    L.bsgpu_vmatch.argtypes = [C.POINTER(vp), C.c_int, vp] + [C.c_int] * 4 + \
        [C.c_int, C.c_int, vp, i64, C.c_int, C.POINTER(Result)]
    L.bsgpu_count_device.argtypes = [C.POINTER(vp), C.c_int, vp] + [C.c_int] * 4 + \
        [vp, vp, C.POINTER(C.c_int)]
    L.bsgpu_match_PDict3Parts_XString.argtypes = [vp, vp, i64] + [C.c_int] * 5 + [C.POINTER(Result)]
    L.bsgpu_match_PDict3Parts_XStringViews.argtypes = [vp, vp, i64, vp, vp, i32] + \
        [C.c_int] * 5 + [C.POINTER(Result)]
    L.bsgpu_vmatch_PDict3Parts_XStringSet.argtypes = [vp, vp, vp, i32] + [C.c_int] * 4 + \
        [C.c_int, C.c_int, vp, i64, C.c_int, C.POINTER(Result)]
    L.bsgpu_shared_counts_create.argtypes = [i64, C.POINTER(vp), vp]
    L.bsgpu_shared_counts_open.argtypes = [vp, C.POINTER(vp)]
    L.bsgpu_shared_counts_close.argtypes = [vp, C.c_int]
    L.bsgpu_mindex_combine.argtypes = [C.POINTER(C.POINTER(Result)), C.c_int, C.POINTER(Result)]
    L.bsgpu_last_plan.restype = C.c_char_p
    L.bsgpu_launch_count.restype = i64
    L.bsgpu_probe_peaks.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), i64]
    L.bsgpu_patterns_build.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.bsgpu_oligo_frequency.argtypes = [vp] + [C.c_int] * 5 + [vp, vp]
    L.bsgpu_oligo_count_device.argtypes = [vp] + [C.c_int] * 3 + [vp, vp]
    L.bsgpu_XString_oligo_frequency.argtypes = [vp, i64] + [C.c_int] * 4 + [vp, vp]
    L.bsgpu_XStringSet_oligo_frequency.argtypes = [vp, vp, i32] + [C.c_int] * 5 + [vp, vp]
    L.bsgpu_subject_fasta.argtypes = [vp, i64, vp, C.c_int, C.POINTER(vp), C.POINTER(FastaInfo)]
    L.bsgpu_fasta_info_free.argtypes = [C.POINTER(FastaInfo)]
    L.bsgpu_fasta_info_free.restype = None
    L.bsgpu_subject_nsequences.argtypes = [vp]
    L.bsgpu_subject_element.argtypes = [vp, i32, C.POINTER(vp)]
    L.bsgpu_subject_download.argtypes = [vp, vp, vp]
    L.bsgpu_subject_letters.argtypes = [vp, i64, vp, C.c_int, C.POINTER(vp)]
    L.bsgpu_count_indels.argtypes = [vp, vp] + [C.c_int] * 4 + [vp]
    L.bsgpu_count_host.argtypes = [C.POINTER(vp), C.c_int, vp, i64, i64, i64, i64, i64] + [C.c_int] * 4 + [vp, vp, vp]
    L.bsgpu_kernel_timing.argtypes = [C.c_int]
    L.bsgpu_kernel_times.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    _lib = L
    return L


def kernel_times():
    """(name, ms) of the hot kernels launched since bsgpu_kernel_timing(1), in launch order."""
    names = (C.c_char_p * 8192)()
    ms = (C.c_float * 8192)()
    n = lib().bsgpu_kernel_times(names, ms, 8192)
    if n < 0:
        raise RuntimeError(lib().bsgpu_last_error().decode())
    return [(names[i].decode(), float(ms[i])) for i in range(n)]


def check(rc):
    if rc != 0:
        raise BsgpuError(rc, lib().bsgpu_last_error().decode("utf-8", "replace"))
