"""biostrings_b200 -- B200-native dictionary matching for Biostrings' PDict path.

PDict preprocessing plus matchPDict / countPDict / whichPDict / vcountPDict /
vwhichPDict, exact and with max.mismatch outside a Trusted Band, computed by hand-written
sm_100a kernels behind the C ABI in include/bsgpu.h (biostrings_b200/lib/libbsgpu.so).
The Python layer mirrors the reference's R layer (same names, argument meaning and error
behaviour); it needs the shared library and a CUDA device -- there is no CPU fallback.
"""
from ._lib import BsgpuError
from .matchpdict import (DeviceSubject, countPDict, matchPDict, vcountPDict, vmatchPDict,
                         vwhichPDict, whichPDict)
from .io import device_subject_element, device_subject_from_letters, download_set, read_fasta_to_device, readDNAStringSet
from .letterfreq import (dinucleotideFrequency, mkAllStrings, oligonucleotideFrequency,
                         oligonucleotideTransitions, trinucleotideFrequency)
from .mindex import ByPos_MIndex, extractAllMatches
from .pdict import MTB_PDict, TB_PDict, make_pdict as PDict
from .xstring import (DNA_BASES, DNAString, DNAStringSet, MaskedDNAString, XStringViews,
                      decode_dna, encode_dna)

__all__ = ["PDict", "TB_PDict", "MTB_PDict", "matchPDict", "countPDict", "whichPDict",
           "vcountPDict", "vwhichPDict", "vmatchPDict", "ByPos_MIndex", "extractAllMatches", "DNAString",
           "DNAStringSet", "XStringViews", "MaskedDNAString", "DeviceSubject", "DNA_BASES",
           "encode_dna", "decode_dna", "BsgpuError", "oligonucleotideFrequency", "dinucleotideFrequency",
           "trinucleotideFrequency", "oligonucleotideTransitions", "mkAllStrings", "readDNAStringSet",
           "read_fasta_to_device", "device_subject_element", "device_subject_from_letters", "download_set"]
