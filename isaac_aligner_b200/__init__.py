"""B200-native candidate-extension path of the Isaac aligner (ungapped scoring, banded Smith-Waterman,
simple indels, shadow rescue) behind a C ABI (include/isaac_ext.h).

The compute lives in isaac_aligner_b200/csrc (CUDA, sm_100a) and is reached through
isaac_aligner_b200.capi (ctypes over libisaac_ext.so).  There is no CPU fallback: importing capi on a
machine without the built library raises, and every compute call needs a CUDA device.
"""
from .types import (CANDIDATE_DTYPE, FRAGMENT_DTYPE, MASK_WORDS, Config, BWA_SCORES, ELAND_SCORES)  # noqa: F401

__version__ = "0.1.0"
