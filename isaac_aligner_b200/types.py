"""numpy / ctypes mirrors of the structs in include/isaac_ext.h."""
import ctypes
import numpy as np

MASK_WORDS = 16  # ISAAC_EXT_MASK_WORDS

# isaac_ext_fragment_t, 64 bytes
FRAGMENT_DTYPE = np.dtype([
    ("position", "<i8"), ("logProbability", "<f8"), ("contigId", "<u4"), ("readId", "<u4"),
    ("cigarOffset", "<u4"), ("smithWatermanScore", "<u4"), ("observedLength", "<u4"),
    ("mismatchCount", "<u2"), ("matchesInARow", "<u2"), ("gapCount", "<u2"), ("editDistance", "<u2"),
    ("uniqueSeedCount", "<u2"), ("repeatSeedsCount", "<u2"), ("nonUniqueSeedOffsetFirst", "<u2"),
    ("nonUniqueSeedOffsetSecond", "<u2"), ("firstSeedIndex", "<i2"), ("lowClipped", "<u2"),
    ("highClipped", "<u2"), ("cigarLength", "<u2"), ("reverse", "u1"), ("readIndex", "u1"), ("matchCount", "<u2"),
])
assert FRAGMENT_DTYPE.itemsize == 64

# isaac_ext_candidate_t, 16 bytes
CANDIDATE_DTYPE = np.dtype([("position", "<i8"), ("readId", "<u4"), ("contigStrand", "<u4")])
assert CANDIDATE_DTYPE.itemsize == 16

# gapMatch, gapMismatch, gapOpen, gapExtend, minGapExtend (AlignOptions.cpp:55-56)
BWA_SCORES = (0, -3, -11, -4, -20)
ELAND_SCORES = (2, -1, -15, -3, -25)


class Config(ctypes.Structure):
    """isaac_ext_config_t"""
    _fields_ = [
        ("gapMatchScore", ctypes.c_int32), ("gapMismatchScore", ctypes.c_int32), ("gapOpenScore", ctypes.c_int32),
        ("gapExtendScore", ctypes.c_int32), ("minGapExtendScore", ctypes.c_int32),
        ("repeatThreshold", ctypes.c_uint32), ("maxSeedsPerRead", ctypes.c_uint32),
        ("gappedMismatchesMax", ctypes.c_uint32), ("semialignedGapLimit", ctypes.c_uint32),
        ("avoidSmithWaterman", ctypes.c_uint32), ("maxReadLength", ctypes.c_uint32),
        ("device", ctypes.c_int32), ("hostThreads", ctypes.c_uint32),
    ]

    @classmethod
    def default(cls, scores=BWA_SCORES, max_read_length=300, device=0, host_threads=0, avoid_smith_waterman=False):
        """The reference's defaults (AlignOptions.cpp:84-133): repeat threshold 10, gapped mismatches 5,
        semialigned gap limit 100, Smith-Waterman always on."""
        return cls(scores[0], scores[1], scores[2], scores[3], scores[4], 10, 8, 5, 100, 1 if avoid_smith_waterman else 0,
                   max_read_length, device, host_threads)


class Adapter(ctypes.Structure):
    """isaac_ext_adapter_t = flowcell::SequencingAdapterMetadata"""
    _fields_ = [("sequence", ctypes.c_char_p), ("reverse", ctypes.c_uint32), ("clipLength", ctypes.c_uint32)]


# flowcell/SequencingAdapterMetadata.cpp:29-39: (sequence, reverse, clipLength; 0 = unbounded)
STANDARD_ADAPTERS = (("AGATCGGAAGAGC", False, 0), ("GCTCTTCCGATCT", True, 0))
NEXTERA_STANDARD_ADAPTERS = (("CTGTCTCTTATACACATCT", False, 0), ("AGATGTGTATAAGAGACAG", True, 0))
NEXTERA_MATEPAIR_ADAPTERS = (("CTGTCTCTTATACACATCT", False, 19), ("AGATGTGTATAAGAGACAG", False, 19))


def adapter_array(adapters):
    """(sequence, reverse, clipLength) tuples -> ctypes array of isaac_ext_adapter_t"""
    arr = (Adapter * max(1, len(adapters)))()
    for i, (seq, reverse, clip) in enumerate(adapters):
        arr[i] = Adapter(seq.encode() if isinstance(seq, str) else seq, 1 if reverse else 0, int(clip))
    return arr


class Reads(ctypes.Structure):
    """isaac_ext_reads_t"""
    _fields_ = [
        ("clusterCount", ctypes.c_uint32), ("readCount", ctypes.c_uint32),
        ("readLength", ctypes.c_uint32 * 2), ("firstCycle", ctypes.c_uint32 * 2),
        ("bcl", ctypes.c_void_p), ("endCyclesMasked", ctypes.c_void_p),
    ]


class ReadSet:
    """Host-side owner of one tile's BCL bytes plus the ctypes view passed over the ABI."""

    def __init__(self, bcl, read_lengths, first_cycles=None, end_cycles_masked=None):
        self.read_lengths = tuple(int(x) for x in read_lengths)
        total = sum(self.read_lengths)
        self.bcl = np.ascontiguousarray(bcl, dtype=np.uint8).reshape(-1, total)
        self.cluster_count = self.bcl.shape[0]
        self.read_count = len(self.read_lengths)
        if first_cycles is None:
            first_cycles, c = [], 1
            for n in self.read_lengths:
                first_cycles.append(c)
                c += n
        self.first_cycles = tuple(int(x) for x in first_cycles)
        self.end_cycles_masked = None
        if end_cycles_masked is not None:
            self.end_cycles_masked = np.ascontiguousarray(end_cycles_masked, dtype=np.uint16).reshape(
                self.cluster_count, self.read_count)
        rl = list(self.read_lengths) + [0] * (2 - self.read_count)
        fc = list(self.first_cycles) + [0] * (2 - self.read_count)
        self.c = Reads(self.cluster_count, self.read_count, (ctypes.c_uint32 * 2)(*rl), (ctypes.c_uint32 * 2)(*fc),
                       self.bcl.ctypes.data,
                       self.end_cycles_masked.ctypes.data if self.end_cycles_masked is not None else None)


def cigar_to_string(words):
    ops = "MIDNSHP=X?"
    return "".join("%d%s" % (int(w) >> 4, ops[min(int(w) & 0xF, 9)]) for w in words)


# isaac_ext_alignment_t (32 bytes): what FragmentBuilder::alignFragments keeps of one candidate
ALIGNMENT_DTYPE = np.dtype([("position", "<i8"), ("logProbability", "<f8"), ("observedLength", "<u2"), ("mismatchCount", "<u2"),
                            ("matchesInARow", "<u2"), ("editDistance", "<u2"), ("smithWatermanScore", "<u2"), ("lowClipped", "<u2"),
                            ("highClipped", "<u2"), ("gapsAndFlags", "u1"), ("cigarLength", "u1")])
ALIGNMENT_GAPS, ALIGNMENT_ALIGNED, ALIGNMENT_GAPPED = 0x3F, 0x40, 0x80
