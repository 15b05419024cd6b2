"""ctypes / numpy mirrors of the batch structs of the two TemplateBuilder-facing calls (include/isaac_ext.h):
isaac_ext_build_fragments and isaac_ext_rescue_shadows."""
import ctypes

import numpy as np

from .synth import MATCH_DTYPE, SEED_DTYPE
from .types import FRAGMENT_DTYPE

# isaac_ext_rescue_request_t, 32 bytes
RESCUE_REQUEST_DTYPE = np.dtype([("orphanPosition", "<i8"), ("bestTemplateLength", "<i8"), ("orphanReadId", "<u4"),
                                 ("orphanContigStrand", "<u4"), ("orphanObservedLength", "<u4"), ("pad", "<u4")])
assert RESCUE_REQUEST_DTYPE.itemsize == 32

# TemplateLengthStatistics::AlignmentModel (TemplateLengthStatistics.hh:48-59)
FFp, FRp, RFp, RRp, FFm, FRm, RFm, RRm = range(8)


class BuildBatch(ctypes.Structure):
    """isaac_ext_build_batch_t"""
    _fields_ = [("matches", ctypes.c_void_p), ("clusterMatchBegin", ctypes.c_void_p), ("seeds", ctypes.c_void_p),
                ("seedCount", ctypes.c_uint32), ("withGaps", ctypes.c_uint32)]


class BuildResult(ctypes.Structure):
    """isaac_ext_build_result_t"""
    _fields_ = [("fragments", ctypes.c_void_p), ("readFragmentBegin", ctypes.c_void_p), ("cigars", ctypes.c_void_p),
                ("built", ctypes.c_void_p), ("fragmentCount", ctypes.c_uint64), ("cigarWords", ctypes.c_uint64)]


class Tls(ctypes.Structure):
    """isaac_ext_tls_t"""
    _fields_ = [("min", ctypes.c_uint32), ("max", ctypes.c_uint32), ("median", ctypes.c_uint32),
                ("lowStdDev", ctypes.c_uint32), ("highStdDev", ctypes.c_uint32), ("bestModel", ctypes.c_uint32 * 2),
                ("mateDriftRange", ctypes.c_int32)]

    @classmethod
    def make(cls, mn=245, mx=455, median=350, low=35, high=35, m0=FRp, m1=RFm, drift=-1):
        """explicit template length statistics min 245, median 350, max 455, FRp / RFm"""
        return cls(mn, mx, median, low, high, (ctypes.c_uint32 * 2)(m0, m1), drift)


class RescueResult(ctypes.Structure):
    """isaac_ext_rescue_result_t"""
    _fields_ = [("fragments", ctypes.c_void_p), ("requestFragmentBegin", ctypes.c_void_p), ("cigars", ctypes.c_void_p),
                ("rescued", ctypes.c_void_p), ("fragmentCount", ctypes.c_uint64), ("cigarWords", ctypes.c_uint64)]


class MatchBatch:
    """Host-side owner of a tile's matches, CSR offsets and seed table plus the ctypes view."""

    def __init__(self, matches, cluster_match_begin, seeds, with_gaps=True):
        self.matches = np.ascontiguousarray(matches, dtype=MATCH_DTYPE)
        self.begin = np.ascontiguousarray(cluster_match_begin, dtype=np.uint64)
        self.seeds = np.ascontiguousarray(seeds, dtype=SEED_DTYPE)
        self.c = BuildBatch(self.matches.ctypes.data, self.begin.ctypes.data, self.seeds.ctypes.data,
                            len(self.seeds), 1 if with_gaps else 0)


class FlatFragments:
    """fragments + CSR + cigar pool copied out of a result struct (or filled by the oracle)."""

    def __init__(self, fragments, begin, cigars, flags):
        self.fragments, self.begin, self.cigars, self.flags = fragments, begin, cigars, flags

    def cigar(self, i):
        f = self.fragments[i]
        return self.cigars[int(f["cigarOffset"]):int(f["cigarOffset"]) + int(f["cigarLength"])]


def copy_result(res, n_groups, begin_field, flag_field, n_flags):
    nf, nc = int(res.fragmentCount), int(res.cigarWords)

    def arr(ptr, dtype, count):
        if not count:
            return np.zeros(0, dtype=dtype)
        buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).copy()

    return FlatFragments(arr(res.fragments, FRAGMENT_DTYPE, nf), arr(getattr(res, begin_field), np.uint64, n_groups + 1),
                         arr(res.cigars, np.uint32, nc), arr(getattr(res, flag_field), np.uint8, n_flags))


# isaac_ext_template_t, 16 bytes
TEMPLATE_DTYPE = np.dtype([("alignmentScore", "<u4"), ("fragmentAlignmentScore", "<u4", (2,)), ("properPair", "u1"),
                           ("built", "u1"), ("hadFragments", "u1"), ("pad", "u1")])
assert TEMPLATE_DTYPE.itemsize == 16

DODGY_ALIGNMENT_SCORE_UNKNOWN, DODGY_ALIGNMENT_SCORE_UNALIGNED = 255, -1      # TemplateBuilder.hh:60-61
CLIP_SEMIALIGNED, CLIP_OVERLAPPING = 1, 2                                     # ISAAC_EXT_CLIP_*


class TemplateOptions(ctypes.Structure):
    """isaac_ext_template_options_t (isaac-align defaults: --scatter-repeats 0, --dodgy-alignment-score 0, --mapq-threshold 0)"""
    _fields_ = [("scatterRepeats", ctypes.c_uint32), ("dodgyAlignmentScore", ctypes.c_int32),
                ("mapqThreshold", ctypes.c_uint32), ("clipFlags", ctypes.c_uint32)]

    @classmethod
    def make(cls, scatter_repeats=False, dodgy=0, mapq_threshold=0, clip_semialigned=False, clip_overlapping=False):
        return cls(1 if scatter_repeats else 0, dodgy, mapq_threshold,
                   (CLIP_SEMIALIGNED if clip_semialigned else 0) | (CLIP_OVERLAPPING if clip_overlapping else 0))


class TemplateResult(ctypes.Structure):
    """isaac_ext_template_result_t"""
    _fields_ = [("templates", ctypes.c_void_p), ("fragments", ctypes.c_void_p), ("cigars", ctypes.c_void_p),
                ("cigarWords", ctypes.c_uint64), ("rescueRequests", ctypes.c_uint64)]


class Templates:
    """templates[cluster], fragments[cluster * readCount + readIndex] and their CIGAR pool"""

    def __init__(self, templates, fragments, cigars, rescue_requests=0):
        self.templates, self.fragments, self.cigars, self.rescue_requests = templates, fragments, cigars, rescue_requests

    def cigar(self, i):
        f = self.fragments[i]
        return self.cigars[int(f["cigarOffset"]):int(f["cigarOffset"]) + int(f["cigarLength"])]


class PackOptionsC(ctypes.Structure):
    """isaac_ext_pack_options_t"""
    _fields_ = [("tile", ctypes.c_uint64), ("barcodeIdx", ctypes.c_uint32), ("keepUnaligned", ctypes.c_uint32),
                ("compact", ctypes.c_uint32), ("pad", ctypes.c_uint32), ("pf", ctypes.c_void_p), ("xy", ctypes.c_void_p), ("barcodeSequence", ctypes.c_void_p),
                ("distributionBinSize", ctypes.c_uint32), ("contigCount", ctypes.c_uint32),
                ("contigBinBegin", ctypes.c_void_p), ("binIndex", ctypes.c_void_p)]


class PackOptions:
    """What FragmentCollector::add reads beyond the template (isaac_ext_pack_options_t); keeps the arrays the struct points to.
    bin_index: one uint32 array per contig = BinIndexMap::at(contigId + 1) (BinIndexMap.hh:45-107)."""

    def __init__(self, tile=0, barcode_idx=0, keep_unaligned=False, pf=None, xy=None, barcode_sequence=None,
                 distribution_bin_size=0, bin_index=None, compact=False):
        self.pf = None if pf is None else np.ascontiguousarray(pf, dtype=np.uint8)
        self.xy = None if xy is None else np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        self.barcode_sequence = None if barcode_sequence is None else np.ascontiguousarray(barcode_sequence, dtype=np.uint64)
        self.bin_begin = self.bin_flat = None
        contigs = 0
        if distribution_bin_size:
            contigs = len(bin_index)
            self.bin_begin = np.zeros(contigs + 1, dtype=np.uint64)
            self.bin_begin[1:] = np.cumsum([len(b) for b in bin_index])
            self.bin_flat = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.uint32) for b in bin_index] + [np.zeros(1, np.uint32)]))
        ptr = lambda a: a.ctypes.data if a is not None else None
        self.c = PackOptionsC(tile, barcode_idx, 1 if keep_unaligned else 0, 1 if compact else 0, 0, ptr(self.pf), ptr(self.xy), ptr(self.barcode_sequence),
                              distribution_bin_size, contigs, ptr(self.bin_begin), ptr(self.bin_flat))


class PackResultC(ctypes.Structure):
    """isaac_ext_pack_result_t"""
    _fields_ = [("records", ctypes.c_void_p), ("fStrandPos", ctypes.c_void_p), ("initialized", ctypes.c_void_p),
                ("recordLength", ctypes.c_uint32), ("readOffset", ctypes.c_uint32 * 2), ("headerLength", ctypes.c_uint32),
                ("storedFragments", ctypes.c_uint64), ("recordOffset", ctypes.c_void_p), ("recordBytes", ctypes.c_uint64),
                ("kernelMs", ctypes.c_float), ("pad", ctypes.c_uint32)]


# io::FragmentHeader as g++ lays it out on x86-64 (Fragment.hh:260-404), 112 bytes; flags: bit 0 paired, 1 unmapped, 2 mateUnmapped,
# 3 reverse, 4 mateReverse, 5 firstRead, 6 secondRead, 7 failFilter, 8 properPair, 9 duplicate
FRAGMENT_HEADER_DTYPE = np.dtype([
    ("bamTlen", "<i4"), ("observedLength", "<u4"), ("fStrandPosition", "<u8"), ("lowClipped", "<u2"), ("highClipped", "<u2"),
    ("alignmentScore", "<u2"), ("templateAlignmentScore", "<u2"), ("mateFStrandPosition", "<u8"), ("readLength", "<u2"),
    ("cigarLength", "<u2"), ("gapCount", "<u2"), ("editDistance", "<u2"), ("flags", "<u2"), ("pad0", "<u2", (3,)), ("tile", "<u8"),
    ("barcode", "<u8"), ("barcodeSequence", "<u8"), ("clusterId", "<u8"), ("clusterX", "<i4"), ("clusterY", "<i4"),
    ("duplicateClusterRank", "<u8"), ("mateAnchor", "<u8"), ("mateStorageBin", "<u4"), ("pad1", "<u4")])
assert FRAGMENT_HEADER_DTYPE.itemsize == 112
NO_MATCH_POSITION = 0x7FFFFF << 41                                            # ReferencePosition(NoMatch).getValue()


def reference_position(value):
    """ReferencePosition::getValue() -> (contigId, position) (ReferencePosition.hh:68-78,105-111); contig 0x7FFFFE + 1 = no match"""
    value = np.asarray(value, dtype=np.uint64)
    return ((value >> np.uint64(41)).astype(np.int64) - 1), ((value >> np.uint64(1)) & np.uint64((1 << 40) - 1)).astype(np.int64)


class PackedFragments:
    """matchSelector::FragmentBuffer of a tile: records [clusters, recordLength] bytes, f_strand_pos / initialized
    [clusters, readCount]; the record of (cluster, readIndex) starts at records[cluster, read_offset[readIndex]].
    Compact results: records is flat, the record of (cluster, readIndex) is records[record_offset[i] : record_offset[i + 1]] with
    i = cluster * readCount + readIndex."""

    def __init__(self, records, f_strand_pos, initialized, record_length, read_offset, header_length, stored, record_offset=None):
        self.records, self.f_strand_pos, self.initialized = records, f_strand_pos, initialized
        self.record_length, self.read_offset, self.header_length, self.stored = record_length, tuple(read_offset), header_length, stored
        self.record_offset = record_offset

    def headers(self, read_index):
        """the io::FragmentHeader of every cluster's record of one read (FragmentBuffer layout) as a structured array"""
        begin = self.read_offset[read_index]
        return np.ascontiguousarray(self.records[:, begin:begin + self.header_length]).view(FRAGMENT_HEADER_DTYPE).reshape(-1)

    def compacted(self):
        """the compact form of a FragmentBuffer-layout result: every initialised record cut to FragmentHeader::getTotalLength()
        (112 + readLength_ + 4 * cigarLength_, Fragment.hh:192-202), back to back in (cluster, readIndex) order -> (bytes, offsets)"""
        n, rc = self.initialized.shape
        parts, offsets = [], [0]
        for c in range(n):
            for r in range(rc):
                length = 0
                if self.initialized[c, r]:
                    rec = self.records[c, self.read_offset[r]:]
                    length = self.header_length + int(rec[32:34].view(np.uint16)[0]) + 4 * int(rec[34:36].view(np.uint16)[0])
                    parts.append(rec[:length])
                offsets.append(offsets[-1] + length)
        return (np.concatenate(parts) if parts else np.zeros(0, np.uint8)), np.array(offsets, dtype=np.uint64)


def expected_alignments(ungapped, ungapped_cigars, gapped, gapped_cigars, cigar_stride, read_length, gapped_mismatches_max=5, band=16, cutoff=5):
    """What isaac_ext_align_batch_packed must return for candidates whose ungapped / gapped records (FRAGMENT_DTYPE, CIGARs at
    strides 3 / cigar_stride) are known: FragmentBuilder::alignFragments' per-candidate decision (FragmentBuilder.cpp:190-209) in
    numpy; one read length.  Returns (ALIGNMENT_DTYPE array, pool words).  Test / bench helper."""
    from .types import ALIGNMENT_ALIGNED, ALIGNMENT_DTYPE, ALIGNMENT_GAPPED
    u, g = ungapped, gapped
    aligned = u["cigarLength"] > 0
    d = u["logProbability"] - g["logProbability"]
    lp_less = ~(np.abs(d) <= 0.0000001) & (u["logProbability"] < g["logProbability"])
    observed = np.where(aligned, u["observedLength"], 0).astype(np.int64)
    accept = (aligned & (u["mismatchCount"] > cutoff) & (g["matchCount"] > 0) & (g["matchCount"].astype(np.int64) + band > observed) &
              (g["mismatchCount"] <= gapped_mismatches_max) & (u["mismatchCount"] > g["mismatchCount"]) & lp_less)
    out = np.zeros(len(u), dtype=ALIGNMENT_DTYPE)
    for name in ("position", "logProbability", "observedLength", "mismatchCount", "matchesInARow", "editDistance", "smithWatermanScore",
                 "lowClipped", "highClipped"):
        out[name] = np.where(accept, g[name], u[name])
    gaps = np.where(accept, g["gapCount"], u["gapCount"]).astype(np.uint8)
    out["gapsAndFlags"] = gaps | np.where(accept, ALIGNMENT_ALIGNED | ALIGNMENT_GAPPED, np.where(aligned, ALIGNMENT_ALIGNED, 0)).astype(np.uint8)
    # the implied CIGAR of a kept ungapped alignment against its real one
    cu = np.asarray(ungapped_cigars).reshape(-1, 3).astype(np.int64)
    left = np.where(u["reverse"] != 0, u["highClipped"], u["lowClipped"]).astype(np.int64)
    right = np.where(u["reverse"] != 0, u["lowClipped"], u["highClipped"]).astype(np.int64)
    mid = read_length - left - right
    w_left, w_mid, w_right = left << 4 | 4, mid << 4, right << 4 | 4
    implied = np.zeros((len(u), 3), dtype=np.int64)
    count = np.zeros(len(u), dtype=np.int64)
    for present, word in ((left > 0, w_left), (mid > 0, w_mid), (right > 0, w_right)):
        implied[np.arange(len(u)), np.minimum(count, 2)] = np.where(present, word, implied[np.arange(len(u)), np.minimum(count, 2)])
        count += present
    used = np.arange(3)[None, :] < u["cigarLength"].astype(np.int64)[:, None]
    same = (count == u["cigarLength"]) & np.all(np.where(used, implied == cu, True), axis=1)
    explicit = aligned & ~accept & ~same
    out["cigarLength"] = np.where(accept, g["cigarLength"], np.where(explicit, u["cigarLength"], 0))
    table = np.asarray(gapped_cigars).reshape(-1, cigar_stride)
    rows = np.nonzero(accept | explicit)[0]
    words = np.zeros((len(rows), cigar_stride), dtype=np.uint32)
    is_gapped = accept[rows]
    words[is_gapped] = table[rows[is_gapped]]
    words[~is_gapped, :3] = cu[rows[~is_gapped]]
    lens = out["cigarLength"][rows].astype(np.int64)
    pool = words[np.arange(cigar_stride)[None, :] < lens[:, None]]
    return out, pool.astype(np.uint32)


def implied_ungapped_cigar(alignment, reverse, read_length):
    """the CIGAR words of a kept ungapped alignment (see isaac_ext_alignment_t)"""
    left = int(alignment["highClipped"] if reverse else alignment["lowClipped"])
    right = int(alignment["lowClipped"] if reverse else alignment["highClipped"])
    words = []
    if left:
        words.append(left << 4 | 4)
    if read_length - left - right:
        words.append((read_length - left - right) << 4)
    if right:
        words.append(right << 4 | 4)
    return words
