"""One tile through the path the way MatchSelector::parallelSelect drives it (MatchSelector.cpp:370-443), on top of the C ABI:
BCL clusters + the tile's match records in, templates + template length statistics + MatchSelectorStats summary out.

    load the tile's matches (raw 16-byte alignment::Match records as io::MatchWriter leaves them, sorted by cluster / location
    / seed, SelectMatchesTransition.cpp:242-254) and its BclClusters buffer
 -> isaac_ext_set_reads                       Read::decodeBcl of every cluster
 -> isaac_ext_trim_low_quality_ends           --base-quality-cutoff (MatchSelector.cpp:300)
 -> isaac_ext_determine_template_length       unless the user gave stable statistics (:401-417); broadcast when tiles are dealt
                                              over several GPUs (distributed.broadcast_tls)
 -> isaac_ext_build_templates                 buildFragments + buildTemplate + end clippers of every cluster (:323-349)
 -> isaac_ext_template_stats                  what threadStats_.recordTemplate collects (:306-358)

The chain itself is the C entry point isaac_ext_select_tile; this module reads the reference's files and calls it."""
import ctypes

import numpy as np

from .batch import MatchBatch, TemplateOptions, TemplateResult, Tls
from .synth import MATCH_DTYPE, SEED_DTYPE
from .types import Reads, ReadSet

CLUSTER_SHIFT, CLUSTER_MASK = 9, (1 << 31) - 1          # SeedId: tile:12 barcode:12 cluster:31 seed:8 reverse:1 (SeedId.hh:37-127)


def read_match_file(path):
    """a match file of the reference: a plain array of 16-byte Match records (Match.hh:38-73, io/MatchWriter.hh)"""
    return np.fromfile(path, dtype=MATCH_DTYPE)


def write_match_file(path, matches):
    np.ascontiguousarray(matches, dtype=MATCH_DTYPE).tofile(path)


def read_bcl_clusters(path, read_lengths):
    """BclClusters (BclClusters.hh:33-124): one byte per base, the reads of a cluster back to back, clusters back to back"""
    total = int(sum(read_lengths))
    return np.fromfile(path, dtype=np.uint8).reshape(-1, total)


def cluster_match_begin(matches, cluster_count):
    """CSR offsets of every cluster's matches from the cluster field of their seed ids (the reference walks the sorted list
    with findNextCluster, MatchSelector.cpp:262-277); clusters without any record get an empty range"""
    cluster = ((matches["seedId"] >> np.uint64(CLUSTER_SHIFT)) & np.uint64(CLUSTER_MASK)).astype(np.int64)
    if cluster.size and (np.any(np.diff(cluster) < 0) or cluster[-1] >= cluster_count):
        raise ValueError("matches must be sorted by cluster and belong to the tile")
    begin = np.zeros(cluster_count + 1, dtype=np.uint64)
    np.cumsum(np.bincount(cluster, minlength=cluster_count), out=begin[1:])
    return begin


class TileResult:
    def __init__(self, templates, tls, tls_stable, stats, end_cycles_masked, packed=None, cycle_stats=None):
        self.templates, self.tls, self.tls_stable, self.stats, self.end_cycles_masked = templates, tls, tls_stable, stats, end_cycles_masked
        self.packed, self.cycle_stats = packed, cycle_stats


class TileC(ctypes.Structure):
    """isaac_ext_tile_t"""
    _fields_ = [("reads", Reads), ("matches", ctypes.c_void_p), ("matchCount", ctypes.c_uint64), ("seeds", ctypes.c_void_p),
                ("seedCount", ctypes.c_uint32), ("withGaps", ctypes.c_uint32), ("pf", ctypes.c_void_p),
                ("baseQualityCutoff", ctypes.c_uint32), ("mateDriftRange", ctypes.c_int32), ("tls", ctypes.c_void_p),
                ("options", TemplateOptions), ("pack", ctypes.c_void_p), ("cycleStats", ctypes.c_uint32), ("pad", ctypes.c_uint32)]


class TileResultC(ctypes.Structure):
    """isaac_ext_tile_result_t"""
    _fields_ = [("templates", TemplateResult), ("tls", Tls), ("tlsStable", ctypes.c_uint32), ("packedValid", ctypes.c_uint32),
                ("stats", ctypes.c_uint64 * 128), ("endCyclesMasked", ctypes.c_void_p), ("cycleStats", ctypes.c_void_p)]


def select_tile(ctx, bcl, read_lengths, matches, seeds, pf=None, base_quality_cutoff=0, tls=None, options=None,
                mate_drift_range=-1, with_gaps=True, pack=None, cycle_stats=False):
    """MatchSelector::parallelSelect for one tile on the context's GPU = ONE call of isaac_ext_select_tile.  tls: user-defined
    template length statistics (batch.Tls) or None = determine them from this tile.  pack: batch.PackOptions = also leave the
    io::FragmentHeader bin records FragmentCollector::add stores for the tile (TileResult.packed), None = skip that pass."""
    from . import capi
    reads = ReadSet(bcl, tuple(read_lengths))
    matches = np.ascontiguousarray(matches, dtype=MATCH_DTYPE)
    seeds = np.ascontiguousarray(seeds, dtype=SEED_DTYPE)
    pf_a = np.ascontiguousarray(pf, dtype=np.uint8) if pf is not None else None
    options = options if options is not None else TemplateOptions.make()
    t = TileC(reads.c, matches.ctypes.data if matches.size else None, matches.size, seeds.ctypes.data, seeds.size, 1 if with_gaps else 0,
              pf_a.ctypes.data if pf_a is not None else None, int(base_quality_cutoff), int(mate_drift_range),
              ctypes.addressof(tls) if tls is not None else None, options, ctypes.addressof(pack.c) if pack is not None else None,
              1 if cycle_stats else 0, 0)
    res = TileResultC()
    ctx._check(capi._lib.isaac_ext_select_tile(ctx._h, ctypes.byref(t), ctypes.byref(res)))
    ctx.reads = reads
    templates = ctx._templates(res.templates)
    out_tls = Tls()
    ctypes.memmove(ctypes.byref(out_tls), ctypes.byref(res.tls), ctypes.sizeof(Tls))
    stats = np.array(res.stats, dtype=np.uint64).reshape(4, 32)
    masked = None
    if res.endCyclesMasked:
        count = reads.cluster_count * reads.read_count
        buf = (ctypes.c_char * (count * 2)).from_address(res.endCyclesMasked)
        masked = np.frombuffer(buf, dtype=np.uint16).copy().reshape(reads.cluster_count, reads.read_count)
    packed = None
    if pack is not None:
        from .batch import PackResultC
        pr = PackResultC()
        ctx._check(capi._lib.isaac_ext_tile_packed(ctx._h, ctypes.byref(pr)))
        packed = ctx._packed(pr, bool(pack.c.compact))
    cycles = None
    if res.cycleStats:
        buf = (ctypes.c_char * (4 * 47105 * 8)).from_address(res.cycleStats)
        cycles = np.frombuffer(buf, dtype=np.uint64).copy().reshape(4, 47105)
    return TileResult(templates, out_tls, bool(res.tlsStable), stats, masked, packed, cycles)
