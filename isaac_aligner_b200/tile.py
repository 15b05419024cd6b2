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

This module is host plumbing only; every step is a call into libisaac_ext.so."""
import numpy as np

from .batch import MatchBatch, TemplateOptions
from .synth import MATCH_DTYPE
from .types import ReadSet

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
    def __init__(self, templates, tls, tls_stable, stats, end_cycles_masked, packed=None):
        self.templates, self.tls, self.tls_stable, self.stats, self.end_cycles_masked = templates, tls, tls_stable, stats, end_cycles_masked
        self.packed = packed


def select_tile(ctx, bcl, read_lengths, matches, seeds, pf=None, base_quality_cutoff=0, tls=None, options=None,
                mate_drift_range=-1, with_gaps=True, pack=None):
    """MatchSelector::parallelSelect for one tile on the context's GPU.  tls: user-defined template length statistics
    (batch.Tls) or None = determine them from this tile.  pack: batch.PackOptions = also leave the io::FragmentHeader bin
    records FragmentCollector::add stores for the tile (TileResult.packed), None = skip that pass."""
    reads = ReadSet(bcl, tuple(read_lengths))
    ctx.set_reads(reads)
    masked = ctx.trim_low_quality_ends(base_quality_cutoff) if base_quality_cutoff else None
    mb = MatchBatch(matches, cluster_match_begin(matches, reads.cluster_count), seeds, with_gaps=with_gaps)
    stable = True
    if tls is None:
        tls, stable = ctx.determine_template_length(mb, pf, mate_drift_range)
    options = options if options is not None else TemplateOptions.make()
    templates = ctx.build_templates(mb, tls, options)
    stats = ctx.template_stats(mb, tls, templates, pf)
    packed = ctx.pack_fragments(templates, pack) if pack is not None else None
    return TileResult(templates, tls, stable, stats, masked, packed)
