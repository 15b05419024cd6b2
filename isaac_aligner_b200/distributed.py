"""Multi-GPU plumbing of the candidate-extension path.

The path shards with no data-path collective: clusters (read pairs) are independent once the reference and the template
length statistics are fixed, so rank r of G takes tiles r, r+G, ... (the reference strides clusters over its threads
the same way, MatchSelector.cpp:279-291) with the packed reference replicated per GPU.  The only exchange is the sum
of the per-tile statistics counters (one all-reduce of 64 u64, NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

STATS_COUNTERS = 64   # ISAAC_EXT_STATS_COUNTERS
STAT_NAMES = ("fragments", "aligned", "gapped", "perfect", "mismatches", "editDistance", "gaps", "bases")


def tiles_of_rank(n_tiles, rank, world):
    """tile indices processed by `rank`: r, r + G, r + 2G, ..."""
    return list(range(rank, n_tiles, world))


def cluster_range_of_rank(n_clusters, rank, world):
    """contiguous cluster range of `rank` when one tile is split (balanced to within one cluster)"""
    return n_clusters * rank // world, n_clusters * (rank + 1) // world


def stats_from_fragments(fragments):
    """numpy statement of the K6 counters (kernels_stats.cuh) for host-side checks"""
    s = np.zeros(STATS_COUNTERS, dtype=np.uint64)
    aligned = fragments["cigarLength"] != 0
    a = fragments[aligned]
    s[0] = fragments.size
    s[1] = a.size
    s[2] = int((a["gapCount"] != 0).sum())
    s[3] = int((a["editDistance"] == 0).sum())
    s[4] = int(a["mismatchCount"].astype(np.uint64).sum())
    s[5] = int(a["editDistance"].astype(np.uint64).sum())
    s[6] = int(a["gapCount"].astype(np.uint64).sum())
    s[7] = int(a["observedLength"].astype(np.uint64).sum())
    s[8:41] = np.bincount(np.minimum(a["mismatchCount"], 32), minlength=33).astype(np.uint64)
    return s


TEMPLATE_STATS_COUNTERS = 32   # ISAAC_EXT_TEMPLATE_STATS_COUNTERS: one matchSelector::TileBarcodeStats per (read, pass filter)
TEMPLATE_STAT_NAMES = ("yield", "yieldQ30", "qualityScoreSum", "clusterCount", "unanchoredClusterCount", "nmnmClusterCount",
                       "rmClusterCount", "qcClusterCount", "alignedFragmentCount", "uniquelyAlignedFragmentCount",
                       "uniquelyAlignedPerfectFragmentCount", "alignmentScoreSum", "basesOutsideIndels",
                       "uniquelyAlignedBasesOutsideIndels", "mismatches", "uniquelyAlignedMismatches")
TLS_WORDS = 8   # isaac_ext_tls_t: min, max, median, lowStdDev, highStdDev, bestModel[2], mateDriftRange


def broadcast_tls(tls, src=0, device="cpu"):
    """The reference determines the template length statistics on the first tile and uses them for all later tiles unless
    --per-tile-tls (MatchSelector.cpp:401-417).  With the tiles dealt over ranks, the rank that owns the first tile (src)
    runs isaac_ext_determine_template_length and every other rank receives the eight words of its isaac_ext_tls_t.
    `tls` is a batch.Tls (ignored on the receiving ranks, may be None there); returns the Tls every rank continues with."""
    import ctypes
    import torch
    import torch.distributed as dist
    from .batch import Tls
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tls
    words = torch.zeros(TLS_WORDS, dtype=torch.int32, device=device)
    if dist.get_rank() == src:
        raw = np.frombuffer(ctypes.string_at(ctypes.addressof(tls), ctypes.sizeof(Tls)), dtype=np.int32)
        words.copy_(torch.from_numpy(raw.copy()))
    dist.broadcast(words, src=src)
    out = Tls()
    ctypes.memmove(ctypes.addressof(out), words.cpu().numpy().tobytes(), ctypes.sizeof(Tls))
    return out


def allreduce_stats(stats):
    """sums the counter vector over all ranks in place.  `stats` is a torch int64 tensor (cuda -> NCCL, cpu -> gloo)
    holding the u64 counters bit for bit; returns it.  A single-process run returns it unchanged."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats
