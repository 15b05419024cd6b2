"""The N > 1 path: tiles / cluster ranges shard over ranks with no data-path collective; the per-tile statistics are
summed with one all-reduce.  The CPU tests run the host logic on a world of 2 gloo ranks (the CPU oracle stands in for
the kernels as the checker); the GPU test checks the K6 counters kernel against its numpy statement."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from common import small_workload
    from isaac_aligner_b200 import distributed
    from isaac_aligner_b200.types import Config
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genome, sim, reads, cand = small_workload(n_pairs=600, L=100, seed=12)
    cfg = Config.default(max_read_length=200)
    g = oracle_lib.GenomeHolder(genome)
    chk = oracle_lib.port()
    # every rank extends only the candidates of its own cluster range
    b, e = distributed.cluster_range_of_rank(reads.cluster_count, rank, world)
    cluster_of = cand["readId"] // 2
    mine = cand[(cluster_of >= b) & (cluster_of < e)]
    frags, _, _ = chk.ungapped(g, reads, cfg, mine)
    local = distributed.stats_from_fragments(frags)
    t = torch.from_numpy(local.view(np.int64).copy())
    distributed.allreduce_stats(t)
    if rank == 0:
        full, _, _ = chk.ungapped(g, reads, cfg, cand)
        want = distributed.stats_from_fragments(full)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(t.numpy().view(np.uint64), want), int(want[0]), int(local[0])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 500
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, total, local = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok == 1 and 0 < local < total


def test_shard_helpers_cover_everything_once():
    from isaac_aligner_b200 import distributed
    for world in (1, 2, 3, 8):
        tiles = sorted(t for r in range(world) for t in distributed.tiles_of_rank(21, r, world))
        assert tiles == list(range(21))
        ranges = [distributed.cluster_range_of_rank(1001, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 1001
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


@pytest.mark.gpu
def test_tile_stats_kernel_matches_numpy():
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import small_workload
    from isaac_aligner_b200 import capi, distributed
    from isaac_aligner_b200.types import Config
    genome, sim, reads, cand = small_workload(n_pairs=2000, L=100, seed=14, indel_rate=5e-3)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    cand = cand[ctx.ungapped(cand)[0]["cigarLength"] > 0]
    frags, _, _ = ctx.gapped(cand)
    d = torch.from_numpy(frags.view(np.uint8).reshape(-1, 64)).cuda()
    stats = torch.zeros(distributed.STATS_COUNTERS, dtype=torch.int64, device="cuda")
    ctx.tile_stats_device(len(frags), d.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(stats.cpu().numpy().view(np.uint64), distributed.stats_from_fragments(frags))
    ctx.close()


def _tls_rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import ctypes
    import torch.distributed as dist
    from isaac_aligner_b200 import distributed
    from isaac_aligner_b200.batch import FRm, RFp, Tls
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 owns the first tile: it determined these statistics; the others start with nothing
    mine = Tls.make(mn=211, mx=498, median=347, low=33, high=36, m0=RFp, m1=FRm, drift=-1) if rank == 0 else None
    got = distributed.broadcast_tls(mine, src=0)
    np.save(os.path.join(out_dir, "tls%d.npy" % rank),
            np.frombuffer(ctypes.string_at(ctypes.addressof(got), ctypes.sizeof(Tls)), dtype=np.int32).copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tls_broadcast(tmp_path):
    """the template length statistics of the first tile reach every rank bit for bit (MatchSelector.cpp:401-417)"""
    import torch.multiprocessing as mp
    port = 30100 + os.getpid() % 500
    mp.spawn(_tls_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (np.load(os.path.join(str(tmp_path), "tls%d.npy" % r)) for r in range(2))
    assert np.array_equal(a, b) and list(a) == [211, 498, 347, 33, 36, 2, 5, -1]


def _template_stats_rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from common_build import build_workload
    from isaac_aligner_b200 import distributed
    from isaac_aligner_b200.batch import MatchBatch, TemplateOptions, Tls
    from isaac_aligner_b200.types import BWA_SCORES, Config, ReadSet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chk = oracle_lib.reference()
    genome, sim, reads, mb = build_workload(n_pairs=1200, L=100, seed=33, masked=False)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    g = oracle_lib.GenomeHolder(genome)
    tls, options = Tls.make(), TemplateOptions.make()
    # every rank takes its contiguous cluster range of the tile (the statistics do not depend on how the tile is cut)
    b, e = distributed.cluster_range_of_rank(reads.cluster_count, rank, world)
    sub_reads = ReadSet(reads.bcl[b:e], reads.read_lengths)
    begin = mb.begin[b:e + 1] - mb.begin[b]
    sub_mb = MatchBatch(mb.matches[int(mb.begin[b]):int(mb.begin[e])], begin, mb.seeds)
    mine = oracle_lib.template_stats(chk, g, sub_reads, cfg, sub_mb, tls, options)
    t = torch.from_numpy(mine.view(np.int64).copy())
    distributed.allreduce_stats(t)
    if rank == 0:
        whole = oracle_lib.template_stats(chk, g, reads, cfg, mb, tls, options)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(t.numpy().view(np.uint64), whole), int(whole[0][3]), int(mine[0][3])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_template_stats_allreduce(tmp_path):
    """the MatchSelectorStats summary of a tile = the sum of the summaries of its parts: two gloo ranks, each with its cluster
    range, one all-reduce (the CPU checker stands in for the kernel as the producer of the per-rank vectors)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    chk = oracle_lib.reference()
    if chk is None or not hasattr(chk.lib, "oracle_template_stats"):
        pytest.skip("needs the reference build of the checker")
    import torch.multiprocessing as mp
    port = 30700 + os.getpid() % 500
    mp.spawn(_template_stats_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, total, local = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok == 1 and total == 1200 and local == 600


def _rank_tile_stats(rank, world, port, out_dir):
    """the full matchSelector::TileStats payload (4 x 47105 u64 + the 4 x 32 summary counters) of two half tiles, summed over two
    gloo ranks, against the whole tile's: the reference's own TileStats / TileBarcodeStats stand in for the kernels"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from common_build import build_workload
    from isaac_aligner_b200 import distributed
    from isaac_aligner_b200.batch import MatchBatch, Tls, TemplateOptions
    from isaac_aligner_b200.types import Config, ReadSet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genome, sim, reads, mb = build_workload(n_pairs=900, L=100, seed=77, masked=False)
    cfg, tls, options = Config.default(max_read_length=200), Tls.make(), TemplateOptions.make(clip_semialigned=True)
    ref = oracle_lib.reference()
    g = oracle_lib.GenomeHolder(genome)

    def payload(b, e):
        sub_reads = ReadSet(reads.bcl[b:e], reads.read_lengths)
        m0, m1 = int(mb.begin[b]), int(mb.begin[e])
        sub_mb = MatchBatch(mb.matches[m0:m1], mb.begin[b:e + 1] - mb.begin[b], mb.seeds, with_gaps=True)
        summary = oracle_lib.template_stats(ref, g, sub_reads, cfg, sub_mb, tls, options)
        cycles = oracle_lib.tile_cycle_stats(ref, g, sub_reads, cfg, sub_mb, tls, options)
        return np.concatenate([summary.reshape(-1), cycles.reshape(-1)])

    b, e = distributed.cluster_range_of_rank(reads.cluster_count, rank, world)
    t = torch.from_numpy(payload(b, e).view(np.int64).copy())
    distributed.allreduce_stats(t)
    if rank == 0:
        want = payload(0, reads.cluster_count)
        got = t.numpy().view(np.uint64)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(got, want), got.size, int(want[128 + 34816:128 + 35840].sum())]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/c++") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libisaac_ref.so")),
                    reason="TileStats has the reference build as its only producer on the CPU")
def test_two_rank_full_tile_stats_payload(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + 17) % 500
    mp.spawn(_rank_tile_stats, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, size, mismatches = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok == 1 and size == 4 * 32 + 4 * 47105 and mismatches > 0


def _bins_rank_main(rank, world, port, out_dir):
    """the gap realigner's sharding: the bins of a run are dealt over the ranks like tiles (bin r, r + G, ...), every rank realigns its
    own against its replica of the reference, nothing of the bins is exchanged; only the count of realigned fragments is summed here,
    as a run's statistics would be.  The reference's own GapRealigner stands in for the kernels."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from isaac_aligner_b200 import bins, distributed
    from isaac_aligner_b200.batch import Tls
    from test_realign_host import make_contigs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = oracle_lib.reference()
    contigs = make_contigs(3, lengths=(3000, 60000))
    genome = oracle_lib.GenomeHolder(contigs)
    n_bins = 5
    regions = [(2000 + 9000 * k, 2000 + 9000 * (k + 1)) for k in range(n_bins)]

    def realigned_in(k):
        b = bins.simulate_bin(contigs, contig=1, region=regions[k], n_pairs=500, read_length=100, seed=60 + k)
        o = bins.RealignOptions(b.bin_start, b.bin_end, [Tls.make()], clip_semialigned=True)
        res = oracle_lib.realign_bin(ref, genome, b, o)
        return int(np.count_nonzero(res.cigar_offset != bins.OWN_CIGAR))

    mine = distributed.tiles_of_rank(n_bins, rank, world)
    local = sum(realigned_in(k) for k in mine)
    t = torch.tensor([local, len(mine)], dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        want = sum(realigned_in(k) for k in range(n_bins))
        np.save(os.path.join(out_dir, "bins.npy"), np.array([int(t[0]), want, int(t[1]), local]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bins_of_the_gap_realigner(tmp_path):
    import oracle_lib
    if oracle_lib.reference() is None:
        pytest.skip("oracle/_ref/libisaac_ref.so not built (needs /root/reference)")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + 77) % 500
    mp.spawn(_bins_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    total, want, bins_done, local = np.load(os.path.join(str(tmp_path), "bins.npy"))
    assert total == want and want > 0 and bins_done == 5 and 0 < local < total
