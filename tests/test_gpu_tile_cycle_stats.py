"""isaac_ext_tile_cycle_stats: matchSelector::TileStats (alignment score histograms of fragments and templates, per-cycle blanks /
mismatches / fragments-with-k-mismatches arrays, each again for uniquely aligned fragments) of a tile's templates on the GPU against
the reference's OWN TileStats fed by its own TemplateBuilder and end clippers (oracle/_ref: oracle_tile_cycle_stats), word for word,
before and after TileStats::finalize."""
import numpy as np
import pytest

import oracle_lib
from common_build import build_workload
from isaac_aligner_b200.batch import DODGY_ALIGNMENT_SCORE_UNALIGNED, DODGY_ALIGNMENT_SCORE_UNKNOWN, Tls, TemplateOptions
from isaac_aligner_b200.types import BWA_SCORES, ELAND_SCORES, Config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


def assert_same(got, want, what):
    for k in range(4):
        at = 0
        for name, size in oracle_lib.TILE_CYCLE_STATS_FIELDS:
            x, y = got[k, at:at + size], want[k, at:at + size]
            bad = np.nonzero(x != y)[0]
            assert not bad.size, "%s: block %d, %s differs at %d places, first [%d]: %d vs %d" % (what, k, name, bad.size, bad[0], x[bad[0]], y[bad[0]])
            at += size
        assert at == oracle_lib.TILE_CYCLE_STATS_WORDS


@pytest.mark.parametrize("scores,L,options,pf_rate", [
    (BWA_SCORES, 100, TemplateOptions.make(), None),
    (BWA_SCORES, 150, TemplateOptions.make(clip_semialigned=True, clip_overlapping=True, mapq_threshold=10), 0.8),
    (ELAND_SCORES, 100, TemplateOptions.make(dodgy=DODGY_ALIGNMENT_SCORE_UNALIGNED, clip_semialigned=True), 0.5),
    (BWA_SCORES, 75, TemplateOptions.make(scatter_repeats=True, dodgy=DODGY_ALIGNMENT_SCORE_UNKNOWN, clip_overlapping=True, mapq_threshold=3), 0.9),
])
def test_tile_cycle_stats_equal_the_references_tilestats(capi, scores, L, options, pf_rate):
    genome, sim, reads, mb = build_workload(n_pairs=5000, L=L, seed=500 + L, indel_rate=6e-3, snp_rate=8e-3)
    cfg = Config.default(scores, max_read_length=2 * L)
    pf = None if pf_rate is None else (np.random.default_rng(L).random(reads.cluster_count) < pf_rate).astype(np.uint8)
    tls = Tls.make()
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    ctx.build_templates(mb, tls, options)
    got = ctx.tile_cycle_stats(pf)
    ref = oracle_lib.Oracle(oracle_lib.REF_SO)
    g = oracle_lib.GenomeHolder(genome)
    want = oracle_lib.tile_cycle_stats(ref, g, reads, cfg, mb, tls, options, pf, threads=8)
    assert_same(got, want, "raw counters, L %d" % L)
    # something was counted everywhere
    for first, size in ((0, 8192), (16384, 8192), (32768, 1024), (34816, 1024), (41984, 1024), (36864, 1024)):
        assert got[0, first:first + size].sum() > 0, first
    assert got[1].sum() < got[0].sum() or pf is None
    assert_same(ctx.tile_cycle_stats(pf, finalize=True), oracle_lib.tile_cycle_stats(ref, g, reads, cfg, mb, tls, options, pf, finalize=True, threads=8),
                "after TileStats::finalize, L %d" % L)
    # the counters need the templates of this tile on the device
    ctx.build_fragments(mb)
    with pytest.raises(capi.ExtError):
        ctx.tile_cycle_stats(pf)
    ctx.close()


def test_single_ended_tile(capi):
    from test_gpu_templates import single_ended
    genome, sim, reads, mb = build_workload(n_pairs=3000, L=100, seed=517, snp_rate=6e-3)
    reads1, mb1 = single_ended(genome, sim, reads, mb)
    cfg = Config.default(max_read_length=200)
    tls, options = Tls.make(), TemplateOptions.make(clip_semialigned=True)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads1)
    ctx.build_templates(mb1, tls, options)
    got = ctx.tile_cycle_stats()
    want = oracle_lib.tile_cycle_stats(oracle_lib.Oracle(oracle_lib.REF_SO), oracle_lib.GenomeHolder(genome), reads1, cfg, mb1, tls, options, threads=4)
    assert_same(got, want, "single-ended")
    assert got[2:].sum() == 0 and got[0].sum() > 0
    ctx.close()
