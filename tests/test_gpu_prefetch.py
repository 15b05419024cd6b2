"""isaac_ext_prefetch_reads: the next tile's BCL bytes are uploaded and decoded next to the kernels of the current tile (the
reference loads the next tile while it processes the current one, SelectMatchesTransition.cpp:316-340).  Same templates as without
it, whatever the order of the calls; and the tiles of a run take less time than with the uploads in line."""
import time

import numpy as np
import pytest

from common_build import build_workload
from isaac_aligner_b200.batch import Tls, TemplateOptions
from isaac_aligner_b200.types import Config, ReadSet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


def same(a, b):
    return a.templates.tobytes() == b.templates.tobytes() and a.fragments.tobytes() == b.fragments.tobytes() and np.array_equal(a.cigars, b.cigars)


def test_prefetched_tiles_give_the_same_templates(capi):
    genome, sim, reads_a, mb_a = build_workload(n_pairs=4000, L=100, seed=901)
    # another tile of another size on the same genome
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch
    sim_b = synth.simulate_pairs(genome, 2500, L=100, seed=9021, indel_rate=4e-3, seed_offsets=synth.auto_seed_offsets(100))
    matches_b, begin_b = synth.make_matches(sim_b, genome, seed=9022, decoy_rate=0.3)
    reads_b, mb_b = ReadSet(sim_b.bcl, (100, 100)), MatchBatch(matches_b, begin_b, synth.seed_table(sim_b), with_gaps=True)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    tls, opt = Tls.make(), TemplateOptions.make(clip_semialigned=True)
    ctx.set_reads(reads_a)
    want_a = ctx.build_templates(mb_a, tls, opt)
    # prefetch B under A, take it over, prefetch A under B, take it over
    ctx.prefetch_reads(reads_b)
    ctx.prefetch_batch(mb_b, reads_b.cluster_count)
    again_a = ctx.build_templates(mb_a, tls, opt)
    assert same(want_a, again_a)
    ctx.set_reads(reads_b)
    got_b = ctx.build_templates(mb_b, tls, opt)                                  # finds its matches on the device
    ctx.prefetch_reads(reads_a)
    ctx.prefetch_batch(mb_b, reads_b.cluster_count)                              # a prefetched batch that is not the tile's is ignored
    ctx.set_reads(reads_a)
    assert same(want_a, ctx.build_templates(mb_a, tls, opt))
    assert same(want_a, ctx.build_templates(mb_a, tls, opt))                     # ... and a second call on the tile uploads again
    # a prefetch that is never taken over changes nothing; a plain set_reads of something else still works
    ctx.prefetch_reads(reads_b)
    other = ReadSet(reads_b.bcl.copy(), reads_b.read_lengths)
    ctx.set_reads(other)
    assert same(got_b, ctx.build_templates(mb_b, tls, opt))
    ctx.close()
    fresh = capi.Context(Config.default(max_read_length=200))
    fresh.set_reference(genome)
    fresh.set_reads(reads_b)
    assert same(got_b, fresh.build_templates(mb_b, tls, opt))
    fresh.close()


def test_uploads_overlap_the_previous_tiles_kernels(capi):
    import torch
    genome, sim, reads, mb = build_workload(n_pairs=400_000, L=150, seed=903, genome_bases=3_000_000, masked=False, repeat_rate=0.0, too_many_rate=0.0)
    pinned = torch.empty(reads.bcl.nbytes, dtype=torch.uint8).pin_memory()
    view = pinned.numpy().reshape(reads.bcl.shape)
    view[...] = reads.bcl
    tiles = [ReadSet(view, reads.read_lengths), None]
    pinned2 = torch.empty(reads.bcl.nbytes, dtype=torch.uint8).pin_memory()
    v2 = pinned2.numpy().reshape(reads.bcl.shape)
    v2[...] = reads.bcl
    tiles[1] = ReadSet(v2, reads.read_lengths)
    ctx = capi.Context(Config.default(max_read_length=300))
    ctx.set_reference(genome)
    tls = Tls.make()

    from isaac_aligner_b200.batch import MatchBatch
    keep = []

    def pin(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        keep.append(t)
        return v

    batches = [MatchBatch(pin(mb.matches), pin(mb.begin), mb.seeds, with_gaps=True), MatchBatch(pin(mb.matches), pin(mb.begin), mb.seeds, with_gaps=True)]

    def run(prefetch, count=6):
        ctx.set_reads(tiles[0])
        t0 = time.perf_counter()
        for k in range(count):
            if prefetch:
                ctx.prefetch_reads(tiles[(k + 1) & 1])
                ctx.prefetch_batch(batches[(k + 1) & 1], reads.cluster_count)
            ctx.build_templates(batches[k & 1], tls, copy=False)
            ctx.set_reads(tiles[(k + 1) & 1])
        return (time.perf_counter() - t0) / count

    run(True, 2); run(False, 2)
    inline, overlapped = min(run(False) for _ in range(3)), min(run(True) for _ in range(3))
    print("per tile: uploads in line %.2f ms, prefetched %.2f ms" % (inline * 1e3, overlapped * 1e3))
    # measured on a B200: 22.9 -> 20.4 ms per 400 k pairs (profiles/); the bound here only catches a prefetch that serialises
    assert overlapped < inline * 1.2, (inline, overlapped)
    ctx.close()


def test_deferred_templates_equal_the_blocking_call(capi):
    """isaac_ext_build_templates_deferred / isaac_ext_fetch_templates: the same templates as the blocking call, with the download of
    one tile queued behind its kernels while the next tile is built; two result sets, a third waiting tile is refused"""
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch
    genome, sim, reads_a, mb_a = build_workload(n_pairs=5000, L=100, seed=931)
    sim_b = synth.simulate_pairs(genome, 3000, L=100, seed=9321, indel_rate=4e-3, seed_offsets=synth.auto_seed_offsets(100))
    matches_b, begin_b = synth.make_matches(sim_b, genome, seed=9322, decoy_rate=0.3)
    reads_b, mb_b = ReadSet(sim_b.bcl, (100, 100)), MatchBatch(matches_b, begin_b, synth.seed_table(sim_b), with_gaps=True)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    tls, opt = Tls.make(), TemplateOptions.make(clip_semialigned=True, clip_overlapping=True)
    ctx.set_reads(reads_a)
    want_a = ctx.build_templates(mb_a, tls, opt)
    ctx.set_reads(reads_b)
    want_b = ctx.build_templates(mb_b, tls, opt)
    # A deferred, B deferred on top of it, then both fetched in order; and once more the other way round
    for first, second, want_first, want_second in ((reads_a, reads_b, want_a, want_b), (reads_b, reads_a, want_b, want_a)):
        mb_first, mb_second = (mb_a, mb_b) if first is reads_a else (mb_b, mb_a)
        ctx.set_reads(first)
        h1 = ctx.build_templates_deferred(mb_first, tls, opt)
        ctx.set_reads(second)
        h2 = ctx.build_templates_deferred(mb_second, tls, opt)
        with pytest.raises(capi.ExtError) as e:
            ctx.build_templates_deferred(mb_second, tls, opt)                   # two tiles are waiting already
        assert e.value.code == 4
        assert same(want_first, ctx.fetch_templates(h1))
        with pytest.raises(capi.ExtError):
            ctx.fetch_templates(h1)                                             # fetched already
        assert same(want_second, ctx.fetch_templates(h2))
    # a blocking call between deferred ones, and the device-resident result of the last tile built
    ctx.set_reads(reads_a)
    h = ctx.build_templates_deferred(mb_a, tls, opt)
    ctx.set_reads(reads_b)
    assert same(want_b, ctx.build_templates(mb_b, tls, opt))
    assert same(want_a, ctx.fetch_templates(h))
    ctx.close()
