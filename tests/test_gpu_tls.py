"""isaac_ext_determine_template_length (MatchSelector::determineTemplateLength) against the reference's own
TemplateBuilder::buildFragments + TemplateLengthDistribution driven by the restated loop of oracle/ref_capi.cpp."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib
from common_build import build_workload
from isaac_aligner_b200 import synth
from isaac_aligner_b200.batch import MatchBatch, Tls
from isaac_aligner_b200.types import BWA_SCORES, Config, ReadSet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


def reference_checker():
    if not os.path.exists(oracle_lib.REF_SO):
        pytest.skip("the reference build of the checker did not travel to this box")
    chk = oracle_lib.Oracle(oracle_lib.REF_SO)
    if not hasattr(chk.lib, "oracle_determine_template_length"):
        pytest.skip("reference checker without oracle_determine_template_length")
    return chk


def swap_reads(sim, reads, mb, seed=3):
    """a real library sequences either strand of a fragment first: in half of the clusters read 1 and read 2 trade places (BCL
    bytes, masking, and the seeds their matches belong to), which turns their FR+ pairs into RF- pairs"""
    rng = np.random.default_rng(seed)
    n, L, S = sim.bcl.shape[0], sim.L, len(sim.seed_offsets)
    flip = rng.random(n) < 0.5
    bcl = sim.bcl.copy()
    bcl[flip] = np.concatenate([sim.bcl[flip, L:], sim.bcl[flip, :L]], axis=1)
    ecm = reads.end_cycles_masked
    if ecm is not None:
        ecm = ecm.copy()
        ecm[flip] = ecm[flip][:, ::-1]
    counts = np.diff(mb.begin.astype(np.int64))
    flip_match = np.repeat(flip, counts)
    matches = mb.matches.copy()
    seed_index = (matches["seedId"] >> np.uint64(1)) & np.uint64(0xFF)
    swapped = (seed_index + np.uint64(S)) % np.uint64(2 * S)
    cleared = matches["seedId"] & ~np.uint64(0xFF << 1)
    matches["seedId"] = np.where(flip_match, cleared | (swapped << np.uint64(1)), matches["seedId"])
    return ReadSet(bcl, (L, L), end_cycles_masked=ecm), MatchBatch(matches, mb.begin, mb.seeds)


def words(tls):
    return list(np.frombuffer(ctypes.string_at(ctypes.addressof(tls), ctypes.sizeof(Tls)), dtype=np.int32))


@pytest.mark.parametrize("n_pairs,drift,with_pf,fixed_insert", [(60000, -1, False, False), (3000, -1, True, False),
                                                                   (26000, 40, True, False), (40000, -1, False, True)])
def test_template_length_statistics_bit_exact(capi, n_pairs, drift, with_pf, fixed_insert):
    """tiles that end before the statistics are stable (finalize), filtered clusters, a mate drift range, and a library of one
    single insert size, whose statistics are stable at the second update: the walk stops there"""
    chk = reference_checker()
    extra = dict(insert=(350, 0, 350, 350), indel_rate=0.0) if fixed_insert else dict(indel_rate=2e-3)
    genome, sim, reads, mb = build_workload(n_pairs=n_pairs, L=100, seed=400 + n_pairs % 97, genome_bases=2_000_000, **extra)
    reads, mb = swap_reads(sim, reads, mb)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    pf = (np.random.default_rng(5).random(n_pairs) < 0.9).astype(np.uint8) if with_pf else None
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    got, stable = ctx.determine_template_length(mb, pf, drift)
    want, want_stable = oracle_lib.determine_template_length(chk, oracle_lib.GenomeHolder(genome), reads, cfg, mb, pf, drift)
    assert words(got) == words(want) and stable == want_stable, (words(got), words(want), stable, want_stable)
    assert stable == fixed_insert
    assert 300 < got.median < 400 and got.min <= got.median <= got.max      # the simulated inserts are N(350, 35)
    assert sorted([got.bestModel[0], got.bestModel[1]]) == [1, 6]            # FRp / RFm
    # the statistics drive the template builder like user-supplied ones
    from isaac_aligner_b200.batch import TemplateOptions
    t = ctx.build_templates(mb, got, TemplateOptions.make())
    assert t.templates["properPair"].mean() > 0.7
    ctx.close()


def test_template_length_statistics_single_ended(capi):
    chk = reference_checker()
    genome, sim, reads, mb = build_workload(n_pairs=500, L=100, seed=77)
    single = ReadSet(sim.bcl[:, :100], (100,))
    read_of = ((mb.matches["seedId"] >> np.uint64(1)) & np.uint64(0xFF)) // np.uint64(len(sim.seed_offsets))
    keep = read_of == 0
    counts = np.diff(mb.begin.astype(np.int64))
    cluster_of = np.repeat(np.arange(500), counts)
    begin = np.zeros(501, dtype=np.uint64)
    np.cumsum(np.bincount(cluster_of[keep], minlength=500), out=begin[1:])
    seeds = synth.seed_table(sim)
    mb1 = MatchBatch(mb.matches[keep], begin, seeds[seeds["readIndex"] == 0])
    cfg = Config.default(BWA_SCORES, max_read_length=100)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(single)
    got, stable = ctx.determine_template_length(mb1)
    want, want_stable = oracle_lib.determine_template_length(chk, oracle_lib.GenomeHolder(genome), single, cfg, mb1)
    assert words(got) == words(want) and not stable and not want_stable
    assert got.min == 0xFFFFFFFF and got.bestModel[0] == 8
    ctx.close()


@pytest.mark.parametrize("clip,with_pf", [(False, False), (True, True)])
def test_template_stats_bit_exact(capi, clip, with_pf):
    """the TileBarcodeStats MatchSelector keeps per (read, pass filter) over the templates of a tile, against the reference's own
    TileBarcodeStats fed by the reference's TemplateBuilder; clusters without matches, with NoMatch / N-seed records and with
    match lists that do not build are in the tile"""
    from isaac_aligner_b200.batch import TemplateOptions
    if not hasattr(reference_checker().lib, "oracle_template_stats"):
        pytest.skip("reference checker without oracle_template_stats")
    chk = reference_checker()
    n = 6000
    genome, sim, reads, mb = build_workload(n_pairs=n, L=100, seed=431, genome_bases=1_000_000, indel_rate=4e-3)
    reads, mb = swap_reads(sim, reads, mb)
    # a few clusters whose match list is a single NoMatch record, half of them with the N-seed id (all seeds contain Ns)
    matches, begin = mb.matches.copy(), mb.begin.astype(np.int64)
    rng = np.random.default_rng(8)
    for c in rng.choice(n, size=60, replace=False):
        if begin[c + 1] > begin[c]:
            m = int(begin[c])
            matches["location"][m] = np.uint64(0xFFFFFE0000000000)                       # ReferencePosition::NoMatch
            if c % 2:
                matches["seedId"][m] |= np.uint64(0xFF << 1)                              # SeedId::setNSeedId
    mb = MatchBatch(matches, mb.begin, mb.seeds)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    pf = (rng.random(n) < 0.85).astype(np.uint8) if with_pf else None
    options = TemplateOptions.make(clip_semialigned=clip, clip_overlapping=clip)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    tls, _ = ctx.determine_template_length(mb, pf)
    templates = ctx.build_templates(mb, tls, options)
    got = ctx.template_stats(mb, tls, templates, pf)
    want = oracle_lib.template_stats(chk, oracle_lib.GenomeHolder(genome), reads, cfg, mb, tls, options, pf, threads=8)
    assert np.array_equal(got, want), (got[:, :30], want[:, :30])
    all_clusters = got[0]                                                                  # read 1, all clusters
    assert all_clusters[3] == n and all_clusters[5] > 0 and all_clusters[7] > 0
    assert all_clusters[16 + 1] + all_clusters[16 + 6] > 0.5 * n and all_clusters[25 + 2] > 0.5 * n     # FRp / RFm, nominal
    assert got[2][29] == n and (got[1][3] == n) == (pf is None)
    ctx.close()
