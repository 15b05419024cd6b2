"""Parity of isaac_ext_build_fragments (FragmentBuilder::build) and isaac_ext_rescue_shadows
(ShadowAligner::rescueShadow) with the CPU oracle, bit-exact, through the C ABI."""
import os

import numpy as np
import pytest

import oracle_lib
from common_build import assert_flat_equal, build_workload, rescue_requests
from isaac_aligner_b200.batch import FFp, FRm, FRp, RFm, RFp, Tls
from isaac_aligner_b200.types import BWA_SCORES, ELAND_SCORES, Config

pytestmark = pytest.mark.gpu


def checkers():
    return oracle_lib.gpu_checkers()        # fails when the reference build did not travel


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


@pytest.mark.parametrize("scores,L,with_gaps,threads", [(BWA_SCORES, 150, True, 0), (ELAND_SCORES, 100, True, 3),
                                                         (BWA_SCORES, 100, False, 1), (BWA_SCORES, 250, True, 0)])
def test_build_fragments_bit_exact(capi, scores, L, with_gaps, threads):
    genome, sim, reads, mb = build_workload(n_pairs=4000, L=L, seed=140 + L, indel_rate=5e-3, with_gaps=with_gaps)
    cfg = Config.default(scores, max_read_length=2 * L, host_threads=threads)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    got = ctx.build_fragments(mb)
    g = oracle_lib.GenomeHolder(genome)
    for chk in checkers():
        want = oracle_lib.build_fragments(chk, g, reads, cfg, mb, threads=8)
        assert_flat_equal(got, want, "build_fragments cuda vs " + chk.kind)
    f = got.fragments
    assert (f["gapCount"] > 0).sum() > 50 and (got.flags == 0).any()
    ctx.close()


def test_build_fragments_empty_and_ragged(capi):
    """no matches at all; a single cluster; clusters whose matches are all filtered"""
    genome, sim, reads, mb = build_workload(n_pairs=64, L=100, seed=7)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    from isaac_aligner_b200.batch import MatchBatch
    empty = MatchBatch(mb.matches[:0], np.zeros(65, dtype=np.uint64), mb.seeds)
    got = ctx.build_fragments(empty)
    assert got.fragments.size == 0 and not got.flags.any() and not got.begin.any()
    g = oracle_lib.GenomeHolder(genome)
    assert_flat_equal(got, oracle_lib.build_fragments(oracle_lib.port(), g, reads, cfg, empty), "empty batch")
    ctx.close()


@pytest.mark.parametrize("scores,L,models", [(BWA_SCORES, 150, (FRp, RFm)), (ELAND_SCORES, 100, (FRp, RFm)),
                                              (BWA_SCORES, 100, (RFp, FRm)), (BWA_SCORES, 100, (FFp, FRp))])
def test_rescue_shadows_bit_exact(capi, scores, L, models):
    genome, sim, reads, mb = build_workload(n_pairs=3000, L=L, seed=160 + L, indel_rate=6e-3)
    cfg = Config.default(scores, max_read_length=2 * L)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    tls = Tls.make(m0=models[0], m1=models[1])
    req = rescue_requests(sim, seed=161)
    got = ctx.rescue_shadows(tls, req)
    g = oracle_lib.GenomeHolder(genome)
    for chk in checkers():
        want = oracle_lib.rescue_shadows(chk, g, reads, cfg, tls, req, threads=8)
        assert_flat_equal(got, want, "rescue_shadows cuda vs " + chk.kind)
    if models == (FRp, RFm):
        assert got.flags.mean() > 0.8 and (got.fragments["gapCount"] > 0).any()
    if models == (FFp, FRp):
        assert not got.flags.any()          # incoherent models: rescuing impossible (ShadowAligner.cpp:164-168)
    ctx.close()


def test_rescue_low_complexity_hits_the_candidate_cap(capi):
    """a poly-A shadow against a poly-A window produces > 10000 scan hits: the 10000-candidate cap and the 1000-shadow
    cap of the reference must cut at the same place (ShadowAligner.cpp:93-97, 212-215)"""
    rng = np.random.default_rng(5)
    L = 100
    contig = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=60000)].copy()
    contig[20000:45000] = ord("A")
    bcl = np.zeros((4, 2 * L), dtype=np.uint8)
    bcl[:, :] = (30 << 2) | 0                 # all A, Q30
    bcl[:, L:] = (30 << 2) | 3                # read 2 all T (reverse strand of poly-A)
    from isaac_aligner_b200.types import ReadSet
    from isaac_aligner_b200.batch import RESCUE_REQUEST_DTYPE
    reads = ReadSet(bcl, (L, L))
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L)
    ctx = capi.Context(cfg)
    ctx.set_reference([contig])
    ctx.set_reads(reads)
    req = np.zeros(4, dtype=RESCUE_REQUEST_DTYPE)
    req["orphanReadId"] = [0, 2, 4, 6]
    req["orphanPosition"] = [21000, 19000, 30000, 44000]
    req["orphanObservedLength"] = L
    req["bestTemplateLength"] = [0, 15000, 30000, 500]
    tls = Tls.make()
    got = ctx.rescue_shadows(tls, req)
    g = oracle_lib.GenomeHolder([contig])
    for chk in checkers():
        want = oracle_lib.rescue_shadows(chk, g, reads, cfg, tls, req, fragments_per_request=1100)
        assert_flat_equal(got, want, "rescue cap cuda vs " + chk.kind)
    assert (np.diff(got.begin.astype(np.int64)) == 1000).any()
    ctx.close()


@pytest.mark.parametrize("indel_rate", [5e-4, 1e-2])
def test_baseline_config0_full_size_bit_exact(capi, indel_rate):
    """BASELINE.json configs[0] at its full size (100 000 simulated 2x150 pairs on the 5 Mbp genome, the generator and seeds
    of bench.py) and the indel-rich variant of configs[3]: every fragment of build + rescue equals both CPU checkers"""
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch
    from isaac_aligner_b200.types import ReadSet
    L, n_pairs = 150, 100_000
    genome = synth.make_genome(5_000_000, n_contigs=1, seed=synth.SEED_G5)
    sim = synth.simulate_pairs(genome, n_pairs, L=L, seed=synth.SEED_READS + 7, indel_rate=indel_rate,
                               seed_offsets=synth.auto_seed_offsets(L))
    matches, begin = synth.make_matches(sim, genome, seed=synth.SEED_READS + 8, decoy_rate=0.2)
    reads = ReadSet(sim.bcl, (L, L))
    mb = MatchBatch(matches, begin, synth.seed_table(sim), with_gaps=True)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    tls = Tls.make()
    built = ctx.build_fragments(mb)
    req = synth.rescue_policy(built.fragments, built.begin, n_pairs)
    rescued = ctx.rescue_shadows(tls, req)
    g = oracle_lib.GenomeHolder(genome)
    for chk in checkers():
        assert_flat_equal(built, oracle_lib.build_fragments(chk, g, reads, cfg, mb, threads=8), "config0 build vs " + chk.kind)
        assert_flat_equal(rescued, oracle_lib.rescue_shadows(chk, g, reads, cfg, tls, req, threads=8),
                          "config0 rescue vs " + chk.kind)
    assert built.flags.mean() > 0.95 and rescued.flags.mean() > 0.8
    assert (built.fragments["gapCount"] > 0).sum() > (2000 if indel_rate > 1e-3 else 200)
    ctx.close()
