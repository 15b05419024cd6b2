"""Pins the restatement of the sequencing-adapter clipper (FragmentSequencingAdapterClipper + SequencingAdapter) and of
GappedAligner::makesSenseToGapAlign (--avoid-smith-waterman, a9) in oracle/isaac_oracle.cpp against the reference's own classes
(oracle/_ref/libisaac_ref.so) on seeded workloads, and replays the reference's adapter unit-test vectors through the restatement."""
import numpy as np
import pytest

import oracle_lib
from common import assert_fragments_equal, small_workload
from common_build import assert_flat_equal, build_workload, rescue_requests
from isaac_aligner_b200 import synth
from isaac_aligner_b200.batch import Tls
from isaac_aligner_b200.types import BWA_SCORES, Config, ReadSet
from test_adapters import ADAPTER_SETS, check_golden, golden_cases, golden_inputs

REF = oracle_lib.reference()
PORT = oracle_lib.port()
needs_ref = pytest.mark.skipif(REF is None or not hasattr(REF.lib, "oracle_set_adapters"),
                               reason="oracle/_ref/libisaac_ref.so not built (needs /root/reference)")


def test_reference_adapter_goldens_through_the_restatement():
    gold = golden_cases()
    try:
        for case in gold["cases"]:
            genome, reads, cand = golden_inputs(case)
            cfg = Config.default(tuple(gold["scores"]), max_read_length=len(case["read"]))
            PORT.set_adapters([tuple(a) for a in case["adapters"]])
            frag, cigar, _ = PORT.ungapped(oracle_lib.GenomeHolder(genome), reads, cfg, cand)
            check_golden(case, frag, cigar)
    finally:
        PORT.set_adapters(())


def both(fn, adapters):
    try:
        for chk in (REF, PORT):
            chk.set_adapters(adapters)
        return fn(REF), fn(PORT)
    finally:
        for chk in (REF, PORT):
            chk.set_adapters(())


@needs_ref
@pytest.mark.parametrize("kind,avoid", [("standard", False), ("nextera", False), ("matepair", False), ("matepair", True), (None, True)])
def test_micro_port_matches_reference(kind, avoid):
    adapters, inserted, read_through = ADAPTER_SETS[kind] if kind else ((), None, True)
    genome, sim, reads, cand = small_workload(n_pairs=2500, L=100, seed=701, indel_rate=8e-3)
    if kind:
        synth.insert_adapters(sim, inserted, fraction=0.5, seed=702, read_through=read_through)
        reads = ReadSet(sim.bcl, (100, 100), end_cycles_masked=reads.end_cycles_masked)
    lens = np.array([g.size for g in genome])
    cand = cand[(cand["position"] >= 0) & (cand["position"] + 100 <= lens[cand["contigStrand"] >> 1])]
    cfg = Config.default(BWA_SCORES, max_read_length=200, avoid_smith_waterman=avoid)
    g = oracle_lib.GenomeHolder(genome)
    r, p = both(lambda chk: chk.ungapped(g, reads, cfg, cand), adapters)
    assert_fragments_equal(r[0], p[0], r[1], p[1], r[2], p[2], "ungapped port vs reference")
    if kind:
        assert ((r[0]["lowClipped"] + r[0]["highClipped"] > 0) & (r[0]["cigarLength"] > 0)).sum() > 500
    gc = cand[r[0]["cigarLength"] > 0]
    r, p = both(lambda chk: chk.gapped(g, reads, cfg, gc, threads=1), adapters)      # one GappedAligner in call order (its table cache)
    assert_fragments_equal(r[0], p[0], r[1], p[1], r[2], p[2], "gapped port vs reference")
    if avoid:
        plain = Config.default(BWA_SCORES, max_read_length=200)
        every = both(lambda chk: chk.gapped(g, reads, plain, gc, threads=1), adapters)[0][0]
        assert ((r[0]["cigarLength"] == 0) & (every["cigarLength"] > 0)).sum() > 500 and (r[0]["gapCount"] > 0).sum() > 50


@needs_ref
@pytest.mark.parametrize("kind,avoid,L", [("standard", False, 150), ("matepair", True, 100), (None, True, 100)])
def test_tile_calls_port_matches_reference(kind, avoid, L):
    adapters, inserted, read_through = ADAPTER_SETS[kind] if kind else ((), None, True)
    genome, sim, reads, mb = build_workload(n_pairs=1500, L=L, seed=710 + L, indel_rate=6e-3)
    if kind:
        synth.insert_adapters(sim, inserted, fraction=0.4, seed=711, read_through=read_through, min_keep=40)
        reads = ReadSet(sim.bcl, (L, L), end_cycles_masked=reads.end_cycles_masked)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L, avoid_smith_waterman=avoid)
    g = oracle_lib.GenomeHolder(genome)
    r, p = both(lambda chk: oracle_lib.build_fragments(chk, g, reads, cfg, mb), adapters)
    assert_flat_equal(r, p, "build port vs reference (adapters %s, avoid %s)" % (kind, avoid))
    assert (r.fragments["gapCount"] > 0).sum() > 20
    tls = Tls.make()
    req = rescue_requests(sim, seed=712)
    r, p = both(lambda chk: oracle_lib.rescue_shadows(chk, g, reads, cfg, tls, req), adapters)
    assert_flat_equal(r, p, "rescue port vs reference (adapters %s, avoid %s)" % (kind, avoid))
    assert r.flags.mean() > 0.5


@needs_ref
@pytest.mark.parametrize("n_pairs,drift,with_pf,fixed_insert", [(3000, -1, True, False), (26000, 40, False, False), (30000, -1, False, True)])
def test_template_length_port_matches_reference(n_pairs, drift, with_pf, fixed_insert):
    """MatchSelector::determineTemplateLength + TemplateLengthDistribution: tiles that end before stability (finalize), a mate
    drift range, filtered clusters, and a single-insert-size library that turns stable at the second update"""
    import ctypes
    from test_gpu_tls import swap_reads, words
    extra = dict(insert=(350, 0, 350, 350), indel_rate=0.0) if fixed_insert else dict(indel_rate=2e-3)
    genome, sim, reads, mb = build_workload(n_pairs=n_pairs, L=100, seed=800 + n_pairs % 89, genome_bases=2_000_000, **extra)
    reads, mb = swap_reads(sim, reads, mb)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    pf = (np.random.default_rng(6).random(n_pairs) < 0.9).astype(np.uint8) if with_pf else None
    g = oracle_lib.GenomeHolder(genome)
    r, r_stable = oracle_lib.determine_template_length(REF, g, reads, cfg, mb, pf, drift)
    p, p_stable = oracle_lib.determine_template_length(PORT, g, reads, cfg, mb, pf, drift)
    assert words(r) == words(p) and r_stable == p_stable, (words(r), words(p), r_stable, p_stable)
    assert r_stable or not fixed_insert
    assert sorted([r.bestModel[0], r.bestModel[1]]) == [1, 6]
