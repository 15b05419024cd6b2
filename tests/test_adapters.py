"""Sequencing-adapter clipping (FragmentSequencingAdapterClipper + SequencingAdapter).

* the literal vectors of the reference's own unit test (tests/golden/adapters.json, transcribed from testSequencingAdapter.cpp
  by tests/golden/make_adapter_goldens.py) replayed through the reference build of the checker (CPU) and through the CUDA
  path (GPU);
* seeded workloads whose reads run into adapters, CUDA path vs the reference's own code, bit-exact: the micro entry points
  (one clipper per candidate), FragmentBuilder::build, rescueShadow and the whole TemplateBuilder (one clipper per read list /
  rescue, the first candidate of a strand locates the adapter)."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from common import assert_fragments_equal, small_workload
from common_build import assert_flat_equal, build_workload, rescue_requests
from isaac_aligner_b200 import synth
from isaac_aligner_b200.types import (BWA_SCORES, CANDIDATE_DTYPE, NEXTERA_MATEPAIR_ADAPTERS, NEXTERA_STANDARD_ADAPTERS,
                                      STANDARD_ADAPTERS, Config, ReadSet, cigar_to_string)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adapters.json")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def reference_checker():
    if not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"):
        pytest.skip("the reference build of the checker did not travel to this box")
    chk = oracle_lib.reference()
    if chk is None or not hasattr(chk.lib, "oracle_set_adapters"):
        pytest.skip("no reference checker with adapter support")
    return chk


def golden_inputs(case):
    """one single-read cluster whose strand sequence is the case's read string, Q30 everywhere"""
    read = case["read"]
    fwd = "".join(COMP[b] for b in reversed(read)) if case["reverse"] else read
    bcl = np.array([(30 << 2) | "ACGT".index(b) for b in fwd], dtype=np.uint8)[None, :]
    reads = ReadSet(bcl, (len(read),))
    genome = [np.frombuffer(case["reference"].encode(), dtype=np.uint8)]
    cand = np.zeros(1, dtype=CANDIDATE_DTYPE)
    cand["contigStrand"] = 1 if case["reverse"] else 0
    return genome, reads, cand


def check_golden(case, frag, cigar):
    assert cigar_to_string(cigar[0][:frag["cigarLength"][0]]) == case["cigar"], case["name"]
    for key in ("mismatchCount", "editDistance", "observedLength", "position"):
        if key in case:
            assert int(frag[key][0]) == case[key], (case["name"], key)


def golden_cases():
    return json.load(open(GOLDEN))


def test_reference_adapter_goldens_through_the_reference_checker():
    chk = reference_checker()
    gold = golden_cases()
    try:
        for case in gold["cases"]:
            genome, reads, cand = golden_inputs(case)
            cfg = Config.default(tuple(gold["scores"]), max_read_length=len(case["read"]))
            chk.set_adapters([tuple(a) for a in case["adapters"]])
            frag, cigar, _ = chk.ungapped(oracle_lib.GenomeHolder(genome), reads, cfg, cand)
            check_golden(case, frag, cigar)
    finally:
        chk.set_adapters(())


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


@pytest.mark.gpu
def test_reference_adapter_goldens_cuda(capi):
    gold = golden_cases()
    for case in gold["cases"]:
        genome, reads, cand = golden_inputs(case)
        ctx = capi.Context(Config.default(tuple(gold["scores"]), max_read_length=len(case["read"])))
        ctx.set_reference(genome)
        ctx.set_reads(reads)
        ctx.set_adapters([tuple(a) for a in case["adapters"]])
        frag, cigar, _ = ctx.ungapped(cand)
        check_golden(case, frag, cigar)
        ctx.close()


@pytest.mark.gpu
def test_set_adapters_rejects_what_the_reference_asserts(capi):
    ctx = capi.Context(Config.default())
    for bad in ([("ACGT", False, 0)], [("ACGTNACGT", False, 0)], [("acgtacgt", False, 0)], [("ACGTACGTAC", False, 5)], [("A" * 127, False, 0)]):
        with pytest.raises(capi.ExtError) as e:
            ctx.set_adapters(bad)
        assert e.value.code == 1
    ctx.set_adapters(STANDARD_ADAPTERS)
    ctx.set_adapters(())
    ctx.close()


ADAPTER_SETS = {
    "standard": (STANDARD_ADAPTERS, ["AGATCGGAAGAGC"], True),
    "nextera": (NEXTERA_STANDARD_ADAPTERS, ["CTGTCTCTTATACACATCT"], True),
    "matepair": (NEXTERA_MATEPAIR_ADAPTERS, ["CTGTCTCTTATACACATCT", "AGATGTGTATAAGAGACAG", "CTGTCTCTTATACACATCTAGATGTGTATAAGAGACAG"], False),
}


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["standard", "nextera", "matepair"])
def test_micro_entry_points_with_adapters(capi, kind):
    """alignUngapped / alignGapped behind one clipper per candidate, bit-exact against the reference's own classes"""
    chk = reference_checker()
    adapters, inserted, read_through = ADAPTER_SETS[kind]
    genome, sim, reads, cand = small_workload(n_pairs=3000, L=100, seed=301, indel_rate=4e-3)
    cuts = synth.insert_adapters(sim, inserted, fraction=0.5, seed=302, read_through=read_through)
    reads = ReadSet(sim.bcl, (100, 100), end_cycles_masked=reads.end_cycles_masked)
    # candidates that stay inside their contig: where the adapter reaches an end of the read the reference compares against
    # its contig without a bounds check (FragmentSequencingAdapterClipper.cpp:190-216)
    lens = np.array([g.size for g in genome])
    cand = cand[(cand["position"] >= 0) & (cand["position"] + 100 <= lens[cand["contigStrand"] >> 1])]
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    ctx.set_adapters(adapters)
    g = oracle_lib.GenomeHolder(genome)
    try:
        chk.set_adapters(adapters)
        fu, cu, mu = ctx.ungapped(cand)
        ru = chk.ungapped(g, reads, cfg, cand, threads=8)
        assert_fragments_equal(fu, ru[0], cu, ru[1], mu, ru[2], "ungapped with %s adapters" % kind)
        clipped = (fu["lowClipped"] + fu["highClipped"] > 0) & (fu["cigarLength"] > 0)
        assert clipped.sum() > 500, clipped.sum()
        gc = cand[fu["cigarLength"] > 0]
        fg, cg, mg = ctx.gapped(gc)
        rg = chk.gapped(g, reads, cfg, gc, threads=8)
        assert_fragments_equal(fg, rg[0], cg, rg[1], mg, rg[2], "gapped with %s adapters" % kind)
        # the chunked end-to-end entry point takes the same path
        f1 = np.zeros(len(cand), dtype=capi.FRAGMENT_DTYPE)
        pool = np.zeros(len(cand) * 3, dtype=np.uint32)
        ctx.extend_compact(cand, False, f1, pool)
        assert np.array_equal(f1["lowClipped"], fu["lowClipped"]) and np.array_equal(f1["mismatchCount"], fu["mismatchCount"])
    finally:
        chk.set_adapters(())
    # without adapters the same reads align unclipped
    ctx.set_adapters(())
    f0 = ctx.ungapped(cand)[0]
    assert (f0["mismatchCount"] >= fu["mismatchCount"]).all() and (f0["mismatchCount"] > fu["mismatchCount"]).sum() > 500
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,L", [("standard", 150), ("matepair", 100)])
def test_tile_calls_with_adapters(capi, kind, L):
    """FragmentBuilder::build, rescueShadow and TemplateBuilder keep one clipper per read list / rescue call"""
    from isaac_aligner_b200.batch import Tls, TemplateOptions
    chk = reference_checker()
    adapters, inserted, read_through = ADAPTER_SETS[kind]
    genome, sim, reads, mb = build_workload(n_pairs=3000, L=L, seed=310 + L, indel_rate=4e-3)
    synth.insert_adapters(sim, inserted, fraction=0.4, seed=311, read_through=read_through, min_keep=40)
    reads = ReadSet(sim.bcl, (L, L), end_cycles_masked=reads.end_cycles_masked)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    ctx.set_adapters(adapters)
    g = oracle_lib.GenomeHolder(genome)
    try:
        chk.set_adapters(adapters)
        got = ctx.build_fragments(mb)
        assert_flat_equal(got, oracle_lib.build_fragments(chk, g, reads, cfg, mb, threads=8), "build_fragments with %s adapters" % kind)
        f = got.fragments
        assert ((f["lowClipped"] + f["highClipped"] > 0) & (f["cigarLength"] > 0)).sum() > 300
        tls = Tls.make()
        req = rescue_requests(sim, seed=312)
        gotr = ctx.rescue_shadows(tls, req)
        assert_flat_equal(gotr, oracle_lib.rescue_shadows(chk, g, reads, cfg, tls, req, threads=8), "rescue_shadows with %s adapters" % kind)
        options = TemplateOptions.make()
        t = ctx.build_templates(mb, tls, options)
        want = oracle_lib.build_templates(chk, g, reads, cfg, mb, tls, options, threads=8)
        for name in ("built", "properPair", "alignmentScore", "fragmentAlignmentScore"):
            assert np.array_equal(t.templates[name], want.templates[name]), name
        for name in ("position", "contigId", "observedLength", "editDistance", "cigarLength", "lowClipped", "highClipped", "mismatchCount"):
            assert np.array_equal(t.fragments[name], want.fragments[name]), name
    finally:
        chk.set_adapters(())
    ctx.close()


# ---- --avoid-smith-waterman: GappedAligner::makesSenseToGapAlign -------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("kind", [None, "matepair"])
def test_avoid_smith_waterman_micro(capi, kind):
    """alignGapped with the 7-mer heuristic in front, candidates in call order through one GappedAligner (its table is cached
    per read and strand); with adapters the clipped range differs between candidates of a read, which is what the cache
    does not notice"""
    chk = reference_checker()
    genome, sim, reads, cand = small_workload(n_pairs=3000, L=100, seed=321, indel_rate=8e-3)
    adapters = ()
    if kind:
        adapters, inserted, read_through = ADAPTER_SETS[kind]
        synth.insert_adapters(sim, inserted, fraction=0.4, seed=322, read_through=read_through)
        reads = ReadSet(sim.bcl, (100, 100), end_cycles_masked=reads.end_cycles_masked)
    lens = np.array([g.size for g in genome])
    cand = cand[(cand["position"] >= 0) & (cand["position"] + 100 <= lens[cand["contigStrand"] >> 1])]
    cfg = Config.default(BWA_SCORES, max_read_length=200, avoid_smith_waterman=True)
    plain = Config.default(BWA_SCORES, max_read_length=200)
    g = oracle_lib.GenomeHolder(genome)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    ctx.set_adapters(adapters)
    try:
        chk.set_adapters(adapters)
        gc = cand[ctx.ungapped(cand)[0]["cigarLength"] > 0]
        fg, cg, mg = ctx.gapped(gc)
        rg = chk.gapped(g, reads, cfg, gc, threads=1)
        assert_fragments_equal(fg, rg[0], cg, rg[1], mg, rg[2], "gapped with --avoid-smith-waterman")
        every = chk.gapped(g, reads, plain, gc, threads=8)[0]
        skipped = (fg["cigarLength"] == 0) & (every["cigarLength"] > 0)
        kept = (fg["cigarLength"] > 0) & (fg["gapCount"] > 0)
        assert skipped.sum() > 1000 and kept.sum() > 100, (skipped.sum(), kept.sum())
    finally:
        chk.set_adapters(())
    ctx.close()


@pytest.mark.gpu
def test_avoid_smith_waterman_tile_calls(capi):
    from isaac_aligner_b200.batch import Tls, TemplateOptions
    chk = reference_checker()
    L = 150
    genome, sim, reads, mb = build_workload(n_pairs=4000, L=L, seed=330, indel_rate=8e-3)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L, avoid_smith_waterman=True)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    g = oracle_lib.GenomeHolder(genome)
    got = ctx.build_fragments(mb)
    assert_flat_equal(got, oracle_lib.build_fragments(chk, g, reads, cfg, mb, threads=8), "build_fragments, --avoid-smith-waterman")
    assert (got.fragments["gapCount"] > 0).sum() > 50
    tls = Tls.make()
    req = rescue_requests(sim, seed=331)
    gotr = ctx.rescue_shadows(tls, req)
    assert_flat_equal(gotr, oracle_lib.rescue_shadows(chk, g, reads, cfg, tls, req, threads=8), "rescue_shadows, --avoid-smith-waterman")
    options = TemplateOptions.make()
    t = ctx.build_templates(mb, tls, options)
    want = oracle_lib.build_templates(chk, g, reads, cfg, mb, tls, options, threads=8)
    for name in ("built", "properPair", "alignmentScore", "fragmentAlignmentScore"):
        assert np.array_equal(t.templates[name], want.templates[name]), name
    for name in ("position", "contigId", "observedLength", "editDistance", "cigarLength", "mismatchCount", "gapCount"):
        assert np.array_equal(t.fragments[name], want.fragments[name]), name
    ctx.close()
