"""Parity on the genome BASELINE configs[2] is quoted on (3.1 Gbp in 24 contigs, 0.1 % N runs of 100-10 000
bases): the packed reference spans more than 2^31 bases, so every global base index of the kernels has to be 64-bit; reads fall
into N runs and next to contig ends.  Sampled (the CPU checkers run 20 000 pairs in seconds); the full-size run of this
configuration is bench.py's pairs_pipeline line."""
import os

import numpy as np
import pytest

import oracle_lib
from common import assert_fragments_equal
from isaac_aligner_b200 import synth
from isaac_aligner_b200.batch import MatchBatch, Tls, TemplateOptions
from isaac_aligner_b200.types import CANDIDATE_DTYPE, Config, ReadSet
from test_gpu_templates import assert_templates_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def human():
    workers = 1          # no fork workers inside a pytest process that already holds a CUDA context
    genome = synth.make_genome_parallel(3_100_000_000, n_contigs=24, seed=synth.SEED_G3100, n_fraction=0.001, workers=workers)
    from isaac_aligner_b200 import capi
    ctx = capi.Context(Config.default(max_read_length=300))
    ctx.set_reference(genome)
    yield genome, ctx, capi
    ctx.close()


def test_candidates_beyond_two_to_the_31st_base(human):
    """ungapped + gapped extension of candidates on the last contigs (global offsets 2.7 .. 3.1 G), inside and next to N runs, at
    both ends of a contig: bit-exact with both checkers"""
    genome, ctx, capi = human
    L = 150
    late = genome[20:]                                                   # the checkers get the last four contigs only (their ids shifted)
    sim = synth.simulate_pairs(late, 3000, L=L, seed=31, indel_rate=5e-3)
    reads = ReadSet(sim.bcl, (L, L))
    cand = synth.microbench_candidates(sim, late, per_read=4, seed=32)
    rng = np.random.default_rng(33)
    extra = np.zeros(2000, dtype=CANDIDATE_DTYPE)                         # around N runs and contig ends
    extra["readId"] = rng.integers(0, 6000, size=extra.size)
    c = rng.integers(0, len(late), size=extra.size)
    n_at = [np.nonzero(late[k][::64] == ord("N"))[0] * 64 for k in range(len(late))]
    pos = np.array([int(rng.choice(n_at[k])) + int(rng.integers(-L, 20)) if (i % 2 and n_at[k].size) else
                    int(rng.choice([rng.integers(-L + 10, 5), late[k].size - rng.integers(1, L)])) for i, k in enumerate(c)])
    extra["position"] = pos
    extra["contigStrand"] = (c.astype(np.uint32) << 1) | rng.integers(0, 2, size=extra.size).astype(np.uint32)
    cand = np.concatenate([cand, extra])
    shifted = cand.copy()
    shifted["contigStrand"] += np.uint32(20 << 1)                         # the same contigs under their ids in the whole genome
    ctx.set_reads(reads)
    fu, cu, mu = ctx.ungapped(shifted)
    keep = fu["cigarLength"] > 0
    fg, cg, mg = ctx.gapped(shifted[keep])
    g = oracle_lib.GenomeHolder(late)
    for chk in oracle_lib.gpu_checkers():
        ru = chk.ungapped(g, reads, ctx.config, cand)
        rg = chk.gapped(g, reads, ctx.config, cand[keep])
        for r in (ru[0], rg[0]):
            r["contigId"] += 20
        assert_fragments_equal(fu, ru[0], cu, ru[1], mu, ru[2], "human-scale ungapped " + chk.kind)
        assert_fragments_equal(fg, rg[0], cg, rg[1], mg, rg[2], "human-scale gapped " + chk.kind)
    assert (fu["mismatchCount"][len(cand) - 2000:] > 0).sum() > 500


def test_templates_on_the_whole_genome(human):
    """the whole TemplateBuilder on 20 000 pairs drawn from all 24 contigs against the reference's own TemplateBuilder holding
    the same 3.1 Gbp"""
    genome, ctx, capi = human
    L = 150
    sim = synth.simulate_pairs(genome, 20000, L=L, seed=41, indel_rate=2e-3, seed_offsets=synth.auto_seed_offsets(L))
    matches, begin = synth.make_matches(sim, genome, seed=42, decoy_rate=0.3, neighbor_rate=0.1)
    reads = ReadSet(sim.bcl, (L, L))
    mb = MatchBatch(matches, begin, synth.seed_table(sim), with_gaps=True)
    ctx.set_reads(reads)
    options = TemplateOptions.make(clip_semialigned=True)
    got = ctx.build_templates(mb, Tls.make(), options)
    ref = oracle_lib.Oracle(oracle_lib.REF_SO)
    want = oracle_lib.build_templates(ref, oracle_lib.GenomeHolder(genome), reads, ctx.config, mb, Tls.make(), options, threads=8)
    assert_templates_equal(got, want, "templates on G3100")
    assert got.templates["built"].mean() > 0.9 and got.rescue_requests > 5000
    assert (got.fragments["contigId"][got.fragments["cigarLength"] > 0] >= 18).sum() > 2000       # beyond 2^31 bases
