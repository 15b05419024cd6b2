"""Parity of isaac_ext_build_templates (TemplateBuilder::buildFragments + buildTemplate per cluster) with the
reference's own TemplateBuilder compiled unmodified (oracle/_ref), bit-exact: template and fragment mapping scores, proper-pair
flags, the chosen fragment records and their CIGARs."""
import os

import numpy as np
import pytest

import oracle_lib
from common import FRAGMENT_FIELDS
from common_build import build_workload
from isaac_aligner_b200.batch import DODGY_ALIGNMENT_SCORE_UNALIGNED, DODGY_ALIGNMENT_SCORE_UNKNOWN, Tls, TemplateOptions
from isaac_aligner_b200.types import BWA_SCORES, ELAND_SCORES, Config, cigar_to_string

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(oracle_lib.REF_SO), reason="the TemplateBuilder checker is the reference build only")]


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


def assert_templates_equal(got, want, what):
    for name in ("hadFragments", "built", "properPair", "alignmentScore", "fragmentAlignmentScore"):
        bad = np.nonzero((got.templates[name] != want.templates[name]).reshape(len(got.templates), -1).any(axis=1))[0]
        assert not bad.size, "%s: template field %s differs at %d clusters, first %d: %r vs %r\n%r\n%r" % (
            what, name, bad.size, bad[0], got.templates[name][bad[0]], want.templates[name][bad[0]],
            got.fragments[2 * bad[0]:2 * bad[0] + 2], want.fragments[2 * bad[0]:2 * bad[0] + 2])
    for name in FRAGMENT_FIELDS:
        if name in ("cigarOffset", "matchCount"):
            continue
        x, y = got.fragments[name], want.fragments[name]
        if name == "logProbability":
            x, y = x.view(np.uint64), y.view(np.uint64)
        bad = np.nonzero(x != y)[0]
        assert not bad.size, "%s: fragment field %s differs at %d records, first %d: %r vs %r\n%r\n%r" % (
            what, name, bad.size, bad[0], got.fragments[name][bad[0]], want.fragments[name][bad[0]],
            got.fragments[bad[0]], want.fragments[bad[0]])
    for i in np.nonzero(got.fragments["cigarLength"])[0]:
        assert np.array_equal(got.cigar(i), want.cigar(i)), "%s: cigar of record %d: %s vs %s" % (
            what, i, cigar_to_string(got.cigar(i)), cigar_to_string(want.cigar(i)))


def run_both(capi, genome, reads, mb, cfg, tls, options):
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    got = ctx.build_templates(mb, tls, options)
    ctx.close()
    ref = oracle_lib.Oracle(oracle_lib.REF_SO)
    want = oracle_lib.build_templates(ref, oracle_lib.GenomeHolder(genome), reads, cfg, mb, tls, options, threads=8)
    return got, want


@pytest.mark.parametrize("scores,L,options", [
    (BWA_SCORES, 150, TemplateOptions.make()),
    (BWA_SCORES, 100, TemplateOptions.make(scatter_repeats=True, dodgy=DODGY_ALIGNMENT_SCORE_UNKNOWN)),
    (ELAND_SCORES, 100, TemplateOptions.make(dodgy=DODGY_ALIGNMENT_SCORE_UNALIGNED, mapq_threshold=20)),
    (BWA_SCORES, 250, TemplateOptions.make(mapq_threshold=3)),
])
def test_build_templates_bit_exact(capi, scores, L, options):
    genome, sim, reads, mb = build_workload(n_pairs=6000, L=L, seed=300 + L, indel_rate=4e-3)
    cfg = Config.default(scores, max_read_length=2 * L)
    got, want = run_both(capi, genome, reads, mb, cfg, Tls.make(), options)
    assert_templates_equal(got, want, "build_templates L=%d" % L)
    t = got.templates
    assert t["built"].mean() > 0.9 and t["properPair"].mean() > 0.8 and got.rescue_requests > 100


def test_build_templates_baseline_config0(capi):
    """BASELINE configs[0]: 100 000 simulated 2x150 pairs on the 5 Mbp genome, the generator and seeds of bench.py"""
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch
    from isaac_aligner_b200.types import ReadSet
    L, n_pairs = 150, 100_000
    genome = synth.make_genome(5_000_000, n_contigs=1, seed=synth.SEED_G5)
    sim = synth.simulate_pairs(genome, n_pairs, L=L, seed=synth.SEED_READS + 7, indel_rate=5e-4, seed_offsets=synth.auto_seed_offsets(L))
    matches, begin = synth.make_matches(sim, genome, seed=synth.SEED_READS + 8, decoy_rate=0.2)
    reads = ReadSet(sim.bcl, (L, L))
    mb = MatchBatch(matches, begin, synth.seed_table(sim), with_gaps=True)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L)
    got, want = run_both(capi, genome, reads, mb, cfg, Tls.make(), TemplateOptions.make())
    assert_templates_equal(got, want, "config0 templates")
    assert got.templates["built"].mean() > 0.99


@pytest.mark.parametrize("seed,kw,options", [
    (411, dict(neighbor_rate=0.7), TemplateOptions.make()),                                  # few well-anchored candidates
    (412, dict(neighbor_rate=0.7, repeat_rate=0.1), TemplateOptions.make(scatter_repeats=True, dodgy=3)),
    (413, dict(genome_bases=40_000, n_contigs=3, indel_rate=1e-2), TemplateOptions.make(mapq_threshold=10)),   # crowded genome
    (414, dict(too_many_rate=0.1, masked=False), TemplateOptions.make(dodgy=DODGY_ALIGNMENT_SCORE_UNALIGNED)),
])
def test_build_templates_hard_cases(capi, seed, kw, options):
    genome, sim, reads, mb = build_workload(n_pairs=5000, L=100, seed=seed, **kw)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    got, want = run_both(capi, genome, reads, mb, cfg, Tls.make(), options)
    assert_templates_equal(got, want, "hard case %d" % seed)


def test_build_templates_tandem_repeats(capi):
    """a genome made of a repeated unit: many equally good placements per read, so the repeat bookkeeping (equal template
    scores, ISAAC_LP_EQUALS ties, unique-probability sums, --scatter-repeats) decides the result"""
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch
    from isaac_aligner_b200.types import ReadSet
    rng = np.random.default_rng(77)
    unit = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=700)]
    contig = np.concatenate([np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=3000)], np.tile(unit, 12),
                             np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=3000)]]).copy()
    genome = [contig]
    L = 100
    sim = synth.simulate_pairs(genome, 3000, L=L, seed=78, indel_rate=1e-3, seed_offsets=synth.auto_seed_offsets(L))
    matches, begin = synth.make_matches(sim, genome, seed=79, decoy_rate=0.1)
    # every seed match also matches at the same offset in the other copies of the unit: add them as extra matches
    reads = ReadSet(sim.bcl, (L, L))
    extra = []
    loc = matches["location"]
    pos = (loc >> np.uint64(1)) & np.uint64((1 << 40) - 1)
    inside = (pos >= 3000) & (pos < 3000 + 700 * 12 - 64)
    for k in range(1, 4):
        m = matches[inside].copy()
        p = (pos[inside].astype(np.int64) - 3000 + 700 * k) % (700 * 12 - 64) + 3000
        m["location"] = (m["location"] & ~np.uint64(((1 << 40) - 1) << 1)) | (p.astype(np.uint64) << np.uint64(1))
        extra.append((m, np.repeat(np.arange(len(begin) - 1), np.diff(begin.astype(np.int64)))[inside]))
    cluster_of = np.concatenate([np.repeat(np.arange(len(begin) - 1), np.diff(begin.astype(np.int64)))] + [e[1] for e in extra])
    allm = np.concatenate([matches] + [e[0] for e in extra])
    # a cluster's matches arrive sorted by (location, seed) (SelectMatchesTransition.cpp:242-254)
    order = np.lexsort((allm["seedId"], allm["location"], cluster_of))
    allm, cluster_of = allm[order], cluster_of[order]
    nb = np.zeros(len(begin), dtype=np.uint64)
    np.cumsum(np.bincount(cluster_of, minlength=len(begin) - 1), out=nb[1:])
    mb = MatchBatch(allm, nb, synth.seed_table(sim), with_gaps=True)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L)
    for options in (TemplateOptions.make(), TemplateOptions.make(scatter_repeats=True)):
        got, want = run_both(capi, genome, reads, mb, cfg, Tls.make(), options)
        assert_templates_equal(got, want, "tandem repeats")


@pytest.mark.parametrize("L,cutoff", [(100, 25), (150, 30), (36, 25), (30, 25), (100, 0)])
def test_trim_low_quality_ends_bit_exact(capi, L, cutoff):
    """alignment::trimLowQualityEnds (Quality.cpp:71-120): reads with decaying, noisy and flat quality profiles"""
    from isaac_aligner_b200.types import ReadSet
    rng = np.random.default_rng(500 + L)
    n = 4000
    decay = np.clip(40 - (np.arange(L)[None, :] * rng.uniform(0, 0.6, size=(2 * n, 1))) + rng.normal(0, 6, size=(2 * n, L)), 2, 41)
    decay[rng.random(2 * n) < 0.1] = 40                         # flat high quality: nothing to trim
    decay[rng.random(2 * n) < 0.05] = 2                          # all bad: everything but the last 35 cycles goes
    q = decay.astype(np.uint8).reshape(n, 2 * L)
    bases = rng.integers(0, 4, size=(n, 2 * L)).astype(np.uint8)
    bcl = (q << 2) | bases
    bcl[rng.random(bcl.shape) < 0.01] = 0                       # N: quality 2 (Read.cpp:63-68)
    reads = ReadSet(bcl, (L, L))
    ctx = capi.Context(Config.default(max_read_length=2 * L))
    ctx.set_reads(reads)
    got = ctx.trim_low_quality_ends(cutoff)
    want = oracle_lib.trim_low_quality_ends(oracle_lib.Oracle(oracle_lib.REF_SO), reads, cutoff)
    assert np.array_equal(got, want)
    if cutoff and L >= 100:
        assert (got > 0).mean() > 0.3 and (got == 0).any() and got.max() == L - 35
    ctx.close()


@pytest.mark.parametrize("L,seed,kw,options", [
    (100, 601, dict(), TemplateOptions.make(clip_semialigned=True)),
    (150, 602, dict(), TemplateOptions.make(clip_overlapping=True)),
    (150, 603, dict(indel_rate=8e-3), TemplateOptions.make(clip_semialigned=True, clip_overlapping=True)),
    (100, 604, dict(genome_bases=40_000, n_contigs=3), TemplateOptions.make(clip_semialigned=True, clip_overlapping=True, mapq_threshold=5,
                                                                           dodgy=DODGY_ALIGNMENT_SCORE_UNALIGNED)),
])
def test_end_clippers_bit_exact(capi, L, seed, kw, options):
    """SemialignedEndsClipper + OverlappingEndsClipper after buildTemplate (MatchSelector.cpp:336-346).  Short inserts make the
    mates overlap; substitution-rich read ends trigger the semialigned clipper."""
    from isaac_aligner_b200 import synth
    genome, sim, reads, mb = build_workload(n_pairs=5000, L=L, seed=seed, **kw)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L)
    tls = Tls.make()
    got, want = run_both(capi, genome, reads, mb, cfg, tls, options)
    assert_templates_equal(got, want, "end clippers %d" % seed)
    plain, _ = run_both(capi, genome, reads, mb, cfg, tls, TemplateOptions.make(dodgy=options.dodgyAlignmentScore,
                                                                                  mapq_threshold=options.mapqThreshold))
    changed = (plain.fragments["observedLength"] != got.fragments["observedLength"]).sum()
    assert changed > 20, "the clippers did not fire (%d)" % changed


def single_ended(genome, sim, reads, mb):
    """read 1 of a paired workload as a single-ended tile: its bases, its seeds, its matches"""
    from isaac_aligner_b200.batch import MatchBatch
    from isaac_aligner_b200.types import ReadSet
    L = reads.read_lengths[0]
    S = len(sim.seed_offsets)
    seed_index = (mb.matches["seedId"] >> np.uint64(1)) & np.uint64(0xFF)
    keep = seed_index < S
    cluster_of = np.repeat(np.arange(len(mb.begin) - 1), np.diff(mb.begin.astype(np.int64)))
    begin = np.zeros(len(mb.begin), dtype=np.uint64)
    np.cumsum(np.bincount(cluster_of[keep], minlength=len(mb.begin) - 1), out=begin[1:])
    ecm = reads.end_cycles_masked[:, :1] if reads.end_cycles_masked is not None else None
    return ReadSet(np.ascontiguousarray(reads.bcl[:, :L]), (L,), end_cycles_masked=ecm), \
        MatchBatch(mb.matches[keep], begin, mb.seeds[:S], with_gaps=True)


@pytest.mark.parametrize("options", [TemplateOptions.make(), TemplateOptions.make(dodgy=DODGY_ALIGNMENT_SCORE_UNALIGNED, mapq_threshold=4,
                                                                               clip_semialigned=True, scatter_repeats=True)])
def test_single_ended_templates_bit_exact(capi, options):
    """single-ended data: pickBestFragment (TemplateBuilder.cpp:1035-1058), one fragment per BamTemplate"""
    genome, sim, reads, mb = build_workload(n_pairs=5000, L=100, seed=700, neighbor_rate=0.5, repeat_rate=0.05)
    reads1, mb1 = single_ended(genome, sim, reads, mb)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads1)
    got = ctx.build_templates(mb1, Tls.make(), options)
    built = ctx.build_fragments(mb1)
    ctx.close()
    ref = oracle_lib.Oracle(oracle_lib.REF_SO)
    g = oracle_lib.GenomeHolder(genome)
    want = oracle_lib.build_templates(ref, g, reads1, cfg, mb1, Tls.make(), options, threads=4)
    from common_build import assert_flat_equal
    assert_flat_equal(built, oracle_lib.build_fragments(ref, g, reads1, cfg, mb1, threads=4), "single-ended build_fragments")
    # one fragment per cluster
    for name in ("hadFragments", "built", "properPair", "alignmentScore"):
        assert np.array_equal(got.templates[name], want.templates[name]), name
    assert np.array_equal(got.templates["fragmentAlignmentScore"][:, 0], want.templates["fragmentAlignmentScore"][:, 0])
    for name in FRAGMENT_FIELDS:
        if name in ("cigarOffset", "matchCount"):
            continue
        x, y = got.fragments[name], want.fragments[name]
        if name == "logProbability":
            x, y = x.view(np.uint64), y.view(np.uint64)
        assert np.array_equal(x, y), name
    for i in np.nonzero(got.fragments["cigarLength"])[0]:
        assert np.array_equal(got.cigar(i), want.cigar(i)), i
    assert got.templates["built"].mean() > 0.5 and got.rescue_requests == 0


@pytest.mark.parametrize("n_pairs", [3001, 700, 5003])
def test_build_templates_tiles_of_different_sizes(capi, n_pairs):
    """the device-resident tile pipeline keeps its buffers between calls: tiles of different sizes, one after the other on
    fresh contexts and larger / smaller than the one before, all give the reference's templates"""
    genome, sim, reads, mb = build_workload(n_pairs=n_pairs, L=100, seed=477, indel_rate=5e-3)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    got, want = run_both(capi, genome, reads, mb, cfg, Tls.make(), TemplateOptions.make(clip_semialigned=True))
    assert_templates_equal(got, want, "build_templates, %d pairs" % n_pairs)
    assert got.rescue_requests > n_pairs // 3
