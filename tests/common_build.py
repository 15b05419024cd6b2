"""Seeded workloads and comparison helpers for the FragmentBuilder::build / ShadowAligner::rescueShadow parity tests."""
import numpy as np

from isaac_aligner_b200 import synth
from isaac_aligner_b200.batch import RESCUE_REQUEST_DTYPE, MatchBatch
from isaac_aligner_b200.types import FRAGMENT_DTYPE, ReadSet, cigar_to_string


def build_workload(n_pairs=1500, L=150, seed=5, genome_bases=300_000, n_contigs=2, indel_rate=4e-3, masked=True,
                   with_gaps=True, neighbor_rate=0.15, repeat_rate=0.01, too_many_rate=0.005, repeat_threshold=10, **simulate):
    """genome + simulated pairs + seed matches with decoys, neighbours, over-represented seeds and TooManyMatch records"""
    genome = synth.make_genome(genome_bases, n_contigs=n_contigs, seed=seed, n_fraction=0.002, n_run=(20, 200))
    sim = synth.simulate_pairs(genome, n_pairs, L=L, seed=seed + 1, indel_rate=indel_rate,
                               seed_offsets=synth.auto_seed_offsets(L), **simulate)
    rng = np.random.default_rng(seed + 2)
    ecm = None
    if masked:
        ecm = np.where(rng.random((n_pairs, 2)) < 0.15, rng.integers(1, 25, size=(n_pairs, 2)), 0).astype(np.uint16)
    reads = ReadSet(sim.bcl, (L, L), end_cycles_masked=ecm)
    matches, begin = synth.make_matches(sim, genome, seed=seed + 3, decoy_rate=0.3, neighbor_rate=neighbor_rate,
                                        repeat_rate=repeat_rate, too_many_rate=too_many_rate,
                                        repeat_matches=repeat_threshold + 2)
    # a few clusters without any match (build() returns false) and a few where only one read has matches
    counts = np.diff(begin.astype(np.int64))
    cluster_of = np.repeat(np.arange(n_pairs), counts)
    drop_cluster = rng.random(n_pairs) < 0.03
    read_of = ((matches["seedId"] >> np.uint64(1)) & np.uint64(0xFF)) // np.uint64(len(sim.seed_offsets))
    drop_read1 = rng.random(n_pairs) < 0.05
    keep = ~(drop_cluster[cluster_of] | (drop_read1[cluster_of] & (read_of == 1)))
    matches = matches[keep]
    begin = np.zeros(n_pairs + 1, dtype=np.uint64)
    np.cumsum(np.bincount(cluster_of[keep], minlength=n_pairs), out=begin[1:])
    return genome, sim, reads, MatchBatch(matches, begin, synth.seed_table(sim), with_gaps=with_gaps)


def rescue_requests(sim, seed=9, fraction=1.0, decoy_fraction=0.05, best_template_fraction=0.3):
    """one rescueShadow request per read taken as the orphan at its true locus (the mate is to be rescued), plus a few
    orphans at wrong loci / wrong strands and requests carrying a bestTemplateLength"""
    rng = np.random.default_rng(seed)
    n = sim.contig.shape[0]
    pick = np.nonzero(rng.random(2 * n) < fraction)[0]
    req = np.zeros(pick.size, dtype=RESCUE_REQUEST_DTYPE)
    cluster, r = pick // 2, pick % 2
    req["orphanReadId"] = pick
    reverse = sim.reverse[cluster, r].astype(np.uint32)
    flip = rng.random(pick.size) < decoy_fraction
    reverse = np.where(flip, 1 - reverse, reverse)
    req["orphanContigStrand"] = (sim.contig[cluster, r] << 1) | reverse
    shift = np.where(rng.random(pick.size) < decoy_fraction, rng.integers(-3000, 3000, size=pick.size), 0)
    req["orphanPosition"] = np.maximum(sim.position[cluster, r] + shift, 0)
    req["orphanObservedLength"] = sim.L + rng.integers(-3, 4, size=pick.size)
    req["bestTemplateLength"] = np.where(rng.random(pick.size) < best_template_fraction, rng.integers(200, 1200, size=pick.size), 0)
    return req


def assert_flat_equal(a, b, what=""):
    """bit-exact comparison of two FlatFragments (isaac_aligner_b200.batch)"""
    assert np.array_equal(a.flags, b.flags), "%s: flags differ at %s" % (what, np.nonzero(a.flags != b.flags)[0][:10])
    if not np.array_equal(a.begin, b.begin):
        i = int(np.nonzero(a.begin != b.begin)[0][0]) - 1
        raise AssertionError("%s: fragment counts differ first at group %d: %d vs %d\n%s\n%s" % (
            what, i, int(a.begin[i + 1] - a.begin[i]), int(b.begin[i + 1] - b.begin[i]),
            [(int(f["position"]), int(f["reverse"]), cigar_to_string(a.cigar(k))) for k, f in
             zip(range(int(a.begin[i]), int(a.begin[i + 1])), a.fragments[int(a.begin[i]):int(a.begin[i + 1])])],
            [(int(f["position"]), int(f["reverse"]), cigar_to_string(b.cigar(k))) for k, f in
             zip(range(int(b.begin[i]), int(b.begin[i + 1])), b.fragments[int(b.begin[i]):int(b.begin[i + 1])])]))
    for name in FRAGMENT_DTYPE.names:
        if name in ("cigarOffset", "matchCount"):
            continue
        x, y = a.fragments[name], b.fragments[name]
        if name == "logProbability":
            x, y = x.view(np.uint64), y.view(np.uint64)
        bad = np.nonzero(x != y)[0]
        if bad.size:
            i = int(bad[0])
            raise AssertionError("%s: field %s differs for %d of %d fragments, first %d: %r vs %r\n%r\n%r\n%s | %s" % (
                what, name, bad.size, x.size, i, a.fragments[name][i], b.fragments[name][i], a.fragments[i], b.fragments[i],
                cigar_to_string(a.cigar(i)), cigar_to_string(b.cigar(i))))
    for i in range(a.fragments.size):
        if not np.array_equal(a.cigar(i), b.cigar(i)):
            raise AssertionError("%s: cigar of fragment %d differs: %s vs %s" % (what, i, cigar_to_string(a.cigar(i)), cigar_to_string(b.cigar(i))))
