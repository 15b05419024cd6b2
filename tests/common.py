"""Shared helpers for the parity tests: seeded workloads and field-by-field comparison."""
import numpy as np

from isaac_aligner_b200 import synth
from isaac_aligner_b200.types import CANDIDATE_DTYPE, FRAGMENT_DTYPE, ReadSet, cigar_to_string

FRAGMENT_FIELDS = [n for n in FRAGMENT_DTYPE.names]


def random_sw_cases(n, seed, lmin=30, lmax=250, with_n=True, band=16, max_indel=8):
    """Random (query, database) pairs like SURVEY Appendix A's validation set: the query is a copy of the database
    window at a random band offset with substitutions, 0-2 indels of 1-8 bases, optional 'n' in the query and 'N' in
    the database."""
    rng = np.random.default_rng(seed)
    queries, dbs = [], []
    for _ in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        db = rng.integers(0, 4, size=L + band - 1 + 2 * band)
        start = int(rng.integers(0, band))
        src = list(db[start:start + L + band])
        for _e in range(int(rng.integers(0, 3))):
            p = int(rng.integers(5, max(6, len(src) - 20)))
            ln = int(rng.integers(1, max_indel + 1))
            if rng.random() < 0.5:
                del src[p:p + ln]
            else:
                src[p:p] = list(rng.integers(0, 4, size=ln))
        q = np.array(src[:L])
        if q.size < L:
            q = np.concatenate([q, rng.integers(0, 4, size=L - q.size)])
        sub = rng.random(L) < rng.choice([0.0, 0.02, 0.1, 0.5])
        q = np.where(sub, (q + rng.integers(1, 4, size=L)) & 3, q)
        qs = bytearray(b"ACGT"[int(c)] for c in q)
        ds = bytearray(b"ACGT"[int(c)] for c in db[:L + band - 1])
        if with_n and rng.random() < 0.2:
            for p in rng.integers(0, L, size=int(rng.integers(1, 4))):
                qs[int(p)] = ord("n")
        if with_n and rng.random() < 0.1:
            p = int(rng.integers(0, L))
            run = min(int(rng.integers(1, 6)), L + band - 1 - p)
            ds[p:p + run] = b"N" * run
        queries.append(bytes(qs))
        dbs.append(bytes(ds))
    return queries, dbs


def small_workload(n_pairs=2000, L=100, seed=7, genome_bases=200_000, n_contigs=2, n_fraction=0.002, indel_rate=3e-3,
                   masked=True):
    """A small genome + simulated pairs + a candidate list that covers true loci, shifted loci, random loci, loci
    hanging over both contig ends and quality-trimmed reads."""
    genome = synth.make_genome(genome_bases, n_contigs=n_contigs, seed=seed, n_fraction=n_fraction, n_run=(20, 200))
    sim = synth.simulate_pairs(genome, n_pairs, L=L, seed=seed + 1, indel_rate=indel_rate)
    rng = np.random.default_rng(seed + 2)
    ecm = None
    if masked:
        ecm = np.where(rng.random((n_pairs, 2)) < 0.2, rng.integers(1, 30, size=(n_pairs, 2)), 0).astype(np.uint16)
    reads = ReadSet(sim.bcl, (L, L), end_cycles_masked=ecm)
    cand = synth.microbench_candidates(sim, genome, per_read=3, seed=seed + 3)
    # edge cases: candidates overhanging the contig start / end, and exactly at the ends
    n_edge = 200
    edge = np.empty(n_edge, dtype=CANDIDATE_DTYPE)
    edge["readId"] = rng.integers(0, 2 * n_pairs, size=n_edge)
    contig = rng.integers(0, n_contigs, size=n_edge)
    edge["contigStrand"] = (contig << 1) | rng.integers(0, 2, size=n_edge)
    lens = np.array([g.size for g in genome])
    at_end = rng.random(n_edge) < 0.5
    edge["position"] = np.where(at_end, lens[contig] - rng.integers(2, L + 5, size=n_edge), rng.integers(-L + 1, 20, size=n_edge))
    return genome, sim, reads, np.concatenate([cand, edge])


def assert_fragments_equal(a, b, cig_a, cig_b, mask_a=None, mask_b=None, what=""):
    """bit-exact comparison of two result sets (isaac_ext_fragment_t arrays + cigar pools [n, stride])"""
    assert a.shape == b.shape
    for name in FRAGMENT_FIELDS:
        x, y = a[name], b[name]
        if name == "logProbability":
            x, y = x.view(np.uint64), y.view(np.uint64)
        if name == "cigarOffset":
            continue
        bad = np.nonzero(x != y)[0]
        if bad.size:
            i = int(bad[0])
            raise AssertionError("%s: field %s differs at %d of %d candidates, first %d: %r vs %r\n%r\n%r\n%s | %s" % (
                what, name, bad.size, a.size, i, a[name][i], b[name][i], a[i], b[i],
                cigar_to_string(cig_a[i][:a["cigarLength"][i]]), cigar_to_string(cig_b[i][:b["cigarLength"][i]])))
    for i in range(a.size):
        n = int(a["cigarLength"][i])
        if n and not np.array_equal(cig_a[i][:n], cig_b[i][:n]):
            raise AssertionError("%s: cigar differs at %d: %s vs %s" % (what, i, cigar_to_string(cig_a[i][:n]), cigar_to_string(cig_b[i][:n])))
    if mask_a is not None and mask_b is not None:
        aligned = a["cigarLength"] > 0
        assert np.array_equal(mask_a[aligned], mask_b[aligned]), what + ": mismatch masks differ"
