"""Synthetic genomes, simulated read pairs and candidate lists.

Everything is generated from numpy's counter-based Philox generator with fixed seeds, so the CPU arm and
the GPU arm of a benchmark regenerate identical inputs.  This is workload generation, not alignment: it is
shared by bench.py and the tests.
"""
import numpy as np

SEED_G5 = 0x15AAC0001
SEED_G3100 = 0x15AAC0002
SEED_READS = 0x15AAC0100

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def make_genome(total_bases, n_contigs=1, seed=SEED_G5, n_fraction=0.0, n_run=(100, 10000)):
    """i.i.d. uniform ACGT contigs of equal length (last one shorter); optionally n_fraction of the bases
    replaced by runs of 'N' so that the N-mask path is exercised.  Returns a list of uint8 ASCII arrays."""
    rng = _rng(seed)
    per = -(-total_bases // n_contigs)
    contigs = []
    for c in range(n_contigs):
        n = min(per, total_bases - c * per)
        a = _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
        if n_fraction > 0:
            target = int(n * n_fraction)
            done = 0
            while done < target:
                run = int(rng.integers(n_run[0], n_run[1] + 1))
                start = int(rng.integers(0, max(1, n - run)))
                a[start:start + run] = ord("N")
                done += run
        contigs.append(a)
    return contigs


class Simulation:
    """Simulated FR read pairs.  bcl: [n_pairs, 2L] bytes (quality << 2 | base, 0 = N) in sequencing order;
    contig/position/reverse: [n_pairs, 2] truth of each read (position = leftmost forward-strand base of the
    strand-order read); shift1: [n_pairs, 2] net indel shift of the read's last base relative to its first."""

    def __init__(self, L):
        self.L = L
        self.bcl = None
        self.contig = None
        self.position = None
        self.reverse = None
        self.events = None  # [n_pairs, 2, K, 3] (strand-order read offset, +del/-ins length, unused)


def simulate_pairs(genome, n_pairs, L=150, seed=SEED_READS, snp_rate=1e-3, indel_rate=5e-4, max_indel=16,
                   insert=(350.0, 35.0, 200, 500), quality_probs=((40, 0.80), (30, 0.15), (20, 0.04), (2, 0.01)),
                   chunk=200_000, seed_offsets=None, seed_length=32):
    rng = _rng(seed)
    lens = np.array([c.size for c in genome], dtype=np.int64)
    margin = insert[3] + 2 * max_indel + 64
    weights = np.maximum(lens - margin, 0).astype(np.float64)
    weights /= weights.sum()
    sim = Simulation(L)
    sim.bcl = np.empty((n_pairs, 2 * L), dtype=np.uint8)
    sim.contig = np.empty((n_pairs, 2), dtype=np.uint32)
    sim.position = np.empty((n_pairs, 2), dtype=np.int64)
    sim.reverse = np.zeros((n_pairs, 2), dtype=np.uint8)
    sim.reverse[:, 1] = 1
    K = 2
    sim.events = np.zeros((n_pairs, 2, K, 2), dtype=np.int32)
    if seed_offsets is not None:
        sim.seed_offsets = tuple(seed_offsets)
        sim.seed_length = seed_length
        sim.seed_clean = np.zeros((n_pairs, 2, len(seed_offsets)), dtype=bool)
        sim.seed_shift = np.zeros((n_pairs, 2, len(seed_offsets)), dtype=np.int32)
    qvals = np.array([q for q, _ in quality_probs], dtype=np.uint8)
    qp = np.array([p for _, p in quality_probs], dtype=np.float64)
    qerr = np.power(10.0, -qvals.astype(np.float64) / 10.0)
    idx = np.arange(L, dtype=np.int64)[None, :]
    for b in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - b)
        contig = rng.choice(len(genome), size=n, p=weights).astype(np.uint32)
        ins = np.clip(np.rint(rng.normal(insert[0], insert[1], size=n)), insert[2], insert[3]).astype(np.int64)
        start = (rng.random(n) * (lens[contig] - margin)).astype(np.int64) + max_indel + 16
        pos = np.stack([start, start + ins - L], axis=1)            # forward leftmost base of each read
        for r in range(2):
            # indel events in strand-order read coordinates
            nev = np.minimum(rng.poisson(L * indel_rate, size=n), K)
            epos = np.sort(rng.integers(8, L - 8, size=(n, K)), axis=1)
            elen = np.minimum(rng.geometric(0.4, size=(n, K)), max_indel).astype(np.int64)
            isdel = rng.random((n, K)) < 0.5
            active = np.arange(K)[None, :] < nev[:, None]
            # keep the second event clear of the first one's inserted bases
            clash = active[:, 1] & (epos[:, 1] < epos[:, 0] + elen[:, 0] + 4)
            active[:, 1] &= ~clash
            goff = np.broadcast_to(idx, (n, L)).copy()
            inserted = np.zeros((n, L), dtype=bool)
            for k in range(K):
                a = active[:, k][:, None]
                p = epos[:, k][:, None]
                ln = elen[:, k][:, None]
                d = isdel[:, k][:, None]
                goff += np.where(a & d & (idx >= p), ln, 0)
                goff -= np.where(a & ~d, np.clip(idx - p, 0, ln), 0)
                inserted |= a & ~d & (idx >= p) & (idx < p + ln)
                sim.events[b:b + n, r, k, 0] = np.where(active[:, k], epos[:, k], -1)
                sim.events[b:b + n, r, k, 1] = np.where(active[:, k], np.where(isdel[:, k], elen[:, k], -elen[:, k]), 0)
            codes = np.empty((n, L), dtype=np.uint8)
            for c in np.unique(contig):
                m = contig == c
                codes[m] = _CODE[genome[c][pos[m, r][:, None] + goff[m]]]
            rnd = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
            codes = np.where(inserted | (codes > 3), rnd, codes)
            qi = rng.choice(len(qvals), size=(n, L), p=qp)
            q = qvals[qi]
            sub = rng.random((n, L)) < (snp_rate + qerr[qi])
            codes = np.where(sub, (codes + rng.integers(1, 4, size=(n, L), dtype=np.uint8)) & 3, codes).astype(np.uint8)
            byte = (q << 2) | codes
            byte = np.where(q == 2, 0, byte).astype(np.uint8)      # Q2 bases are read as N (bcl byte 0)
            if seed_offsets is not None:
                # a seed yields a match at the true locus iff none of its 32 bases is an error, an N or an inserted base
                # and no deletion falls inside it; its locus is shifted by the net indel length left of it
                dirty = sub | inserted | (q == 2)
                csum = np.concatenate([np.zeros((n, 1), dtype=np.int32), np.cumsum(dirty, axis=1, dtype=np.int32)], axis=1)
                for s, o in enumerate(seed_offsets):
                    fo = o if r == 0 else L - o - seed_length
                    clean = (csum[:, fo + seed_length] - csum[:, fo]) == 0
                    for k in range(K):
                        clean &= ~(active[:, k] & isdel[:, k] & (epos[:, k] > fo) & (epos[:, k] < fo + seed_length))
                    sim.seed_clean[b:b + n, r, s] = clean
                    sim.seed_shift[b:b + n, r, s] = goff[:, fo] - fo
            if r == 1:   # read 2 is sequenced from the other strand: reverse-complement into sequencing order
                byte = np.where(byte == 0, 0, byte ^ 3)[:, ::-1]
            sim.bcl[b:b + n, r * L:(r + 1) * L] = byte
        sim.contig[b:b + n] = contig[:, None]
        sim.position[b:b + n] = pos
    return sim


def insert_adapters(sim, sequences, fraction=0.3, seed=SEED_READS + 7, read_through=True, min_keep=20):
    """Puts sequencing adapters into a fraction of the simulated reads, in sequencing order (the BCL order): bases
    [cut, cut + len) become one of 'sequences' (chosen at random); with read_through the bases behind it are random (a short
    insert: the read runs through the adapter into the flowcell oligos), otherwise they stay as simulated (a mate-pair style
    junction adapter: both sides still match the genome).  Qualities and N calls are kept.  Returns the cut points
    (n_pairs, 2), -1 = untouched."""
    rng = _rng(seed)
    L = sim.L
    n = sim.bcl.shape[0]
    cuts = np.full((n, 2), -1, dtype=np.int32)
    codes = [np.array([_CODE[ord(ch)] for ch in s], dtype=np.uint8) for s in sequences]
    for r in range(2):
        view = sim.bcl[:, r * L:(r + 1) * L]
        pick = np.nonzero(rng.random(n) < fraction)[0]
        for i in pick:
            a = codes[int(rng.integers(0, len(codes)))]
            cut = int(rng.integers(min_keep, L - 4))
            cuts[i, r] = cut
            row = view[i]
            new = (row & 3).copy()
            m = min(a.size, L - cut)
            new[cut:cut + m] = a[:m]
            if read_through and cut + m < L:
                new[cut + m:] = rng.integers(0, 4, size=L - cut - m)
            view[i] = np.where(row == 0, 0, (row & 0xFC) | new)
    return cuts


def microbench_candidates(sim, genome, per_read=4, seed=SEED_READS + 2, fractions=(0.6, 0.2, 0.2)):
    """for every read 'per_read' (read, window) pairs: true locus / true locus shifted by
    +-(1..7) bp / uniform random locus with the given fractions.  Returns a CANDIDATE_DTYPE array."""
    from .types import CANDIDATE_DTYPE
    rng = _rng(seed)
    n_reads = sim.contig.size
    n = n_reads * per_read
    read_id = np.repeat(np.arange(n_reads, dtype=np.uint32), per_read)
    contig = sim.contig.reshape(-1)[read_id]
    true_pos = sim.position.reshape(-1)[read_id]
    reverse = sim.reverse.reshape(-1)[read_id].astype(np.uint32)
    kind = rng.random(n)
    shift = rng.integers(1, 8, size=n) * np.where(rng.random(n) < 0.5, -1, 1)
    lens = np.array([c.size for c in genome], dtype=np.int64)
    rnd = (rng.random(n) * (lens[contig] - sim.L - 64)).astype(np.int64) + 16
    pos = np.where(kind < fractions[0], true_pos, np.where(kind < fractions[0] + fractions[1], true_pos + shift, rnd))
    cand = np.empty(n, dtype=CANDIDATE_DTYPE)
    cand["position"] = pos
    cand["readId"] = read_id
    cand["contigStrand"] = (contig << 1) | reverse
    return cand


def rescue_policy(fragments, begin, n_clusters):
    """Stand-in for the decision TemplateBuilder takes between the two calls of the path (out of scope):
    a pair goes to buildDisjoinedTemplate whenever its best pair has any edit distance (TemplateBuilder.cpp:1073-1081),
    which rescues the mate of EVERY candidate of both reads (:737-757).  Here: every aligned fragment of a cluster
    becomes an orphan unless both reads of the cluster have a perfect (edit distance 0) candidate.  Both arms of the
    benchmark are driven by this same function.  Returns an isaac_ext_rescue_request_t array."""
    counts = np.diff(begin.astype(np.int64))
    list_of = np.repeat(np.arange(2 * n_clusters), counts)
    aligned = fragments["cigarLength"] > 0
    perfect = aligned & (fragments["editDistance"] == 0)
    read_perfect = np.bincount(list_of, weights=perfect, minlength=2 * n_clusters) > 0
    cluster_done = read_perfect[0::2] & read_perfect[1::2]
    need = aligned & ~cluster_done[list_of // 2]
    f = fragments[need]
    req = np.zeros(f.size, dtype=[("orphanPosition", "<i8"), ("bestTemplateLength", "<i8"), ("orphanReadId", "<u4"),
                                  ("orphanContigStrand", "<u4"), ("orphanObservedLength", "<u4"), ("pad", "<u4")])
    req["orphanPosition"] = f["position"]
    req["orphanReadId"] = f["readId"]
    req["orphanContigStrand"] = (f["contigId"] << 1) | f["reverse"]
    req["orphanObservedLength"] = f["observedLength"]
    return req


MATCH_DTYPE = np.dtype([("seedId", "<u8"), ("location", "<u8")])    # isaac_ext_match_t == the reference's Match
SEED_DTYPE = np.dtype([("offset", "<u2"), ("length", "<u2"), ("readIndex", "<u4")])   # isaac_ext_seed_t


def auto_seed_offsets(L, seed_length=32):
    """--seeds auto of the reference (SeedDescriptorOption.cpp:86-151): offset 0, the last full seed, then every
    seed_length bases in between: (0, 118, 32, 64) at 150 bp."""
    offs = [0, L - seed_length]
    o = seed_length
    while o + seed_length <= L - seed_length:
        offs.append(o)
        o += seed_length
    return tuple(offs)


def seed_table(sim):
    """isaac_ext_seed_t array: seeds of read 0 first, then read 1 (index = readIndex * S + k)."""
    S = len(sim.seed_offsets)
    seeds = np.zeros(2 * S, dtype=SEED_DTYPE)
    for r in range(2):
        for k, o in enumerate(sim.seed_offsets):
            seeds[r * S + k] = (o, sim.seed_length, r)
    return seeds


def make_matches(sim, genome, seed=SEED_READS + 3, decoy_rate=0.2, neighbor_rate=0.0, repeat_rate=0.0,
                 too_many_rate=0.0, repeat_matches=12):
    """Seed-free stand-in for the seed-matching stage: every error-free seed yields a Match at the true
    locus shifted by the net indel length left of it, plus decoys at uniform random loci; optionally seeds with
    neighbours, over-represented seeds and TooManyMatch records.  Returns (matches sorted like
    SelectMatchesTransition.cpp:242-254, clusterMatchBegin)."""
    rng = _rng(seed)
    n = sim.contig.shape[0]
    S = len(sim.seed_offsets)
    L, sl = sim.L, sim.seed_length
    lens = np.array([c.size for c in genome], dtype=np.int64)
    cl, rd, sd = np.nonzero(sim.seed_clean)                       # cluster, read, seed
    fo = np.where(rd == 0, np.array(sim.seed_offsets)[sd], L - np.array(sim.seed_offsets)[sd] - sl)
    position = sim.position[cl, rd] + fo + sim.seed_shift[cl, rd, sd]
    parts = [(cl, rd * S + sd, sim.reverse[cl, rd].astype(np.int64), sim.contig[cl, rd].astype(np.int64), position,
              (rng.random(cl.size) < neighbor_rate).astype(np.int64))]

    def random_matches(mask_rate, per):
        pick = np.nonzero(rng.random(2 * n) < mask_rate)[0]
        c, r = np.repeat(pick // 2, per), np.repeat(pick % 2, per)
        s = np.repeat(rng.integers(0, S, size=pick.size), per)
        contig = rng.integers(0, len(genome), size=c.size)
        pos = (rng.random(c.size) * (lens[contig] - L - 64)).astype(np.int64) + 16
        return (c, r * S + s, rng.integers(0, 2, size=c.size), contig, pos, np.zeros(c.size, dtype=np.int64))

    if decoy_rate > 0:
        parts.append(random_matches(decoy_rate, 1))
    if repeat_rate > 0:
        parts.append(random_matches(repeat_rate, repeat_matches))
    cluster = np.concatenate([p[0] for p in parts]).astype(np.uint64)
    seed_index = np.concatenate([p[1] for p in parts]).astype(np.uint64)
    reverse = np.concatenate([p[2] for p in parts]).astype(np.uint64)
    contig = np.concatenate([p[3] for p in parts]).astype(np.uint64)
    pos = np.concatenate([p[4] for p in parts]).astype(np.uint64)
    neighbors = np.concatenate([p[5] for p in parts]).astype(np.uint64)
    location = ((((contig + np.uint64(1)) << np.uint64(40)) | pos) << np.uint64(1)) | neighbors
    if too_many_rate > 0:
        pick = np.nonzero(rng.random(2 * n) < too_many_rate)[0]
        cluster = np.concatenate([cluster, (pick // 2).astype(np.uint64)])
        seed_index = np.concatenate([seed_index, ((pick % 2) * S + rng.integers(0, S, size=pick.size)).astype(np.uint64)])
        reverse = np.concatenate([reverse, np.zeros(pick.size, dtype=np.uint64)])
        location = np.concatenate([location, np.zeros(pick.size, dtype=np.uint64)])      # ReferencePosition::TooManyMatch
    order = np.lexsort((seed_index, location, cluster))
    m = np.empty(order.size, dtype=MATCH_DTYPE)
    m["seedId"] = (cluster[order] << np.uint64(9)) | (seed_index[order] << np.uint64(1)) | reverse[order]
    m["location"] = location[order]
    begin = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(np.bincount(cluster.astype(np.int64), minlength=n), out=begin[1:])
    return m, begin


# ---- bench-size workloads generated on several host cores (fork workers; call these BEFORE the process creates a CUDA context) ----
_SHARED = {}


def _genome_job(job):
    c, n, seed, n_fraction, n_run = job
    a = _SHARED["contigs"][c]
    a[:] = make_genome(n, 1, seed=seed, n_fraction=n_fraction, n_run=n_run)[0]
    return c


def make_genome_parallel(total_bases, n_contigs=24, seed=SEED_G3100, n_fraction=0.001, n_run=(100, 10000), workers=8):
    """like make_genome, one generator stream per contig (seed + contig index) so that the contigs can be
    drawn side by side by fork workers that write into anonymous shared mappings."""
    import mmap
    import multiprocessing
    per = -(-total_bases // n_contigs)
    sizes = [min(per, total_bases - c * per) for c in range(n_contigs)]
    contigs = [np.frombuffer(mmap.mmap(-1, max(1, n)), dtype=np.uint8, count=n) for n in sizes]
    _SHARED["contigs"] = contigs
    jobs = [(c, sizes[c], seed + c, n_fraction, n_run) for c in range(n_contigs)]
    if workers <= 1:
        for j in jobs:
            _genome_job(j)
    else:
        with multiprocessing.get_context("fork").Pool(min(workers, n_contigs)) as pool:
            pool.map(_genome_job, jobs, chunksize=1)
    _SHARED.pop("contigs")
    return contigs


_SIM_FIELDS = ("bcl", "contig", "position", "reverse", "events", "seed_clean", "seed_shift")


def _simulate_job(job):
    n, seed, kw = job
    sim = simulate_pairs(_SHARED["genome"], n, seed=seed, **kw)
    return {f: getattr(sim, f) for f in _SIM_FIELDS if hasattr(sim, f)}


def simulate_pairs_parallel(genome, n_pairs, seed=SEED_READS, workers=8, min_batch=50_000, **kw):
    """simulate_pairs over sub-batches with their own generator streams (seed + 7919 * batch), drawn by fork workers and
    concatenated in batch order: the result depends on (seed, workers, min_batch, n_pairs), not on scheduling."""
    import multiprocessing
    batches = max(1, min(workers, n_pairs // max(1, min_batch)))
    sizes = [n_pairs * (k + 1) // batches - n_pairs * k // batches for k in range(batches)]
    jobs = [(sizes[k], seed + 7919 * k, kw) for k in range(batches)]
    _SHARED["genome"] = genome
    if batches == 1:
        parts = [_simulate_job(jobs[0])]
    else:
        with multiprocessing.get_context("fork").Pool(batches) as pool:
            parts = pool.map(_simulate_job, jobs, chunksize=1)
    _SHARED.pop("genome")
    sim = Simulation(kw.get("L", 150))
    for f in parts[0]:
        setattr(sim, f, np.concatenate([p[f] for p in parts], axis=0))
    if kw.get("seed_offsets") is not None:
        sim.seed_offsets = tuple(kw["seed_offsets"])
        sim.seed_length = kw.get("seed_length", 32)
    return sim
