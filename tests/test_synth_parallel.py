"""The bench-size workload generators that draw on several host cores (synth.make_genome_parallel / simulate_pairs_parallel):
the result depends on the seeds and the batch layout only, not on the number of workers or on scheduling."""
import numpy as np

from isaac_aligner_b200 import synth


def test_parallel_genome_is_per_contig_seeded():
    a = synth.make_genome_parallel(3_000_000, n_contigs=6, seed=11, n_fraction=0.002, workers=1)
    b = synth.make_genome_parallel(3_000_000, n_contigs=6, seed=11, n_fraction=0.002, workers=3)
    assert len(a) == len(b) == 6 and all(x.size == 500_000 for x in a)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert all(np.array_equal(a[c], synth.make_genome(500_000, 1, seed=11 + c, n_fraction=0.002)[0]) for c in range(6))
    n = np.mean([(x == ord("N")).mean() for x in a])
    assert 0 < n < 0.05 and set(np.unique(a[0])) <= set(b"ACGTN")       # runs of 100 .. 10 000 bases overshoot a small contig's target


def test_parallel_simulation_is_the_concatenation_of_its_batches():
    genome = synth.make_genome(400_000, n_contigs=2, seed=5)
    offs = synth.auto_seed_offsets(100)
    sim = synth.simulate_pairs_parallel(genome, 3001, seed=21, workers=3, min_batch=500, L=100, seed_offsets=offs)
    assert sim.bcl.shape == (3001, 200) and sim.seed_clean.shape == (3001, 2, len(offs)) and sim.seed_offsets == offs
    sizes = [3001 * (k + 1) // 3 - 3001 * k // 3 for k in range(3)]
    at = 0
    for k, size in enumerate(sizes):
        part = synth.simulate_pairs(genome, size, seed=21 + 7919 * k, L=100, seed_offsets=offs)
        assert np.array_equal(sim.bcl[at:at + size], part.bcl) and np.array_equal(sim.position[at:at + size], part.position)
        at += size
    matches, begin = synth.make_matches(sim, genome, seed=3)
    assert begin[-1] == len(matches) > 3001
