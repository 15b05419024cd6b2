"""The reference's first FragmentBuilder unit test (lib/alignment/cppunit/testFragmentBuilder.cpp:90-597) restated on the flat
build call: its fixture (BuilderInit.hh:121-172: noise contigs of 210, 220, 230, 5 + 230, 60 bases; reads cut out of them, Q40),
its match lists and every value it asserts -- empty match list, one seed per read, seed offsets, three seeds consolidated into one
fragment (uniqueSeedCount 3), the same read on two contigs (c3 = "AAAAA" + c2), two mismatches with their log probabilities to
1e-9, leading / trailing / both soft clips against the 60-base contig.  The asserted values depend on the geometry only, so noise
from a fixed seed stands in for the fixture's rand().  Replayed through both CPU checkers and, on the GPU, through
isaac_ext_build_fragments."""
import os

import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200.batch import MatchBatch
from isaac_aligner_b200.synth import MATCH_DTYPE, SEED_DTYPE
from isaac_aligner_b200.types import ELAND_SCORES, Config, ReadSet

ALIGN, SOFT_CLIP = 0, 4
COMPLEMENT = np.array([3, 2, 1, 0], dtype=np.uint8)
SEEDS = [(0, 32, 0), (32, 32, 0), (64, 32, 0), (0, 32, 1), (32, 32, 1), (64, 32, 1)]     # getSeedMetadataList (BuilderInit.hh:34-46)


def rc(codes):
    return COMPLEMENT[codes[::-1]]


def contigs():
    rng = np.random.default_rng(2102202305)
    c2 = rng.integers(0, 4, size=230).astype(np.uint8)
    c3 = np.concatenate([np.zeros(5, dtype=np.uint8), c2])
    c4 = rng.integers(0, 4, size=60).astype(np.uint8)
    c1 = rng.integers(0, 4, size=220).astype(np.uint8)
    c0 = rng.integers(0, 4, size=210).astype(np.uint8)
    c0[6], c0[111] = 3, 3              # testMismatches replaces a 'T' of read 1 by 'A' and an 'A' of read 2 by 'C' (:44-48)
    return [c0, c1, c2, c3, c4]


def get_bcl(contig, offset0, offset1):
    """getBcl(readMetadataList, contigList, contigId, offset0, offset1, false, true) (BuilderInit.hh:147-172)"""
    bases = np.concatenate([contig[offset0:offset0 + 100], rc(contig)[offset1:offset1 + 100]])
    return ((40 << 2) | bases).astype(np.uint8)


def clusters(c):
    bcl0 = get_bcl(c[0], 2, 3)
    bcl3 = bcl0.copy()
    bcl3[4], bcl3[195] = ord("x"), ord("Q")                                                # 30 << 2 | A, 20 << 2 | C (:44-48)
    assert (bcl0[4] & 3) != (bcl3[4] & 3) and (bcl0[195] & 3) != (bcl3[195] & 3)          # :375-381
    q40 = lambda codes: ((40 << 2) | codes).astype(np.uint8)
    c4, r4 = c[4], rc(c[4])
    return {
        "cluster0": bcl0, "cluster2": get_bcl(c[2], 1, 2), "cluster3": bcl3,
        "cluster4l": q40(np.concatenate([c4[0:44], c4[0:56], r4[0:42], r4[0:58]])),       # :49-50
        "cluster4t": q40(np.concatenate([c4[16:60], c4[0:56], r4[18:60], r4[0:58]])),     # :51-52
        "cluster4lt": q40(np.concatenate([np.zeros(10, np.uint8), c4, np.full(30, 1, np.uint8), np.full(15, 2, np.uint8), r4,
                                          np.full(25, 3, np.uint8)])),                    # :53-54
    }


def match(seed, reverse, contig, position):
    """Match(SeedId(tile, 0, cluster, seed, reverse), ReferencePosition(contig, position))"""
    return ((seed << 1) | (1 if reverse else 0), (((contig + 1) << 40) | position) << 1)


def word(length, op):
    return (length << 4) | op


LP_PERFECT = (-0.0100005, 1e-6)
# name, cluster, repeat threshold, matches, expected fragments per read: (contig, uniqueSeedCount or None, position, observedLength,
# reverse, cigar words, mismatchCount, (logProbability, tolerance) or None)
SCENARIOS = [
    ("testEmptyMatchList", "cluster0", 123, [], [[], []]),                                                          # :90-108
    ("testSingleSeed", "cluster0", 456, [match(0, False, 0, 2), match(3, True, 0, 175)],                           # :110-189
     [[(0, 1, 2, 100, False, [word(100, ALIGN)], 0, LP_PERFECT)], [(0, 1, 107, 100, True, [word(100, ALIGN)], 0, LP_PERFECT)]]),
    ("testSeedOffset", "cluster0", 456, [match(1, False, 0, 2 + 32), match(5, True, 0, 175 - 64)],                 # :191-195
     [[(0, 1, 2, 100, False, [word(100, ALIGN)], 0, LP_PERFECT)], [(0, 1, 107, 100, True, [word(100, ALIGN)], 0, LP_PERFECT)]]),
    ("testMultiSeed", "cluster0", 123,                                                                              # :197-264
     [match(0, False, 0, 2), match(1, False, 0, 34), match(2, False, 0, 66), match(3, True, 0, 175), match(4, True, 0, 143), match(5, True, 0, 111)],
     [[(0, 3, 2, 100, False, [word(100, ALIGN)], 0, LP_PERFECT)], [(0, 3, 107, 100, True, [word(100, ALIGN)], 0, LP_PERFECT)]]),
    ("testRepeats", "cluster2", 123,                                                                                # :266-371
     [match(0, False, 2, 1), match(1, False, 2, 33), match(2, False, 2, 65), match(3, True, 2, 196), match(4, True, 2, 164), match(5, True, 2, 132),
      match(0, False, 3, 6), match(2, False, 3, 70), match(3, True, 3, 201), match(4, True, 3, 169)],
     [[(2, 3, 1, 100, False, [word(100, ALIGN)], 0, LP_PERFECT), (3, 2, 6, 100, False, [word(100, ALIGN)], 0, LP_PERFECT)],
      [(2, 3, 128, 100, True, [word(100, ALIGN)], 0, LP_PERFECT), (3, 2, 133, 100, True, [word(100, ALIGN)], 0, None)]]),
    ("testMismatches", "cluster3", 123, [match(0, False, 0, 2), match(3, True, 0, 175)],                           # :373-460
     [[(0, 1, 2, 100, False, [word(100, ALIGN)], 1, (-8.016268063, 1e-9))], [(0, 1, 107, 100, True, [word(100, ALIGN)], 1, (-5.713682970, 1e-9))]]),
    ("testLeadingSoftClips", "cluster4l", 123, [match(2, False, 4, 20), match(5, True, 4, 6)],                     # :462-505
     [[(4, None, 0, 56, False, [word(44, SOFT_CLIP), word(56, ALIGN)], 0, None)], [(4, None, 2, 58, True, [word(58, ALIGN), word(42, SOFT_CLIP)], 0, None)]]),
    ("testTrailingSoftClips", "cluster4t", 123, [match(0, False, 4, 16), match(3, True, 4, 10)],                   # :507-550
     [[(4, None, 16, 44, False, [word(44, ALIGN), word(56, SOFT_CLIP)], 0, None)], [(4, None, 0, 42, True, [word(58, SOFT_CLIP), word(42, ALIGN)], 0, None)]]),
    ("testLeadingAndTrailingSoftClips", "cluster4lt", 123, [match(1, False, 4, 22), match(4, True, 4, 11)],        # :552-597
     [[(4, None, 0, 60, False, [word(10, SOFT_CLIP), word(60, ALIGN), word(30, SOFT_CLIP)], 0, None)],
      [(4, None, 0, 60, True, [word(25, SOFT_CLIP), word(60, ALIGN), word(15, SOFT_CLIP)], 0, None)]]),
]


def replay(build):
    """build(genome, reads, config, match_batch) -> FlatFragments"""
    codes = contigs()
    genome = [np.frombuffer(b"ACGT", dtype=np.uint8)[c] for c in codes]
    bcl = clusters(codes)
    seeds = np.array(SEEDS, dtype=SEED_DTYPE)
    for name, cluster, repeat_threshold, matches, expected in SCENARIOS:
        reads = ReadSet(bcl[cluster][None, :], (100, 100))
        m = np.zeros(len(matches), dtype=MATCH_DTYPE)
        for i, (seed_id, location) in enumerate(matches):
            m[i] = (seed_id, location)
        mb = MatchBatch(m, np.array([0, len(matches)], dtype=np.uint64), seeds, with_gaps=True)
        config = Config.default(ELAND_SCORES, max_read_length=200)       # FragmentBuilder(flowcells, threshold, 3, 8, false, ELAND..., 20000)
        config.repeatThreshold, config.maxSeedsPerRead, config.gappedMismatchesMax, config.semialignedGapLimit = repeat_threshold, 3, 8, 20000
        flat = build(genome, reads, config, mb)
        assert bool(flat.flags[0]) == bool(matches), name
        for r in range(2):
            got = flat.fragments[int(flat.begin[r]):int(flat.begin[r + 1])]
            assert len(got) == len(expected[r]), (name, r, len(got))
            for k, (contig, unique, position, observed, reverse, cigar, mismatches, lp) in enumerate(expected[r]):
                f, what = got[k], (name, r, k)
                assert (int(f["contigId"]), int(f["position"]), int(f["observedLength"]), bool(f["reverse"]), int(f["readIndex"]),
                        int(f["mismatchCount"])) == (contig, position, observed, reverse, r, mismatches), what
                if unique is not None:
                    assert int(f["uniqueSeedCount"]) == unique, what
                assert flat.cigar(int(flat.begin[r]) + k).tolist() == cigar, what
                if lp is not None:
                    assert abs(float(f["logProbability"]) - lp[0]) <= lp[1], what
        if name == "testEmptyMatchList":
            assert flat.cigars.size == 0, name


def test_restatement_reproduces_the_fragment_builder_scenarios():
    chk = oracle_lib.port()
    replay(lambda genome, reads, config, mb: oracle_lib.build_fragments(chk, oracle_lib.GenomeHolder(genome), reads, config, mb))


def test_reference_build_reproduces_the_fragment_builder_scenarios():
    if not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"):
        pytest.skip("the reference build of the checker did not travel to this box")
    chk = oracle_lib.reference()
    replay(lambda genome, reads, config, mb: oracle_lib.build_fragments(chk, oracle_lib.GenomeHolder(genome), reads, config, mb))


@pytest.mark.gpu
def test_cuda_reproduces_the_fragment_builder_scenarios():
    from isaac_aligner_b200 import capi

    def build(genome, reads, config, mb):
        ctx = capi.Context(config)
        ctx.set_reference(genome)
        ctx.set_reads(reads)
        flat = ctx.build_fragments(mb)
        ctx.close()
        return flat

    replay(build)
