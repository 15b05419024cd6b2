"""The band-width-parametrised model of BandedSmithWaterman::align (oracle/isaac_oracle.cpp: BandedSwT<WIDTH>), the checker of the
warp-wavefront kernel with the widened band (BASELINE configs[4]; the reference hard-wires 16 lanes, BandedSmithWaterman.hh:88-89,
so a wider band has no reference behaviour to match).  Pinned here: its WIDTH = 16 instance equals the reference's OWN code
(oracle/_ref) on 120 000 random cases with both score presets; its wider instances find what a wider band must find."""
import os

import numpy as np
import pytest

import oracle_lib
from common import random_sw_cases
from isaac_aligner_b200.types import cigar_to_string

HAVE_REF = os.path.exists(oracle_lib.REF_SO) or os.path.isdir("/root/reference/src/c++")


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference build")
@pytest.mark.parametrize("scores,seed,n", [((0, -3, 11, 4), 611, 60000), ((2, -1, 15, 3), 612, 60000)])
def test_width_16_equals_the_reference(scores, seed, n):
    ref, port = oracle_lib.reference(), oracle_lib.port()
    queries, dbs = random_sw_cases(n, seed=seed)
    want = ref.banded_sw(queries, dbs, scores, max_read_length=300, threads=8)
    got = port.banded_sw(queries, dbs, scores, max_read_length=300, threads=8, band=16)
    for w, g in zip(want, got):
        assert np.array_equal(w, g)


def test_wider_band_contains_the_narrow_one():
    """an alignment the 16-lane band finds without touching its edge lanes is found by the 32-lane band too, shifted by the
    eight extra database bases in front"""
    port = oracle_lib.port()
    rng = np.random.default_rng(5)
    L = 120
    for trial in range(200):
        genome = bytes(b"ACGT"[int(c)] for c in rng.integers(0, 4, size=L + 200))
        at = 100
        q = bytearray(genome[at:at + L])
        if trial % 2:
            del q[40:43]
            q += genome[at + L:at + L + 3]                 # a 3-base deletion in the read
        d16, d32 = genome[at - 8:at - 8 + L + 15], genome[at - 16:at - 16 + L + 31]
        c16, l16, o16 = port.banded_sw([bytes(q)], [d16], (0, -3, 11, 4), band=16)
        c32, l32, o32 = port.banded_sw([bytes(q)], [d32], (0, -3, 11, 4), band=32)
        assert cigar_to_string(c16[0][:l16[0]]) == cigar_to_string(c32[0][:l32[0]])
        assert int(o16[0]) + 8 == int(o32[0])


def test_wider_band_finds_longer_gaps():
    """a 20-base deletion lies outside a 16-lane band and inside a 32-lane one"""
    port = oracle_lib.port()
    rng = np.random.default_rng(9)
    L = 150
    genome = bytes(b"ACGT"[int(c)] for c in rng.integers(0, 4, size=600))
    at = 200
    q = genome[at:at + 70] + genome[at + 90:at + 90 + 80]                  # 70M20D80M
    d32 = genome[at - 5:at - 5 + L + 31]
    c32, l32, o32 = port.banded_sw([q], [d32], (0, -3, 11, 4), band=32)
    got = cigar_to_string(c32[0][:l32[0]])
    assert got.count("D") == 1 and "20D" in got and int(o32[0]) == 5, got      # the gap may sit a base or two off where the flanks repeat
    assert sum(int(w >> 4) for w in c32[0][:l32[0]] if (w & 0xF) == 0) == L
    d16 = genome[at - 5:at - 5 + L + 15]
    c16, l16, o16 = port.banded_sw([q], [d16], (0, -3, 11, 4), band=16)
    assert "20D" not in cigar_to_string(c16[0][:l16[0]])
