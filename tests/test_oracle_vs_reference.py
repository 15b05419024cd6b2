"""Pins the CPU restatement (oracle/isaac_oracle.cpp) against the reference's own sources compiled unmodified
(oracle/_ref/libisaac_ref.so, built by oracle/Makefile from /root/reference).  Skipped where the reference library
was never built."""
import numpy as np
import pytest

import oracle_lib
from common import assert_fragments_equal, random_sw_cases, small_workload
from isaac_aligner_b200.types import BWA_SCORES, ELAND_SCORES, Config

REF = oracle_lib.reference()
PORT = oracle_lib.port()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/libisaac_ref.so not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("scores", [(0, -3, 11, 4), (2, -1, 15, 3)])
def test_banded_sw_port_matches_reference(scores):
    queries, dbs = random_sw_cases(6000, seed=11 + scores[0])
    cr, lr, orf = REF.banded_sw(queries, dbs, scores, max_read_length=300)
    cp, lp, opf = PORT.banded_sw(queries, dbs, scores, max_read_length=300)
    assert np.array_equal(lr, lp)
    assert np.array_equal(orf, opf)
    assert np.array_equal(cr, cp)
    # the set must exercise gaps, and the _mm_max_epi16-on-bytes quirk must matter for some of them
    ops = cr & 0xF
    assert (ops == 1).any() and (ops == 2).any()


@needs_ref
@pytest.mark.parametrize("scores", [BWA_SCORES, ELAND_SCORES])
def test_ungapped_port_matches_reference(scores):
    genome, sim, reads, cand = small_workload(seed=21)
    g = oracle_lib.GenomeHolder(genome)
    cfg = Config.default(scores, max_read_length=200)
    fr, cr, mr = REF.ungapped(g, reads, cfg, cand)
    fp, cp, mp = PORT.ungapped(g, reads, cfg, cand)
    assert_fragments_equal(fr, fp, cr, cp, mr, mp, "ungapped port vs reference")
    assert (fr["cigarLength"] == 0).any() and (fr["mismatchCount"] > 20).any() and (fr["mismatchCount"] == 0).any()


@needs_ref
@pytest.mark.parametrize("scores", [BWA_SCORES, ELAND_SCORES])
def test_gapped_port_matches_reference(scores):
    genome, sim, reads, cand = small_workload(seed=33, indel_rate=6e-3)
    g = oracle_lib.GenomeHolder(genome)
    cfg = Config.default(scores, max_read_length=200)
    fr, cr, mr = REF.gapped(g, reads, cfg, cand)
    fp, cp, mp = PORT.gapped(g, reads, cfg, cand)
    assert_fragments_equal(fr, fp, cr, cp, mr, mp, "gapped port vs reference")
    assert (fr["gapCount"] > 0).any() and (fr["matchCount"] == 0).any()
