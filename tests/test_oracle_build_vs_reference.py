"""Pins the restatement of FragmentBuilder::build (incl. SimpleIndelAligner and the gapped acceptance rules) and of
ShadowAligner::rescueShadow against the reference's own code (oracle/_ref/libisaac_ref.so)."""
import numpy as np
import pytest

import oracle_lib
from common_build import assert_flat_equal, build_workload, rescue_requests
from isaac_aligner_b200.batch import FRm, FRp, RFm, RFp, Tls
from isaac_aligner_b200.types import BWA_SCORES, ELAND_SCORES, Config

REF = oracle_lib.reference()
PORT = oracle_lib.port()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/libisaac_ref.so not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("scores,L,with_gaps", [(BWA_SCORES, 150, True), (ELAND_SCORES, 100, True), (BWA_SCORES, 100, False)])
def test_build_port_matches_reference(scores, L, with_gaps):
    genome, sim, reads, mb = build_workload(n_pairs=1500, L=L, seed=40 + L, indel_rate=5e-3, with_gaps=with_gaps)
    g = oracle_lib.GenomeHolder(genome)
    cfg = Config.default(scores, max_read_length=2 * L)
    r = oracle_lib.build_fragments(REF, g, reads, cfg, mb)
    p = oracle_lib.build_fragments(PORT, g, reads, cfg, mb)
    assert_flat_equal(r, p, "build port vs reference")
    f = r.fragments
    assert (f["gapCount"] > 0).sum() > 20, "the workload must exercise simple indels / gapped alignment"
    assert (f["uniqueSeedCount"] > 1).any() and (f["repeatSeedsCount"] > 0).any() and (f["nonUniqueSeedOffsetFirst"] != 0xFFFF).any()
    assert (r.flags == 0).any() and (r.flags == 1).any()


@needs_ref
def test_build_threads_do_not_change_results():
    genome, sim, reads, mb = build_workload(n_pairs=800, L=100, seed=91)
    g = oracle_lib.GenomeHolder(genome)
    cfg = Config.default(BWA_SCORES, max_read_length=200)
    assert_flat_equal(oracle_lib.build_fragments(PORT, g, reads, cfg, mb, threads=1),
                      oracle_lib.build_fragments(PORT, g, reads, cfg, mb, threads=4), "threads")


@needs_ref
@pytest.mark.parametrize("scores,L,models", [(BWA_SCORES, 150, (FRp, RFm)), (ELAND_SCORES, 100, (FRp, RFm)), (BWA_SCORES, 100, (RFp, FRm))])
def test_rescue_port_matches_reference(scores, L, models):
    genome, sim, reads, mb = build_workload(n_pairs=600, L=L, seed=60 + L, indel_rate=6e-3)
    g = oracle_lib.GenomeHolder(genome)
    cfg = Config.default(scores, max_read_length=2 * L)
    tls = Tls.make(m0=models[0], m1=models[1])
    req = rescue_requests(sim, seed=61)
    r = oracle_lib.rescue_shadows(REF, g, reads, cfg, tls, req)
    p = oracle_lib.rescue_shadows(PORT, g, reads, cfg, tls, req)
    assert_flat_equal(r, p, "rescue port vs reference")
    if models == (FRp, RFm):
        assert r.flags.mean() > 0.8 and (r.fragments["gapCount"] > 0).any()
        assert (np.diff(r.begin.astype(np.int64)) > 1).any()
