"""The tile driver (isaac_aligner_b200/tile.py = MatchSelector::parallelSelect for one tile): the reference's
on-disk inputs (raw Match records, BclClusters bytes) in, templates + template length statistics + MatchSelectorStats summary
out; checked against the same steps composed from the reference's own classes."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib
from common_build import build_workload
from isaac_aligner_b200 import tile
from isaac_aligner_b200.batch import MatchBatch, TemplateOptions, Tls
from isaac_aligner_b200.types import BWA_SCORES, Config, ReadSet
from test_gpu_tls import swap_reads, words


def test_cluster_match_begin_and_file_round_trip(tmp_path):
    genome, sim, reads, mb = build_workload(n_pairs=300, L=100, seed=21)
    path = os.path.join(str(tmp_path), "matches.dat")
    tile.write_match_file(path, mb.matches)
    assert os.path.getsize(path) == 16 * len(mb.matches)
    back = tile.read_match_file(path)
    assert back.tobytes() == mb.matches.tobytes()
    assert np.array_equal(tile.cluster_match_begin(back, 300), mb.begin)
    bcl_path = os.path.join(str(tmp_path), "clusters.bcl")
    sim.bcl.tofile(bcl_path)
    assert np.array_equal(tile.read_bcl_clusters(bcl_path, (100, 100)), sim.bcl)
    with pytest.raises(ValueError):
        tile.cluster_match_begin(back[::-1], 300)


@pytest.mark.gpu
@pytest.mark.parametrize("user_tls", [False, True])
def test_select_tile_matches_the_reference_steps(tmp_path, user_tls):
    if not os.path.exists(oracle_lib.REF_SO):
        pytest.skip("the reference build of the checker did not travel to this box")
    from isaac_aligner_b200 import capi
    chk = oracle_lib.Oracle(oracle_lib.REF_SO)
    n, L, cutoff = 5000, 100, 25
    genome, sim, reads, mb = build_workload(n_pairs=n, L=L, seed=611, genome_bases=1_000_000, indel_rate=3e-3, masked=False)
    reads, mb = swap_reads(sim, reads, mb)
    path = os.path.join(str(tmp_path), "matches.dat")
    tile.write_match_file(path, mb.matches)
    reads.bcl.tofile(os.path.join(str(tmp_path), "clusters.bcl"))
    pf = (np.random.default_rng(2).random(n) < 0.9).astype(np.uint8)
    cfg = Config.default(BWA_SCORES, max_read_length=2 * L)
    options = TemplateOptions.make(clip_semialigned=True, clip_overlapping=True)
    tls_in = Tls.make() if user_tls else None
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    got = tile.select_tile(ctx, tile.read_bcl_clusters(os.path.join(str(tmp_path), "clusters.bcl"), (L, L)), (L, L),
                           tile.read_match_file(path), mb.seeds, pf=pf, base_quality_cutoff=cutoff, tls=tls_in, options=options, cycle_stats=True)
    ctx.close()
    # the same steps through the reference's own code
    g = oracle_lib.GenomeHolder(genome)
    masked = oracle_lib.trim_low_quality_ends(chk, ReadSet(reads.bcl, (L, L)), cutoff)
    assert np.array_equal(got.end_cycles_masked, masked) and masked.any()
    trimmed = ReadSet(reads.bcl, (L, L), end_cycles_masked=masked)
    want_tls, stable = (tls_in, True) if user_tls else oracle_lib.determine_template_length(chk, g, trimmed, cfg, mb, pf)
    assert words(got.tls) == words(want_tls) and got.tls_stable == stable
    want = oracle_lib.build_templates(chk, g, trimmed, cfg, mb, want_tls, options, threads=8)
    for name in ("built", "properPair", "alignmentScore", "fragmentAlignmentScore"):
        assert np.array_equal(got.templates.templates[name], want.templates[name]), name
    for name in ("position", "contigId", "observedLength", "editDistance", "cigarLength", "lowClipped", "highClipped", "mismatchCount"):
        assert np.array_equal(got.templates.fragments[name], want.fragments[name]), name
    assert np.array_equal(got.stats, oracle_lib.template_stats(chk, g, trimmed, cfg, mb, want_tls, options, pf, threads=8))
    assert got.stats[0][3] == n and 0 < got.stats[1][3] < n
    assert np.array_equal(got.cycle_stats, oracle_lib.tile_cycle_stats(chk, g, trimmed, cfg, mb, want_tls, options, pf, threads=8))
