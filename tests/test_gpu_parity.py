"""Parity of the CUDA path (through the C ABI, libisaac_ext.so) with the CPU oracle, bit-exact.
The checker is the scalar restatement oracle/isaac_oracle.cpp and, when it travelled to this box, the reference's own
code (oracle/_ref/libisaac_ref.so).  /root/reference is never read here."""
import os

import numpy as np
import pytest

import oracle_lib
from common import assert_fragments_equal, random_sw_cases, small_workload
from isaac_aligner_b200.types import BWA_SCORES, ELAND_SCORES, Config

pytestmark = pytest.mark.gpu


def checkers():
    return oracle_lib.gpu_checkers()        # fails when the reference build did not travel


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


@pytest.mark.parametrize("scores", [(0, -3, 11, 4), (2, -1, 15, 3),
                                    (70, -20, 105, 105)])    # match * length too large for row-relative cell values: the plain kernel
def test_banded_sw_bit_exact(capi, scores):
    queries, dbs = random_sw_cases(20000, seed=101 + scores[0])
    ctx = capi.Context(Config.default(max_read_length=300))
    cg, lg, og = ctx.banded_sw(queries, dbs, scores)
    for chk in checkers():
        cr, lr, orf = chk.banded_sw(queries, dbs, scores, max_read_length=300, threads=8)
        assert np.array_equal(lg, lr), chk.kind
        assert np.array_equal(og, orf), chk.kind
        assert np.array_equal(cg, cr), chk.kind
    ctx.close()


@pytest.mark.parametrize("scores,L", [(BWA_SCORES, 100), (ELAND_SCORES, 100), (BWA_SCORES, 150), (BWA_SCORES, 250)])
def test_ungapped_bit_exact(capi, scores, L):
    genome, sim, reads, cand = small_workload(n_pairs=3000, L=L, seed=55 + L)
    ctx = capi.Context(Config.default(scores, max_read_length=2 * L))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    fg, cg, mg = ctx.ungapped(cand)
    g = oracle_lib.GenomeHolder(genome)
    for chk in checkers():
        fr, cr, mr = chk.ungapped(g, reads, ctx.config, cand, threads=8)
        assert_fragments_equal(fg, fr, cg, cr, mg, mr, "ungapped cuda vs " + chk.kind)
    ctx.close()


@pytest.mark.parametrize("scores,L", [(BWA_SCORES, 100), (ELAND_SCORES, 100), (BWA_SCORES, 150), (BWA_SCORES, 250)])
def test_gapped_bit_exact(capi, scores, L):
    genome, sim, reads, cand = small_workload(n_pairs=3000, L=L, seed=77 + L, indel_rate=6e-3)
    ctx = capi.Context(Config.default(scores, max_read_length=2 * L))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    # GappedAligner is only ever handed fragments whose ungapped alignment kept a match (FragmentBuilder.cpp:179)
    cand = cand[ctx.ungapped(cand)[0]["cigarLength"] > 0]
    fg, cg, mg = ctx.gapped(cand)
    g = oracle_lib.GenomeHolder(genome)
    for chk in checkers():
        fr, cr, mr = chk.gapped(g, reads, ctx.config, cand, threads=8)
        assert_fragments_equal(fg, fr, cg, cr, mg, mr, "gapped cuda vs " + chk.kind)
    assert (fg["gapCount"] > 0).any()
    ctx.close()


def test_errors_are_loud(capi):
    # the reference throws InvalidParameterException for overflow-prone score/length combinations
    # (BandedSmithWaterman.cpp:47-53, testBandedSmithWaterman.cpp:214-225)
    with pytest.raises(capi.ExtError) as e:
        capi.Context(Config.default((2, -1, -15, -3, -25), max_read_length=1024 * 3))
    assert e.value.code == 1
    ctx = capi.Context(Config.default(max_read_length=300))
    with pytest.raises(capi.ExtError) as e:
        ctx.ungapped(np.zeros(1, dtype=capi.CANDIDATE_DTYPE))
    assert e.value.code == 6      # no reference yet
    ctx.close()


@pytest.mark.parametrize("gapped", [False, True])
def test_compact_end_to_end_variant_matches_fixed_stride(capi, gapped):
    """isaac_ext_*_batch_compact: same records, CIGARs in a dense pool; several chunks and an odd tail"""
    genome, sim, reads, cand = small_workload(n_pairs=3000, L=100, seed=91, indel_rate=6e-3)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    cand = cand[ctx.ungapped(cand)[0]["cigarLength"] > 0]
    cand = np.concatenate([cand] * 130)[:(1 << 21) + 12345]          # > 2 chunks of 2^20
    f0, c0, _ = (ctx.gapped(cand, with_masks=False) if gapped else ctx.ungapped(cand, with_masks=False))
    f1 = np.zeros(len(cand), dtype=capi.FRAGMENT_DTYPE)
    pool = np.zeros(len(cand) * 10, dtype=np.uint32)
    words = ctx.extend_compact(cand, gapped, f1, pool)
    assert words == int(f0["cigarLength"].sum())
    for name in capi.FRAGMENT_DTYPE.names:
        if name != "cigarOffset":
            assert np.array_equal(f0[name], f1[name]), name
    off = np.concatenate([[0], np.cumsum(f0["cigarLength"][:-1], dtype=np.uint64)])
    assert np.array_equal(f1["cigarOffset"], off.astype(np.uint32))
    idx = np.random.default_rng(3).integers(0, len(cand), size=5000)
    for i in idx:
        n = int(f0["cigarLength"][i])
        assert np.array_equal(c0[i][:n], pool[int(f1["cigarOffset"][i]):int(f1["cigarOffset"][i]) + n])
    # a pool that is too small is reported, with the required size
    with pytest.raises(capi.ExtError) as e:
        ctx.extend_compact(cand, gapped, f1, pool[:1000])
    assert e.value.code == 5
    ctx.close()


@pytest.mark.parametrize("gapped", [False, True])
def test_compact_rejects_bad_candidates(capi, gapped):
    """the chunked entry points validate on the device: one bad candidate anywhere fails the whole call, and the next call is clean"""
    genome, sim, reads, cand = small_workload(n_pairs=500, L=100, seed=92)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    cand = cand[ctx.ungapped(cand)[0]["cigarLength"] > 0]
    f1 = np.zeros(len(cand), dtype=capi.FRAGMENT_DTYPE)
    pool = np.zeros(len(cand) * 10, dtype=np.uint32)
    for field, value in (("readId", 1 << 30), ("contigStrand", 77 << 1), ("position", 1 << 40), ("position", -5000)):
        bad = cand.copy()
        bad[field][len(bad) // 2] = value
        with pytest.raises(capi.ExtError) as e:
            ctx.extend_compact(bad, gapped, f1, pool)
        assert e.value.code == 1, field
    words = ctx.extend_compact(cand, gapped, f1, pool)
    f0 = (ctx.gapped(cand, with_masks=False) if gapped else ctx.ungapped(cand, with_masks=False))[0]
    assert words == int(f0["cigarLength"].sum())
    assert np.array_equal(f0["mismatchCount"], f1["mismatchCount"])
    ctx.close()


def test_compact_both_passes_in_one_call(capi):
    """isaac_ext_extend_batch_compact = the two *_batch_compact calls over the same candidates, several chunks"""
    genome, sim, reads, cand = small_workload(n_pairs=3000, L=100, seed=93, indel_rate=6e-3)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    cand = cand[ctx.ungapped(cand)[0]["cigarLength"] > 0]
    cand = np.concatenate([cand] * 180)[:3 * 1212416 + 4321]
    n = len(cand)
    fu, fg = np.zeros(n, dtype=capi.FRAGMENT_DTYPE), np.zeros(n, dtype=capi.FRAGMENT_DTYPE)
    pu, pg = np.zeros(n * 3, dtype=np.uint32), np.zeros(n * 10, dtype=np.uint32)
    wu, wg = ctx.extend_compact_both(cand, fu, pu, fg, pg)
    su, sg = np.zeros(n, dtype=capi.FRAGMENT_DTYPE), np.zeros(n, dtype=capi.FRAGMENT_DTYPE)
    qu, qg = np.zeros(n * 3, dtype=np.uint32), np.zeros(n * 10, dtype=np.uint32)
    assert wu == ctx.extend_compact(cand, False, su, qu) and wg == ctx.extend_compact(cand, True, sg, qg)
    assert fu.tobytes() == su.tobytes() and fg.tobytes() == sg.tobytes()
    assert np.array_equal(pu[:wu], qu[:wu]) and np.array_equal(pg[:wg], qg[:wg])
    assert (fg["gapCount"] > 0).sum() > 1000
    # too small a pool for one of the passes is reported with the sizes both need
    with pytest.raises(capi.ExtError) as e:
        ctx.extend_compact_both(cand, fu, pu, fg, pg[:1000])
    assert e.value.code == 5
    ctx.close()


@pytest.mark.parametrize("L,seed", [(100, 95), (150, 96)])
def test_packed_alignments_are_what_align_fragments_keeps(capi, L, seed):
    """isaac_ext_align_batch_packed: one 32-byte record per candidate = the ungapped alignment, or the gapped one where
    FragmentBuilder::alignFragments would have run the gapped aligner and accepts it (FragmentBuilder.cpp:190-209); expected
    records from BOTH checkers' ungapped / gapped results; the implied CIGAR of a kept ungapped alignment is its real one;
    several chunks, masked read ends and candidates across both contig ends included"""
    from isaac_aligner_b200.batch import expected_alignments, implied_ungapped_cigar
    from isaac_aligner_b200.types import ALIGNMENT_DTYPE, ALIGNMENT_GAPPED
    genome, sim, reads, cand = small_workload(n_pairs=2500, L=L, seed=seed, indel_rate=8e-3)
    cfg = Config.default(max_read_length=2 * L)
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    n = len(cand)
    got, pool = np.zeros(n, dtype=ALIGNMENT_DTYPE), np.zeros(n * 8, dtype=np.uint32)
    words = ctx.align_packed(cand, got, pool)
    g = oracle_lib.GenomeHolder(genome)
    for chk in checkers():
        fu, cu, _ = chk.ungapped(g, reads, cfg, cand)
        fg, cg, _ = chk.gapped(g, reads, cfg, cand, cigar_stride=32)
        want, want_pool = expected_alignments(fu, cu, fg, cg, 32, L)
        for name in ALIGNMENT_DTYPE.names:
            x, y = got[name], want[name]
            if name == "logProbability":
                x, y = x.view(np.uint64), y.view(np.uint64)
            assert np.array_equal(x, y), (chk.kind, name)
        assert words == len(want_pool) and np.array_equal(pool[:words], want_pool), chk.kind
        implied = np.nonzero((fu["cigarLength"] > 0) & (got["cigarLength"] == 0))[0]
        assert len(implied) > n // 4
        for i in implied[:: max(1, len(implied) // 3000)]:
            k = int(fu["cigarLength"][i])
            assert implied_ungapped_cigar(got[i], bool(fu["reverse"][i]), L) == [int(w) for w in cu.reshape(-1, 3)[i][:k]], i
        # a kept ungapped alignment with explicit words: soft clips at a contig end (common.small_workload places candidates there)
        assert ((got["cigarLength"] > 0) & ((got["gapsAndFlags"] & ALIGNMENT_GAPPED) == 0)).sum() > 0
    assert ((got["gapsAndFlags"] & ALIGNMENT_GAPPED) != 0).sum() > 100
    # many chunks of the same candidates: the same records, the pool in candidate order
    big = np.concatenate([cand] * (1 + (3 * 1212416) // n))[:2 * 1212416 + 777]
    gb, pb = np.zeros(len(big), dtype=ALIGNMENT_DTYPE), np.zeros(len(big) * 4, dtype=np.uint32)
    wb = ctx.align_packed(big, gb, pb)
    assert gb[:n].tobytes() == got.tobytes() and gb[n:2 * n].tobytes() == got.tobytes()
    assert np.array_equal(pb[:words], pool[:words]) and wb == int(gb["cigarLength"].astype(np.int64).sum())
    with pytest.raises(capi.ExtError) as e:
        ctx.align_packed(big, gb, pb[:100])
    assert e.value.code == 5
    ctx.close()
