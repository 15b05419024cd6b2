"""BASELINE.json configs[1] at its full size (10 M candidates, 150 bp) on the CUDA path, checked through properties that
do not need the oracle to run 10 M alignments: structural invariants of every record, sub-batch invariance across the
chunk boundaries of the split Smith-Waterman, linearity of the tile statistics, idempotence -- plus bit-exact parity of a
random sample against both CPU checkers."""
import os

import numpy as np
import pytest

import oracle_lib
from common import assert_fragments_equal
from isaac_aligner_b200 import synth
from isaac_aligner_b200.types import CANDIDATE_DTYPE, FRAGMENT_DTYPE, Config, ReadSet

pytestmark = pytest.mark.gpu

STRIDE = 32
SIZES = [(10_000_000, 150), (4_000_000, 250)]        # BASELINE configs[1]; the read length of configs[4] at a size that keeps the suite short


def checkers():
    return oracle_lib.gpu_checkers()        # fails when the reference build did not travel


@pytest.fixture(scope="module", params=SIZES, ids=["10M_x_150bp", "4M_x_250bp"])
def full(request):
    """the workload of bench.py (same generator, same seeds) and the results of one full-size pass"""
    import torch
    from isaac_aligner_b200 import capi
    N, L = request.param
    genome = synth.make_genome(5_000_000, n_contigs=1, seed=synth.SEED_G5)
    sim = synth.simulate_pairs(genome, -(-N // 16), L=L, seed=synth.SEED_READS + 1)
    reads = ReadSet(sim.bcl, (L, L))
    cand = synth.microbench_candidates(sim, genome, per_read=8, seed=synth.SEED_READS + 2)[:N]
    assert len(cand) == N
    ctx = capi.Context(Config.default(max_read_length=2 * L))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    dev = torch.device("cuda", 0)
    d_cand = torch.from_numpy(cand.view(np.uint8).reshape(N, 16)).to(dev)
    stream = torch.cuda.current_stream().cuda_stream

    def run(gapped, first=0, count=N):
        frag = torch.empty((count, 64), dtype=torch.uint8, device=dev)
        cig = torch.zeros((count, STRIDE if gapped else 3), dtype=torch.int32, device=dev)
        ptr = d_cand.data_ptr() + 16 * first
        if gapped:
            ctx.gapped_device(count, ptr, STRIDE, frag.data_ptr(), cig.data_ptr(), 0, stream)
        else:
            ctx.ungapped_device(count, ptr, frag.data_ptr(), cig.data_ptr(), 0, stream)
        torch.cuda.synchronize()
        return frag, cig

    state = {"ctx": ctx, "genome": genome, "reads": reads, "cand": cand, "run": run, "torch": torch, "dev": dev, "N": N, "L": L}
    state["ungapped"] = run(False)
    state["gapped"] = run(True)
    yield state
    ctx.close()


def host(frag, cig):
    return frag.cpu().numpy().reshape(-1).view(FRAGMENT_DTYPE), cig.cpu().numpy().view(np.uint32)


def cigar_lengths(cig, lengths):
    """bases of the read and of the reference covered by each CIGAR row (ops M0 I1 D2 S4)"""
    k = np.arange(cig.shape[1])[None, :] < lengths[:, None]
    op, ln = cig & 0xF, (cig >> 4).astype(np.int64)
    read = (ln * (k & ((op == 0) | (op == 1) | (op == 4)))).sum(axis=1)
    ref = (ln * (k & ((op == 0) | (op == 2)))).sum(axis=1)
    gaps = (k & ((op == 1) | (op == 2))).sum(axis=1)
    return read, ref, gaps


@pytest.mark.parametrize("which", ["ungapped", "gapped"])
def test_every_record_is_well_formed(full, which):
    N, L = full["N"], full["L"]
    f, c = host(*full[which])
    aligned = f["cigarLength"] > 0
    assert aligned.mean() > 0.7
    for lo in range(0, N, 2_000_000):                      # bounded numpy temporaries
        s = slice(lo, lo + 2_000_000)
        a = aligned[s]
        read, ref, gaps = cigar_lengths(c[s][a], f["cigarLength"][s][a].astype(np.int64))
        assert np.array_equal(read, np.full(read.shape, L)), "a CIGAR does not cover the read"
        assert np.array_equal(ref, f["observedLength"][s][a].astype(np.int64)), "observedLength differs from the CIGAR"
        assert np.array_equal(gaps, f["gapCount"][s][a].astype(np.int64))
        assert (f["editDistance"][s][a] >= f["mismatchCount"][s][a]).all()
        assert (f["matchCount"][s][a].astype(np.int64) + f["mismatchCount"][s][a] <= L).all()
        assert (f["matchesInARow"][s][a] <= f["matchCount"][s][a]).all()
        assert (f["logProbability"][s][a] < 0).all()
    assert np.array_equal(f["readId"], full["cand"]["readId"])
    if which == "gapped":
        assert (f["gapCount"] > 0).sum() > N // 10           # the shifted 20 % need a gap


def test_sampled_parity_with_the_cpu_checkers(full):
    """30 000 random candidates of the batch: the full-size results equal the oracle's, bit for bit"""
    N = full["N"]
    fu, cu = host(*full["ungapped"])
    fg, cg = host(*full["gapped"])
    rng = np.random.default_rng(2024)
    pick = np.sort(rng.choice(N, 30_000, replace=False))
    # GappedAligner is only ever handed fragments whose ungapped alignment kept a match (FragmentBuilder.cpp:179)
    pick_g = pick[fu["cigarLength"][pick] > 0]
    g = oracle_lib.GenomeHolder(full["genome"])
    cand = full["cand"]
    for chk in checkers():
        fr, cr, _ = chk.ungapped(g, full["reads"], full["ctx"].config, cand[pick], threads=8)
        a = fu[pick].copy(); a["cigarOffset"] = fr["cigarOffset"]
        assert_fragments_equal(a, fr, cu[pick], cr, None, None, "full-size ungapped vs " + chk.kind)
        fr, cr, _ = chk.gapped(g, full["reads"], full["ctx"].config, cand[pick_g], cigar_stride=STRIDE, threads=8)
        a = fg[pick_g].copy(); a["cigarOffset"] = fr["cigarOffset"]
        assert_fragments_equal(a, fr, cg[pick_g], cr, None, None, "full-size gapped vs " + chk.kind)


def test_sub_batches_equal_the_full_batch(full):
    """slices that start and end off the chunk / pair boundaries of the split kernels give the same records"""
    chunk = 2 * 148 * 4 * 128 * 2
    N = full["N"]
    fg, cg = host(*full["gapped"])
    for first, count in [(0, chunk - 1), (1, chunk), (chunk - 3, 2 * chunk + 5), (N - 777, 777), (N // 2 + 1, 1)]:
        f, c = host(*full["run"](True, first, count))
        ref = fg[first:first + count].copy()
        ref["cigarOffset"] -= first * STRIDE
        assert f.tobytes() == ref.tobytes(), (first, count)
        assert np.array_equal(c, cg[first:first + count]), (first, count)


def test_idempotent(full):
    for which, gapped in (("ungapped", False), ("gapped", True)):
        f, c = full["run"](gapped)
        assert full["torch"].equal(f, full[which][0]) and full["torch"].equal(c, full[which][1]), which


def test_tile_statistics_are_linear(full):
    """K6 counters of the whole batch = sum over four tiles = the same sums taken on the host"""
    torch, ctx, N = full["torch"], full["ctx"], full["N"]
    frag = full["gapped"][0]
    stream = torch.cuda.current_stream().cuda_stream
    whole = torch.zeros(64, dtype=torch.int64, device=full["dev"])
    ctx.tile_stats_device(N, frag.data_ptr(), whole.data_ptr(), stream)
    parts = torch.zeros(64, dtype=torch.int64, device=full["dev"])
    q = N // 4
    for k in range(4):
        ctx.tile_stats_device(q, frag.data_ptr() + 64 * q * k, parts.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    f, _ = host(*full["gapped"])
    from isaac_aligner_b200 import distributed
    expect = distributed.stats_from_fragments(f)
    assert np.array_equal(whole.cpu().numpy().view(np.uint64)[:len(expect)], expect)
