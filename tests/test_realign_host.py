"""realign_device.cuh -- the functions every thread of realignBinKernel runs -- on the CPU (tests/cpp/test_realign_host.cu) against the
reference's own build::RealignerGaps / build::GapRealigner / build::SemialignedEndsClipper (oracle_realign_bin) on synthetic bins:
every byte of the bin's data afterwards, Index::pos_ and the CIGAR of every index entry, the two gap lists."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200 import bins
from isaac_aligner_b200.batch import Tls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib():
    so = os.path.join(ROOT, "build", "libtest_realign_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-Xcompiler", "-fPIC", "-shared", "-Wno-deprecated-gpu-targets", "-I",
                           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_realign_host.cu"), "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def ref():
    r = oracle_lib.reference()
    if r is None:
        pytest.skip("oracle/_ref/libisaac_ref.so not built (needs /root/reference)")
    return r


def make_contigs(seed, lengths=(3000, 60000), n_runs=True):
    rng = np.random.default_rng(seed)
    contigs = []
    for n in lengths:
        c = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()
        if n_runs and n > 10000:
            for _ in range(3):
                s = int(rng.integers(1000, n - 200)); c[s:s + int(rng.integers(1, 60))] = ord("N")
        contigs.append(c)
    return contigs


def host_realign(lib, contigs, bin_, options):
    flat = np.concatenate(contigs)
    begin = np.zeros(len(contigs) + 1, dtype=np.uint64)
    begin[1:] = np.cumsum([c.size for c in contigs])
    head = (ctypes.c_uint32(len(contigs)), ctypes.c_void_p(flat.ctypes.data), ctypes.c_void_p(begin.ctypes.data))
    return oracle_lib._realign_call(lib.realign_bin_host, head, bin_, options)


def compare(bin_, got, want):
    assert np.array_equal(got.gaps, want.gaps)
    assert np.array_equal(got.deletions, want.deletions)
    bad = np.flatnonzero(got.position != want.position)
    assert bad.size == 0, (bad[:5], [bins.position_of(got.position[i]) for i in bad[:5]], [bins.position_of(want.position[i]) for i in bad[:5]])
    assert np.array_equal(got.cigar_length, want.cigar_length)
    assert np.array_equal(got.cigar_offset == bins.OWN_CIGAR, want.cigar_offset == bins.OWN_CIGAR)
    for i in np.flatnonzero(want.cigar_offset != bins.OWN_CIGAR):
        assert np.array_equal(got.cigar(i, bin_), want.cigar(i, bin_)), (i, bins.cigar_string(got.cigar(i, bin_)), bins.cigar_string(want.cigar(i, bin_)))
    if not np.array_equal(got.data, want.data):
        for o in bin_.record_offset:
            g, w = bin_.header(int(o), got.data), bin_.header(int(o), want.data)
            assert g == w, (int(o), g, w)
    return int(np.count_nonzero(want.cigar_offset != bins.OWN_CIGAR))


@pytest.mark.parametrize("seed,vigorous,clip,dodgy", [(1, False, False, False), (2, False, True, False), (3, True, True, True),
                                                        (4, True, False, False), (5, False, True, True)])
def test_realigner_restatement_equals_reference(host_lib, ref, seed, vigorous, clip, dodgy):
    contigs = make_contigs(seed)
    genome = oracle_lib.GenomeHolder(contigs)
    bin_ = bins.simulate_bin(contigs, contig=1, region=(2000, 30000), n_pairs=2500, read_length=100, seed=seed, barcodes=2)
    options = bins.RealignOptions(bin_.bin_start, bin_.bin_end, [Tls.make(), Tls.make()], vigorous=vigorous, dodgy=dodgy,
                                  clip_semialigned=clip, gap_groups=[0, 1] if seed % 2 else None)
    want = oracle_lib.realign_bin(ref, genome, bin_, options)
    got, counts = host_realign(host_lib, contigs, bin_, options)
    assert int(counts[4]) == 0, "error flags %d" % int(counts[4])
    realigned = compare(bin_, got, want)
    assert realigned > 50, realigned                       # the bins are built so that the realigner has work
    assert int(counts[3]) == realigned
