"""The host half of isaac_ext_build_templates on the CPU: csrc/template_worker.cuh (pair selection, rescue decisions, mapping
scores: TemplateBuilder.cpp:97-1089 restated) driven by tests/cpp/test_template_worker.cu the way isaac_ext_templates.cuh drives
it -- plan, ONE batch of rescueShadow calls, finish -- with the checker standing in for the two GPU batch calls
(oracle_build_fragments / oracle_rescue_shadows of the reference build).  The templates must be the reference's own
TemplateBuilder's (oracle_build_templates), bit for bit.  No GPU involved; the GPU tests (test_gpu_templates.py) check the same
code behind the real kernels."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from common_build import build_workload
from isaac_aligner_b200.batch import (DODGY_ALIGNMENT_SCORE_UNALIGNED, DODGY_ALIGNMENT_SCORE_UNKNOWN, RESCUE_REQUEST_DTYPE, TEMPLATE_DTYPE,
                                      BuildResult, RescueResult, Templates, Tls, TemplateOptions)
from isaac_aligner_b200.types import FRAGMENT_DTYPE, Config
from test_gpu_templates import assert_templates_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"),
                                reason="the TemplateBuilder checker is the reference build only")


@pytest.fixture(scope="module")
def worker_lib():
    so = os.path.join(ROOT, "build", "libtest_template_worker.so")
    src = os.path.join(ROOT, "tests", "cpp", "test_template_worker.cu")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-cudart", "shared", "-shared", "-Xcompiler", "-fPIC", src, "-o", so])
    return ctypes.CDLL(so)


def p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size else None


def flat_view(flat, cls):
    """FlatFragments (the checker's result) as the C struct the worker reads"""
    return cls(flat.fragments.ctypes.data if flat.fragments.size else None, flat.begin.ctypes.data,
               flat.cigars.ctypes.data if flat.cigars.size else None, flat.flags.ctypes.data, flat.fragments.size, flat.cigars.size)


def worker_templates(lib, ref, genome, reads, config, mb, tls, options, threads=3, device_finish=True):
    g = oracle_lib.GenomeHolder(genome)
    n, rc = reads.cluster_count, reads.read_count
    built = oracle_lib.build_fragments(ref, g, reads, config, mb, threads=4)
    built_c = flat_view(built, BuildResult)
    read_length = np.array(list(reads.read_lengths) + [0] * (2 - rc), dtype=np.uint32)
    contig_length = np.array([len(c) for c in genome], dtype=np.uint64)
    requests = np.zeros(4 * n + 16, dtype=RESCUE_REQUEST_DTYPE)
    request_begin = np.zeros(n + 1, dtype=np.uint64)
    head = [ctypes.c_uint32(n), ctypes.c_uint32(rc), p(read_length), ctypes.c_uint32(len(contig_length)), p(contig_length),
            ctypes.byref(tls), ctypes.byref(options), ctypes.byref(built_c)]
    assert lib.template_worker_plan(*head, ctypes.c_uint64(requests.size), p(requests), p(request_begin), ctypes.c_uint(threads)) == 0
    requests = requests[:int(request_begin[-1])].copy()
    rescued = oracle_lib.rescue_shadows(ref, g, reads, config, tls, requests, threads=4, fragments_per_request=128)
    rescued_c = flat_view(rescued, RescueResult)
    templates = np.zeros(n, dtype=TEMPLATE_DTYPE)
    fragments = np.zeros(n * rc, dtype=FRAGMENT_DTYPE)
    cigars = np.zeros(64 * n + 1024, dtype=np.uint32)
    words = ctypes.c_uint64()
    assert lib.template_worker_finish(*head, ctypes.byref(rescued_c), p(request_begin), p(templates), p(fragments),
                                      ctypes.c_uint64(cigars.size), p(cigars), ctypes.byref(words), ctypes.c_uint(threads)) == 0
    if device_finish:
        # the same inputs through finish_device.cuh (the body of finishTemplatesKernel): byte-identical records wanted
        t2, f2, c2, w2 = np.zeros_like(templates), np.zeros_like(fragments), np.zeros_like(cigars), ctypes.c_uint64()
        assert lib.finish_device_templates(*head, ctypes.byref(rescued_c), p(request_begin), p(t2), p(f2), ctypes.c_uint64(c2.size), p(c2),
                                           ctypes.byref(w2)) == 0
        assert t2.tobytes() == templates.tobytes(), "finish_device.cuh: templates differ from template_worker.cuh"
        for name in FRAGMENT_DTYPE.names:
            assert np.array_equal(f2[name], fragments[name]), "finish_device.cuh: fragment field %s differs" % name
        assert w2.value == words.value and np.array_equal(c2[:w2.value], cigars[:words.value])
    return Templates(templates, fragments, cigars[:words.value].copy(), len(requests)), g


@pytest.mark.parametrize("L,seed,options", [
    (100, 41, dict()),
    (150, 42, dict(mapq_threshold=10)),
    (75, 43, dict(scatter_repeats=True, dodgy=DODGY_ALIGNMENT_SCORE_UNKNOWN)),
    (100, 44, dict(dodgy=DODGY_ALIGNMENT_SCORE_UNALIGNED, mapq_threshold=3)),
])
def test_plan_and_finish_give_the_references_templates(worker_lib, L, seed, options):
    ref = oracle_lib.reference()
    genome, sim, reads, mb = build_workload(n_pairs=1200, L=L, seed=seed)
    config = Config.default(max_read_length=2 * L)
    tls, opt = Tls.make(), TemplateOptions.make(**options)
    got, g = worker_templates(worker_lib, ref, genome, reads, config, mb, tls, opt)
    want = oracle_lib.build_templates(ref, g, reads, config, mb, tls, opt, threads=4)
    assert_templates_equal(got, want, "template worker on the CPU, L %d" % L)
    for i in np.nonzero(want.fragments["cigarLength"])[0]:
        assert np.array_equal(got.cigar(i), want.cigar(i)), i
    assert got.rescue_requests > 0 and int(got.templates["built"].sum()) > 0


def test_thread_count_changes_nothing(worker_lib):
    ref = oracle_lib.reference()
    genome, sim, reads, mb = build_workload(n_pairs=700, L=100, seed=47)
    config = Config.default(max_read_length=200)
    tls, opt = Tls.make(), TemplateOptions.make()
    a, _ = worker_templates(worker_lib, ref, genome, reads, config, mb, tls, opt, threads=1)
    b, _ = worker_templates(worker_lib, ref, genome, reads, config, mb, tls, opt, threads=5)
    assert_templates_equal(a, b, "1 thread against 5")


def test_single_ended_templates(worker_lib):
    """single-ended data: pickBestFragment (TemplateBuilder.cpp:1035-1058), no rescueShadow call at all"""
    from test_gpu_templates import single_ended
    ref = oracle_lib.reference()
    genome, sim, reads, mb = build_workload(n_pairs=800, L=100, seed=51)
    reads1, mb1 = single_ended(genome, sim, reads, mb)
    config = Config.default(max_read_length=200)
    tls, opt = Tls.make(), TemplateOptions.make(mapq_threshold=4)
    got, g = worker_templates(worker_lib, ref, genome, reads1, config, mb1, tls, opt)
    want = oracle_lib.build_templates(ref, g, reads1, config, mb1, tls, opt, threads=4)
    for name in ("hadFragments", "built", "properPair", "alignmentScore"):
        assert np.array_equal(got.templates[name], want.templates[name]), name
    assert np.array_equal(got.templates["fragmentAlignmentScore"][:, 0], want.templates["fragmentAlignmentScore"][:, 0])
    for name in ("position", "contigId", "observedLength", "editDistance", "cigarLength", "reverse", "mismatchCount"):
        assert np.array_equal(got.fragments[name], want.fragments[name]), name
    assert got.rescue_requests == 0


@pytest.mark.parametrize("L,seed,scatter,kw", [(100, 71, False, {}), (100, 72, True, {"repeat_rate": 0.2, "neighbor_rate": 0.4}),
                                               (150, 73, True, {"indel_rate": 1e-2}), (75, 74, False, {"neighbor_rate": 0.6})])
def test_device_plan_pass_records_the_same_requests(worker_lib, L, seed, scatter, kw):
    """csrc/plan_device.cuh (the plan pass without libm and without std::vector, for a one-thread-per-cluster kernel) against the
    planning mode of template_worker.cuh: the same rescueShadow calls in the same order, byte for byte"""
    ref = oracle_lib.reference()
    genome, sim, reads, mb = build_workload(n_pairs=1500, L=L, seed=seed, **kw)
    config = Config.default(max_read_length=2 * L)
    g = oracle_lib.GenomeHolder(genome)
    built = oracle_lib.build_fragments(ref, g, reads, config, mb, threads=4)
    built_c = flat_view(built, BuildResult)
    n, rc = reads.cluster_count, reads.read_count
    read_length = np.array(list(reads.read_lengths), dtype=np.uint32)
    contig_length = np.array([len(c) for c in genome], dtype=np.uint64)
    for tls in (Tls.make(), Tls.make(mn=100, mx=300, median=200, low=20, high=20), Tls.make(mn=0xFFFFFFFF, mx=0xFFFFFFFF, median=0xFFFFFFFF, m0=8, m1=8)):
        options = TemplateOptions.make(scatter_repeats=scatter)
        want, want_begin = np.zeros(4 * n + 16, dtype=RESCUE_REQUEST_DTYPE), np.zeros(n + 1, dtype=np.uint64)
        assert worker_lib.template_worker_plan(ctypes.c_uint32(n), ctypes.c_uint32(rc), p(read_length), ctypes.c_uint32(len(contig_length)),
                                               p(contig_length), ctypes.byref(tls), ctypes.byref(options), ctypes.byref(built_c),
                                               ctypes.c_uint64(want.size), p(want), p(want_begin), ctypes.c_uint(3)) == 0
        got, got_begin = np.zeros(4 * n + 16, dtype=RESCUE_REQUEST_DTYPE), np.zeros(n + 1, dtype=np.uint64)
        assert worker_lib.plan_device_requests(ctypes.c_uint32(n), ctypes.c_uint32(rc), ctypes.byref(tls), ctypes.byref(options),
                                               ctypes.byref(built_c), ctypes.c_uint64(got.size), p(got), p(got_begin)) == 0
        assert np.array_equal(got_begin, want_begin)
        total = int(want_begin[-1])
        assert got[:total].tobytes() == want[:total].tobytes()
    assert total >= 0 and int(want_begin[-1]) == total


def test_plan_device_compiles_for_the_device():
    """plan_device.cuh through nvcc for sm_100a inside a one-thread-per-cluster kernel (device code generation only, no GPU needed)"""
    src = os.path.join(ROOT, "build", "plan_device_kernel.cu")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    with open(src, "w") as f:
        f.write('#include "../isaac_aligner_b200/csrc/plan_device.cuh"\n'
                '__global__ void planKernel(const isaac_b200::PlanView v, unsigned clusters, const unsigned long long *begin, isaac_ext_rescue_request_t *out, unsigned *counts)\n'
                '{\n'
                '    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;\n'
                '    if (c < clusters) counts[c] = isaac_b200::planClusterRequests(v, c, out + begin[c], unsigned(begin[c + 1] - begin[c]));\n'
                '}\n')
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--extended-lambda",
                           "-Xptxas", "-v", "-c", src, "-o", os.path.join(ROOT, "build", "plan_device_kernel.o")])


def test_device_plan_pass_scatter_repeats_on_hand_made_ties(worker_lib):
    """equally good pairs of different template lengths (a read placed on several copies of a tandem repeat): --scatter-repeats picks
    the pair by cluster id, which changes the best template length the requests carry; equally good orphans likewise"""
    from isaac_aligner_b200.batch import FlatFragments
    from test_template_worker_goldens import fragment
    n = 12
    frags, begin, rng = [], [0], np.random.default_rng(5)
    for c in range(n):
        copies = 1 + c % 4
        r1 = [fragment(0, 1000, 100, 0, 0, 0, 1, -9.0, 2)]
        r2 = [fragment(0, 1200 + 37 * k, 100, 1, 1, 0, 1, -11.0 + (1e-9 if k % 2 else 0.0), 1) for k in range(copies)]
        if c % 5 == 4:
            r1.append(fragment(0, 1003, 100, 0, 0, 0, 1, -9.0, 0))     # a second, equally good placement of read 1
        if c == 7:
            r2 = []                                                      # orphan only: TemplateBuilder::rescueShadow
        for f in r1 + r2:
            f["readId"], f["editDistance"], f["smithWatermanScore"] = c * 2 + int(f["readIndex"]), 1 + int(rng.integers(0, 2)), 3
        frags += r1 + r2
        begin += [begin[-1] + len(r1), begin[-1] + len(r1) + len(r2)]
    built = FlatFragments(np.array(frags, dtype=FRAGMENT_DTYPE), np.array(begin, dtype=np.uint64), np.full(8, 1600, dtype=np.uint32), np.ones(n, dtype=np.uint8))
    built_c = flat_view(built, BuildResult)
    read_length, contig_length = np.array([100, 100], dtype=np.uint32), np.array([100000], dtype=np.uint64)
    lists = []
    for scatter in (False, True):
        tls, options = Tls.make(), TemplateOptions.make(scatter_repeats=scatter)
        want, want_begin = np.zeros(256, dtype=RESCUE_REQUEST_DTYPE), np.zeros(n + 1, dtype=np.uint64)
        assert worker_lib.template_worker_plan(ctypes.c_uint32(n), ctypes.c_uint32(2), p(read_length), ctypes.c_uint32(1), p(contig_length),
                                               ctypes.byref(tls), ctypes.byref(options), ctypes.byref(built_c), ctypes.c_uint64(want.size),
                                               p(want), p(want_begin), ctypes.c_uint(1)) == 0
        got, got_begin = np.zeros(256, dtype=RESCUE_REQUEST_DTYPE), np.zeros(n + 1, dtype=np.uint64)
        assert worker_lib.plan_device_requests(ctypes.c_uint32(n), ctypes.c_uint32(2), ctypes.byref(tls), ctypes.byref(options), ctypes.byref(built_c),
                                               ctypes.c_uint64(got.size), p(got), p(got_begin)) == 0
        total = int(want_begin[-1])
        assert total > n and np.array_equal(got_begin, want_begin) and got[:total].tobytes() == want[:total].tobytes(), scatter
        lists.append(want[:total].copy())
    assert lists[0].tobytes() != lists[1].tobytes()                      # the option does change the requests of this tile


def test_device_shadow_windows_against_the_references_rescue_range(worker_lib):
    """csrc/shadow_window_device.cuh (R1 of the rescue pass for a one-thread-per-request kernel) against the reference's own
    calculateShadowRescueRange and TemplateLengthStatistics::mateOrientation (oracle_shadow_rescue_range), with the clamps of
    rescueShadow (ShadowAligner.cpp:179-197) restated here: orphans of both reads and strands, with and without a best template
    length, near the contig ends, every coherent pair of models and mate drift ranges"""
    ref = oracle_lib.reference()
    rng = np.random.default_rng(81)
    n, L0, L1 = 4000, 100, 76
    contig_length = np.array([5000, 300, 120000], dtype=np.uint64)
    bcl = (rng.integers(2, 42, size=(50, L0 + L1)).astype(np.uint8) << 2) | rng.integers(0, 4, size=(50, L0 + L1)).astype(np.uint8)
    from isaac_aligner_b200.types import ReadSet
    reads = ReadSet(bcl, (L0, L1))
    read_length = np.array([L0, L1], dtype=np.uint32)
    req = np.zeros(n, dtype=RESCUE_REQUEST_DTYPE)
    contig = rng.integers(0, 3, size=n)
    req["orphanReadId"] = rng.integers(0, 100, size=n)
    req["orphanContigStrand"] = (contig << 1) | rng.integers(0, 2, size=n)
    req["orphanPosition"] = np.where(rng.random(n) < 0.2, rng.integers(0, 30, size=n), (rng.random(n) * (contig_length[contig] - 1)).astype(np.int64))
    req["orphanObservedLength"] = np.where(rng.random(n) < 0.1, 0, rng.integers(60, 140, size=n))
    req["bestTemplateLength"] = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 3000, size=n))
    coherent = [(1, 6), (6, 1), (2, 5), (5, 2), (0, 7), (3, 4), (1, 2)]          # FRp/RFm, RFp/FRm, FFp/RRm, RRp/FFm, and one incoherent pair
    for m0, m1 in coherent:
        for drift in (-1, 0, 40):
            tls = Tls.make(mn=int(rng.integers(50, 250)), mx=int(rng.integers(300, 900)), median=280, low=30, high=30, m0=m0, m1=m1, drift=drift)
            tasks, got_range = np.zeros((n, 4), dtype=np.int64), np.zeros((n, 2), dtype=np.int64)
            rc_ = worker_lib.shadow_windows_device(p(read_length), ctypes.byref(tls), ctypes.c_uint32(n), p(req), p(contig_length), p(tasks), p(got_range))
            assert rc_ == (1 if (m0, m1) == (1, 2) else 0)                       # TemplateLengthStatistics::isCoherent
            want_range, orientation = np.zeros((n, 2), dtype=np.int64), np.zeros(n, dtype=np.uint8)
            assert ref.lib.oracle_shadow_rescue_range(ctypes.byref(reads.c), ctypes.byref(tls), ctypes.c_uint32(n), p(req), p(want_range), p(orientation)) == 0
            assert np.array_equal(got_range, want_range), (m0, m1, drift)
            first, second = want_range[:, 0], want_range[:, 1]
            shadow_index = (req["orphanReadId"] + 1) % 2
            begin = np.maximum(0, first)
            end = np.minimum(contig_length[contig].astype(np.int64), second + 1)
            empty = (second < first) | (second + 1 + read_length[shadow_index].astype(np.int64) < 0)
            end = np.where(empty, begin, end)
            assert np.array_equal(tasks[:, 0], begin) and np.array_equal(tasks[:, 1], end), (m0, m1, drift)
            assert np.array_equal(tasks[:, 2], req["orphanReadId"] - req["orphanReadId"] % 2 + shadow_index)
            assert np.array_equal(tasks[:, 3], (contig << 1) | orientation), (m0, m1, drift)


def test_shadow_window_device_compiles_for_the_device():
    src = os.path.join(ROOT, "build", "shadow_window_kernel.cu")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    with open(src, "w") as f:
        f.write('#include "../isaac_aligner_b200/csrc/shadow_window_device.cuh"\n'
                'struct Task { long long windowBegin, windowEnd; unsigned shadowReadId, contigStrand; };\n'
                '__global__ void windowKernel(const isaac_b200::ShadowWindowModel m, unsigned n, const isaac_ext_rescue_request_t *q, const unsigned *len,\n'
                '                             const unsigned long long *contigLength, Task *tasks)\n'
                '{\n'
                '    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;\n'
                '    if (i < n) isaac_b200::shadowWindowOf(m, q[i], len, long(contigLength[q[i].orphanContigStrand >> 1]), tasks[i]);\n'
                '}\n')
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-c", src, "-o",
                           os.path.join(ROOT, "build", "shadow_window_kernel.o")])
