"""The end clippers of isaac_ext_build_templates on the CPU: clipTemplateEndsOfCluster (csrc/kernels_clip.cuh, the body of
clipTemplateEndsKernel: SemialignedEndsClipper.cpp:32-205 + OverlappingEndsClipper.cpp:45-180) run by tests/cpp/test_clip_host.cu
over host-packed reference and reads.

* the two literal blocks of the reference's testOverlappingEndsClipper.cpp (:113-153: 4-base reads, hand-made templates; the one
  unit test of the path that cannot go through the seeded GPU calls);
* whole tiles: the CPU-run template worker (tests/test_template_worker.py) followed by the CPU-run clippers must give what the
  reference's TemplateBuilder followed by its own clippers gives (oracle_build_templates with the clip flags), bit for bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from common_build import build_workload
from isaac_aligner_b200.batch import CLIP_OVERLAPPING, CLIP_SEMIALIGNED, TEMPLATE_DTYPE, Templates, Tls, TemplateOptions
from isaac_aligner_b200.types import FRAGMENT_DTYPE, Config, ReadSet, cigar_to_string
from test_gpu_templates import assert_templates_equal
from test_template_worker import worker_lib, worker_templates      # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clip_lib():
    so = os.path.join(ROOT, "build", "libtest_clip_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared",
                           "-shared", "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "cpp", "test_clip_host.cu"), "-o", so])
    return ctypes.CDLL(so)


def clip(lib, genome, reads, flags, templates):
    """the product's clippers on the CPU -> batch.Templates"""
    bases = np.concatenate([np.asarray(c, dtype=np.uint8) for c in genome])
    begin = np.zeros(len(genome) + 1, dtype=np.uint64)
    begin[1:] = np.cumsum([len(c) for c in genome])
    fragments = templates.fragments.copy()
    cigars_in = np.ascontiguousarray(templates.cigars, dtype=np.uint32)
    cigars_out = np.zeros(cigars_in.size + 4 * len(fragments) + 4, dtype=np.uint32)
    t = np.ascontiguousarray(templates.templates)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert lib.clip_templates_host(ctypes.c_uint32(len(genome)), p(bases), p(begin), ctypes.byref(reads.c), ctypes.c_uint32(flags), p(t),
                                   p(fragments), p(cigars_in), p(cigars_out)) == 0
    return Templates(t, fragments, cigars_out, templates.rescue_requests)


def test_overlapping_ends_clipper_literals(clip_lib):
    """testOverlappingEndsClipper.cpp:113-153; its ReadInit (:57-68) keeps the given quality string for a reverse read and reverses it
    for a forward one, and reverses (never complements) the bases of a reverse read: ACGT is its own reverse complement"""
    for q1, q2, want in (("CFCE", "BDBE", (("4M", 0), ("3S1M", 4))), ("BAAA", "CFCE", (("1M3S", 0), ("4M", 1)))):
        forward_q1, forward_q2 = q1[::-1], q2
        bcl = np.array([[((ord(q) - 33) << 2) | "ACGT".index(b) for b, q in zip("ACGT" + "ACGT", forward_q1 + forward_q2)]], dtype=np.uint8)
        reads = ReadSet(bcl, (4, 4))
        genome = [np.frombuffer(b"ACGT", dtype=np.uint8)]
        t = np.zeros(1, dtype=TEMPLATE_DTYPE)
        t["built"] = 1
        f = np.zeros(2, dtype=FRAGMENT_DTYPE)
        f["readId"], f["readIndex"], f["observedLength"], f["cigarLength"] = (0, 1), (0, 1), 4, 1
        f["position"], f["reverse"], f["cigarOffset"] = (0, 1), (0, 1), (0, 1)                  # read 2 " ACGT": position 1, reverse
        got = clip(clip_lib, genome, reads, CLIP_OVERLAPPING, Templates(t, f, np.array([4 << 4, 4 << 4], dtype=np.uint32)))
        for i, (cigar, position) in enumerate(want):
            assert (cigar_to_string(got.cigar(i)), int(got.fragments["position"][i])) == (cigar, position), (q1, q2, i)


@pytest.mark.skipif(not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"),
                    reason="the clippers' checker is the reference build only")
@pytest.mark.parametrize("L,seed,flags,kw", [(100, 61, CLIP_SEMIALIGNED, {}),
                                             (100, 62, CLIP_OVERLAPPING, {"insert": (160.0, 30.0, 100, 260)}),      # mates overlap
                                             (75, 63, CLIP_SEMIALIGNED | CLIP_OVERLAPPING, {"indel_rate": 1e-2, "insert": (120.0, 25.0, 80, 200)}),
                                             (150, 64, CLIP_SEMIALIGNED | CLIP_OVERLAPPING, {"insert": (250.0, 40.0, 150, 400)})])
def test_worker_and_clippers_give_the_references_clipped_templates(worker_lib, clip_lib, L, seed, flags, kw):
    ref = oracle_lib.reference()
    genome, sim, reads, mb = build_workload(n_pairs=1000, L=L, seed=seed, **kw)
    config = Config.default(max_read_length=2 * L)
    tls = Tls.make()
    unclipped, g = worker_templates(worker_lib, ref, genome, reads, config, mb, tls, TemplateOptions.make())
    got = clip(clip_lib, genome, reads, flags, unclipped)
    options = TemplateOptions.make(clip_semialigned=bool(flags & CLIP_SEMIALIGNED), clip_overlapping=bool(flags & CLIP_OVERLAPPING))
    want = oracle_lib.build_templates(ref, g, reads, config, mb, tls, options, threads=4)
    assert_templates_equal(got, want, "CPU worker + CPU clippers, flags %d" % flags)
    changed = 0
    for i in np.nonzero(want.fragments["cigarLength"])[0]:
        assert np.array_equal(got.cigar(i), want.cigar(i)), i
        changed += not np.array_equal(got.cigar(i), unclipped.cigar(i))
    assert changed > 0                                                   # the clippers did something on this tile
