"""Differential campaign of the product code that runs on the CPU harnesses (no GPU) against the reference build of the checker:
  * csrc/template_worker.cuh (plan / finish) + csrc/kernels_clip.cuh (clipTemplateEndsOfCluster) against the reference's
    TemplateBuilder + SemialignedEndsClipper + OverlappingEndsClipper on whole tiles;
  * csrc/plan_device.cuh against the planning mode of the template worker (same rescueShadow requests, byte for byte) and
    csrc/shadow_window_device.cuh on those requests against the reference's calculateShadowRescueRange;
  * csrc/pack_fragments.cuh (the warp functions of packFragmentsKernel, lane after lane) against the reference's io::FragmentHeader.
Random read lengths, insert sizes, indel / neighbour / repeat rates, score presets, template length statistics, options, lane
counts and buffer alignments.  Development aid:
  python tools/fuzz_host_code.py [rounds [campaign seed]]"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib                                                    # noqa: E402
import test_clip_host as C                                           # noqa: E402
import test_template_worker as W                                     # noqa: E402
import test_tile_write_bin_records as P                              # noqa: E402
from common_build import build_workload                              # noqa: E402
from isaac_aligner_b200.batch import (CLIP_OVERLAPPING, CLIP_SEMIALIGNED, DODGY_ALIGNMENT_SCORE_UNALIGNED, DODGY_ALIGNMENT_SCORE_UNKNOWN,  # noqa: E402
                                      RESCUE_REQUEST_DTYPE, BuildResult, Tls, TemplateOptions)
from isaac_aligner_b200.types import ELAND_SCORES, Config, ReadSet   # noqa: E402

NVCC = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared",
        "-shared", "-Xcompiler", "-fPIC"]


def build(name, source, compiler):
    so = os.path.join(ROOT, "build", name)
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(compiler + [os.path.join(ROOT, "tests", "cpp", source), "-o", so])
    return ctypes.CDLL(so)


def main(rounds, campaign=20261018):
    ref = oracle_lib.reference()
    worker = build("libtest_template_worker.so", "test_template_worker.cu", NVCC)
    clipper = build("libtest_clip_host.so", "test_clip_host.cu", NVCC)
    packer = build("libtest_pack_fragments.so", "test_pack_fragments.cpp", ["g++", "-std=c++14", "-O2", "-Wall", "-shared", "-fPIC"])
    rng = np.random.default_rng(campaign)
    for k in range(rounds):
        seed = int(rng.integers(1, 1 << 30))
        L = int(rng.choice([50, 75, 100, 125, 150]))
        mean = float(rng.choice([1.3 * L, 2.0 * L, 350.0]))
        kw = dict(n_pairs=int(rng.integers(200, 700)), L=L, seed=seed, indel_rate=float(rng.choice([5e-4, 4e-3, 1e-2])),
                  neighbor_rate=float(rng.choice([0.0, 0.15, 0.5])), repeat_rate=float(rng.choice([0.0, 0.01, 0.1])),
                  insert=(mean, mean / 8, int(max(L, mean / 2)), int(mean * 1.6)))
        genome, sim, reads, mb = build_workload(**kw)
        config = Config.default(max_read_length=2 * L) if rng.random() < 0.7 else Config.default(scores=ELAND_SCORES, max_read_length=2 * L)
        tls = Tls.make() if rng.random() < 0.5 else Tls.make(mn=int(mean * 0.6), mx=int(mean * 1.5), median=int(mean), low=int(mean / 10),
                                                              high=int(mean / 10), drift=int(rng.choice([-1, 0, 20])))
        scatter = bool(rng.integers(0, 2))
        dodgy = int(rng.choice([0, 10, DODGY_ALIGNMENT_SCORE_UNALIGNED, DODGY_ALIGNMENT_SCORE_UNKNOWN]))
        mapq = int(rng.choice([0, 0, 5, 30]))
        flags = int(rng.integers(0, 4))
        what = "round %d: %r, flags %d, scatter %d, dodgy %d, mapq %d" % (k, kw, flags, scatter, dodgy, mapq)
        # ---- template worker + clippers against the reference's TemplateBuilder + clippers
        plain = TemplateOptions.make(scatter_repeats=scatter, dodgy=dodgy, mapq_threshold=mapq)
        unclipped, g = W.worker_templates(worker, ref, genome, reads, config, mb, tls, plain, threads=int(rng.integers(1, 6)))
        got = C.clip(clipper, genome, reads, flags, unclipped) if flags else unclipped
        clipped = TemplateOptions.make(scatter_repeats=scatter, dodgy=dodgy, mapq_threshold=mapq, clip_semialigned=bool(flags & CLIP_SEMIALIGNED),
                                       clip_overlapping=bool(flags & CLIP_OVERLAPPING))
        want = oracle_lib.build_templates(ref, g, reads, config, mb, tls, clipped, threads=4)
        W.assert_templates_equal(got, want, what)
        for i in np.nonzero(want.fragments["cigarLength"])[0]:
            assert np.array_equal(got.cigar(i), want.cigar(i)), (what, i)
        # ---- the device-side plan pass against the worker's planning mode
        built = oracle_lib.build_fragments(ref, g, reads, config, mb, threads=4)
        built_c = W.flat_view(built, BuildResult)
        n, rc = reads.cluster_count, reads.read_count
        read_length = np.array(list(reads.read_lengths), dtype=np.uint32)
        contig_length = np.array([len(c) for c in genome], dtype=np.uint64)
        a, ab = np.zeros(4 * n + 16, dtype=RESCUE_REQUEST_DTYPE), np.zeros(n + 1, dtype=np.uint64)
        b, bb = np.zeros(4 * n + 16, dtype=RESCUE_REQUEST_DTYPE), np.zeros(n + 1, dtype=np.uint64)
        assert worker.template_worker_plan(ctypes.c_uint32(n), ctypes.c_uint32(rc), W.p(read_length), ctypes.c_uint32(len(contig_length)),
                                           W.p(contig_length), ctypes.byref(tls), ctypes.byref(plain), ctypes.byref(built_c),
                                           ctypes.c_uint64(a.size), W.p(a), W.p(ab), ctypes.c_uint(2)) == 0
        assert worker.plan_device_requests(ctypes.c_uint32(n), ctypes.c_uint32(rc), ctypes.byref(tls), ctypes.byref(plain), ctypes.byref(built_c),
                                           ctypes.c_uint64(b.size), W.p(b), W.p(bb)) == 0
        assert np.array_equal(ab, bb) and a[:int(ab[-1])].tobytes() == b[:int(ab[-1])].tobytes(), what
        # ---- R1 of the rescue pass on those requests against the reference's calculateShadowRescueRange
        total = int(ab[-1])
        if total:
            req = a[:total].copy()
            tasks, got_range = np.zeros((total, 4), dtype=np.int64), np.zeros((total, 2), dtype=np.int64)
            worker.shadow_windows_device(W.p(read_length), ctypes.byref(tls), ctypes.c_uint32(total), W.p(req), W.p(contig_length), W.p(tasks), W.p(got_range))
            want_range, orientation = np.zeros((total, 2), dtype=np.int64), np.zeros(total, dtype=np.uint8)
            assert ref.lib.oracle_shadow_rescue_range(ctypes.byref(reads.c), ctypes.byref(tls), ctypes.c_uint32(total), W.p(req), W.p(want_range),
                                                      W.p(orientation)) == 0
            assert np.array_equal(got_range, want_range), what
            assert np.array_equal(tasks[:, 3] & 1, orientation) and np.array_equal(tasks[:, 0], np.maximum(0, want_range[:, 0])), what
        # ---- the record packer on the templates of this tile and on random template records
        keep, compact = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        lanes, misalign = int(rng.choice([1, 3, 7, 31, 32])), int(rng.integers(0, 8))
        options, barcode = P.make_options(rng, n, keep, with_arrays=False, compact=compact)
        wantp, mask = oracle_lib.pack_fragments(ref, reads, got, options, barcode_bytes=barcode)
        gotp = P.pack_lanes(packer, reads, got, options, lanes, misalign)
        (P.assert_compact_equal(gotp, wantp, mask, what) if compact else P.assert_packed_equal(gotp, wantp, mask, reads.read_lengths, what))
        rl = tuple(int(x) for x in rng.integers(32, 301, size=int(rng.integers(1, 3))))
        rreads = ReadSet(P.random_bcl(rng, 150, sum(rl)), rl)
        rtemplates = P.random_templates(rng, 150, rl)
        options, barcode = P.make_options(rng, 150, keep, with_arrays=True, compact=compact)
        wantp, mask = oracle_lib.pack_fragments(ref, rreads, rtemplates, options, barcode_bytes=barcode)
        gotp = P.pack_lanes(packer, rreads, rtemplates, options, lanes, misalign)
        (P.assert_compact_equal(gotp, wantp, mask, what) if compact else P.assert_packed_equal(gotp, wantp, mask, rl, what))
        if (k + 1) % 10 == 0:
            print("%d rounds clean" % (k + 1), flush=True)
    print("campaign %d: %d rounds, no difference" % (campaign, rounds))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 20261018)
