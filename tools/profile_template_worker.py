"""Times the plan and finish passes of isaac_ext_build_templates (csrc/template_worker.cuh) on the CPU, around build / rescue
results of the reference build (tests/cpp/test_template_worker.cu; no GPU).  Development aid for DESIGN.md section 9 item 1.

    python tools/profile_template_worker.py [pairs] [threads] [repeats]
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib                                                    # noqa: E402
import test_template_worker as T                                     # noqa: E402
from isaac_aligner_b200.batch import RESCUE_REQUEST_DTYPE, TEMPLATE_DTYPE, BuildResult, RescueResult, Tls, TemplateOptions   # noqa: E402
from isaac_aligner_b200.types import FRAGMENT_DTYPE, Config         # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    import bench
    args = bench.parse_args.__wrapped__() if hasattr(bench.parse_args, "__wrapped__") else None
    sys.argv = [sys.argv[0]]
    args = bench.parse_args()
    genome, reads, mb, tls = bench.make_pairs_workload(args, 0, n_pairs)
    lib = ctypes.CDLL(os.path.join(ROOT, "build", "libtest_template_worker.so"))
    ref = oracle_lib.reference()
    g = oracle_lib.GenomeHolder(genome)
    config = Config.default(max_read_length=2 * args.read_length)
    options = TemplateOptions.make()
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    built = oracle_lib.build_fragments(ref, g, reads, config, mb, threads=cores)
    print("checker build_fragments: %.2f s" % (time.perf_counter() - t0))
    built_c = T.flat_view(built, BuildResult)
    n, rc = reads.cluster_count, reads.read_count
    read_length = np.array(list(reads.read_lengths), dtype=np.uint32)
    contig_length = np.array([len(c) for c in genome], dtype=np.uint64)
    requests = np.zeros(4 * n + 16, dtype=RESCUE_REQUEST_DTYPE)
    request_begin = np.zeros(n + 1, dtype=np.uint64)
    head = [ctypes.c_uint32(n), ctypes.c_uint32(rc), T.p(read_length), ctypes.c_uint32(len(contig_length)), T.p(contig_length),
            ctypes.byref(tls), ctypes.byref(options), ctypes.byref(built_c)]
    plan = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        assert lib.template_worker_plan(*head, ctypes.c_uint64(requests.size), T.p(requests), T.p(request_begin), ctypes.c_uint(threads)) == 0
        plan.append(time.perf_counter() - t0)
    req = requests[:int(request_begin[-1])].copy()
    t0 = time.perf_counter()
    rescued = oracle_lib.rescue_shadows(ref, g, reads, config, tls, req, threads=cores, fragments_per_request=96)
    print("checker rescue_shadows of %d requests: %.2f s" % (len(req), time.perf_counter() - t0))
    rescued_c = T.flat_view(rescued, RescueResult)
    templates = np.zeros(n, dtype=TEMPLATE_DTYPE)
    fragments = np.zeros(n * rc, dtype=FRAGMENT_DTYPE)
    cigars = np.zeros(64 * n + 1024, dtype=np.uint32)
    words = ctypes.c_uint64()
    finish = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        assert lib.template_worker_finish(*head, ctypes.byref(rescued_c), T.p(request_begin), T.p(templates), T.p(fragments),
                                          ctypes.c_uint64(cigars.size), T.p(cigars), ctypes.byref(words), ctypes.c_uint(threads)) == 0
        finish.append(time.perf_counter() - t0)
    per = lambda t: min(t) * threads / n * 1e6
    print("%d pairs, %d threads: plan %.2f ms (%.2f us per cluster and thread), finish %.2f ms (%.2f us); %d requests, %d candidate "
          "fragments, %d shadow fragments, %d templates built"
          % (n, threads, min(plan) * 1e3, per(plan), min(finish) * 1e3, per(finish), len(req), built.fragments.size,
             rescued.fragments.size, int(templates["built"].sum())))


if __name__ == "__main__":
    main()
