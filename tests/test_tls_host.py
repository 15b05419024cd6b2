"""The reference's TemplateLengthStatistics unit test (lib/alignment/cppunit/testTemplateLengthStatistics.cpp:108-171: addTemplates,
testStatistics, testMateDriftRange) replayed on the product's host-side TemplateLengthDistribution
(csrc/template_length_host.cuh, what isaac_ext_determine_template_length keeps on the host) on the CPU: 10 000 hand-made pairs of
one-base fragments at growing distance, the asserted statistics 14 / 5001 / 9987 / 3414 / 3413, stability at the second update."""
import ctypes
import os
import subprocess

import numpy as np

from isaac_aligner_b200.types import FRAGMENT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    so = os.path.join(ROOT, "build", "libtest_tls_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-shared", "-fPIC", os.path.join(ROOT, "tests", "cpp", "test_tls_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.tls_host_new.restype = ctypes.c_void_p
    lib.tls_host_free.argtypes = [ctypes.c_void_p]
    lib.tls_host_add.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int]
    lib.tls_host_add.restype = ctypes.c_uint
    lib.tls_host_get.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return lib


def add_templates(lib, h):
    """TestTemplateLengthStatistics::addTemplates (:108-146)"""
    cigars = np.array([16], dtype=np.uint32)                            # cigarBuffer(1, 16) = 1M
    f = np.zeros(2, dtype=FRAGMENT_DTYPE)
    f["observedLength"], f["cigarLength"] = 1, 1
    f["reverse"][1] = 1
    p = lambda i: ctypes.c_void_p(f[i:i + 1].ctypes.data)
    stats = np.zeros(8, dtype=np.uint32)
    get = lambda: (lib.tls_host_get(h, ctypes.c_void_p(stats.ctypes.data)), stats.copy())[1]
    assert lib.tls_host_add(h, p(0), p(1), cigars.ctypes.data, 9999, 1) == 0              # :123-127
    assert int(f["position"][1]) == 9999
    assert lib.tls_host_add(h, p(1), p(0), cigars.ctypes.data, 1, -1) == 0                # std::swap(f[0], f[1]) (:128-129)
    assert get()[:5].tolist() == [14, 5001, 9987, 3414, 3413]                             # :130-134: min, median, max, low, high
    f["position"][1] = f["position"][0]                                                   # :135-136
    assert lib.tls_host_add(h, p(0), p(1), cigars.ctypes.data, 9999, 1) == 0              # :137-141
    assert lib.tls_host_add(h, p(0), p(1), cigars.ctypes.data, 1, -1) == 1                # :142: stable at the second update
    return get()


def test_statistics():
    lib = load()
    h = lib.tls_host_new(-1)
    s = add_templates(lib, h)
    assert s[:5].tolist() == [14, 5001, 9987, 3414, 3413] and s[7] == 1                   # testStatistics (:148-158)
    assert sorted(s[5:7].tolist()) == [1, 6]                                              # the two models the pairs were: FR+ and RF-
    lib.tls_host_free(h)


def test_mate_drift_range():
    """testMateDriftRange (:160-171): mateMin / mateMax = median -/+ the drift range (TemplateLengthStatistics.hh:205-214); the
    library hands the range through in isaac_ext_tls_t::mateDriftRange"""
    lib = load()
    h = lib.tls_host_new(123)
    s = add_templates(lib, h)
    assert int(s[1]) == 5001
    lib.tls_host_free(h)
