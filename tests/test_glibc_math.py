"""isaac_aligner_b200/csrc/glibc_math.cuh replays glibc 2.39's exp() / log10() (the FMA variants every x86-64 host with FMA + AVX2
runs) so that the device computes the mapping scores of TemplateBuilder with the host library's own bits.  CPU part: the host build
of the header against the libm of this box (a campaign of 300 M rounds, ~1.5 G evaluations, was clean: build/glibc_math_300M.log).
GPU part (tests/test_gpu_glibc_math.py): the device build against the same libm."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_has_fma():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " fma " in flags and " avx2 " in flags


@pytest.mark.skipif(not host_has_fma(), reason="glibc picks its non-FMA exp/log on this CPU; the replay is of the FMA variants")
def test_replay_equals_libm_bit_for_bit():
    exe = os.path.join(ROOT, "build", "test_glibc_math")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=gnu++17", "-O2", "-mfma", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_glibc_math.cpp"), "-o", exe])
    out = subprocess.run([exe, "5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr


@pytest.mark.skipif(not host_has_fma(), reason="glibc picks its non-FMA exp/log on this CPU")
def test_replay_does_not_depend_on_host_contraction():
    """built without -mfma (software fma, no contraction by the compiler) the header must give the same bits"""
    exe = os.path.join(ROOT, "build", "test_glibc_math_nofma")
    subprocess.check_call(["g++", "-std=gnu++17", "-O2", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_glibc_math.cpp"), "-o", exe])
    out = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
