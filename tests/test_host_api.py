"""The C++ classes that carry the reference's names (isaac_aligner_b200/host/isaac_b200.hh): they must compile against
the C ABI on any box, and on a GPU box the reference's unit tests restated on them (tests/cpp/test_host_api.cpp) must
pass."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
EXE = os.path.join(ROOT, "build", "test_host_api")
LIB_DIR = os.path.join(ROOT, "isaac_aligner_b200")


def compile_host_test():
    import __graft_entry__
    __graft_entry__.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", SRC, "-o", EXE, "-L" + LIB_DIR, "-lisaac_ext",
                           "-Wl,-rpath," + LIB_DIR, "-Wl,-rpath,/usr/local/cuda/lib64"])
    return EXE


def test_host_classes_compile_and_link():
    assert os.path.exists(compile_host_test())


@pytest.mark.gpu
def test_reference_unit_tests_through_host_classes():
    exe = compile_host_test()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all checks passed" in out.stdout
