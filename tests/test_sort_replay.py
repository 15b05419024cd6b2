"""isaac_aligner_b200/csrc/sort_replay.cuh: libstdc++'s std::sort replayed for host and device code must give the same
permutation as std::sort itself (the candidate lists of FragmentBuilder are full of equivalent entries and the first of each
group survives).  Host-only check, compiled with g++."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sort_replay_matches_std_sort():
    exe = os.path.join(ROOT, "build", "test_sort_replay")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_sort_replay.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr


def test_sort_replay_compiles_for_the_device():
    """the same header through nvcc for sm_100a (device code generation only, no GPU needed)"""
    src = os.path.join(ROOT, "build", "sort_replay_device.cu")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    with open(src, "w") as f:
        f.write('#include "../isaac_aligner_b200/csrc/sort_replay.cuh"\n'
                'struct Key { int k, id; };\n'
                '__global__ void sortLists(Key *lists, const unsigned *begin, unsigned n)\n'
                '{\n'
                '    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;\n'
                '    if (i < n) isaac_b200::sort_replay::sort(lists + begin[i], begin[i + 1] - begin[i], [](const Key &a, const Key &b) { return a.k < b.k; });\n'
                '}\n')
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--extended-lambda",
                           "-c", src, "-o", os.path.join(ROOT, "build", "sort_replay_device.o")])


def test_consolidate_replay_matches_the_host_consolidate():
    """consolidate_device.cuh (the replay under consolidateDuplicateFragments) against the std::sort-based function the product's
    host phases use, on 300 k random candidate lists; host code of an nvcc-compiled binary, no GPU involved"""
    exe = os.path.join(ROOT, "build", "test_consolidate_replay")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-cudart", "shared", os.path.join(ROOT, "tests", "cpp", "test_consolidate_replay.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
