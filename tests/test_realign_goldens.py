"""Replay of tests/golden/realign_*.npz (the reference's own GapRealigner on three small bins, tests/golden/make_realign_goldens.py):
on the CPU through the device functions of csrc/realign_device.cuh (tests/cpp/test_realign_host.cu), on the GPU through
isaac_ext_realign_bin and isaac_ext_realign_bins.  Needs neither /root/reference nor the reference build of the checker."""
import ctypes
import glob
import os

import numpy as np
import pytest

from isaac_aligner_b200 import bins
from isaac_aligner_b200.batch import Tls
from test_realign_host import host_lib, host_realign         # noqa: F401  (host_lib is a fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "realign_*.npz")))


def load(path):
    g = np.load(path)
    contigs = [g["contig0"], g["contig1"]]
    b = bins.Bin(g["data"].copy(), g["record_offset"], g["index"], int(g["bin"][0]), int(g["bin"][1]))
    words = g["tls"].reshape(-1, 8)
    tls = [Tls(int(w[0]), int(w[1]), int(w[2]), int(w[3]), int(w[4]), (ctypes.c_uint32 * 2)(int(w[5]), int(w[6])), int(np.int32(w[7]))) for w in words]
    vigorous, dodgy, clip = (bool(x) for x in g["flags"])
    o = bins.RealignOptions(b.bin_start, b.bin_end, tls, vigorous=vigorous, dodgy=dodgy, clip_semialigned=clip,
                            gap_groups=list(g["groups"]) if g["groups"].size else None)
    return g, contigs, b, o


def check(g, b, got, with_gaps=True):
    assert np.array_equal(got.position, g["position"])
    assert np.array_equal(got.cigar_length, g["cigar_length"])
    realigned = np.flatnonzero(got.cigar_offset != bins.OWN_CIGAR)
    assert np.array_equal(realigned, g["realigned"])
    cigars = np.concatenate([got.cigar(i, b) for i in realigned]) if realigned.size else np.zeros(0, np.uint32)
    assert np.array_equal(cigars, g["cigars"])
    assert np.array_equal(got.data, g["data_after"])
    if with_gaps:
        assert np.array_equal(got.gaps, g["gaps"]) and np.array_equal(got.deletions, g["deletions"])


def test_goldens_are_there():
    assert len(CASES) == 3


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[8:-4] for p in CASES])
def test_device_functions_on_the_cpu_replay_the_goldens(host_lib, path):          # noqa: F811
    g, contigs, b, o = load(path)
    got, counts = host_realign(host_lib, contigs, b, o)
    assert int(counts[4]) == 0
    check(g, b, got)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[8:-4] for p in CASES])
def test_gpu_replays_the_goldens(path):
    from isaac_aligner_b200 import capi
    from isaac_aligner_b200.types import Config
    g, contigs, b, o = load(path)
    ctx = capi.Context(Config.default(max_read_length=512))
    ctx.set_reference(contigs)
    check(g, b, ctx.realign_bin(b, o))
    check(g, b, ctx.realign_bins([b, b], [o, o])[1], with_gaps=False)
    ctx.close()
