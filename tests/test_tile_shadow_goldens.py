"""The literal vectors of the reference's ShadowAligner unit test (tests/golden/shadow_aligner.json, transcribed from
testShadowAligner.cpp by tests/golden/make_shadow_goldens.py): an orphan at the start of a contig, its mate 81 / 92 bases at the
shortest / longest distance the template length statistics allow, both orientations, each block rescuing the mate and then the
orphan back from the rescued mate.  Replayed through both CPU checkers (restatement and reference build) and, on the GPU,
through isaac_ext_rescue_shadows."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200.batch import RESCUE_REQUEST_DTYPE, Tls
from isaac_aligner_b200.types import Config, ReadSet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shadow_aligner.json")
COMPLEMENT = np.array([3, 2, 1, 0], dtype=np.uint8)


def golden():
    return json.load(open(GOLDEN))


def fixture(gold, case):
    """getContigList(190, 300, 422) + getBcl() of the reference's BuilderInit.hh: noise contigs, two reads cut out of one of them"""
    rng = np.random.default_rng(190300422)
    lengths = gold["contigLengths"]
    codes = [rng.integers(0, 4, size=l).astype(np.uint8) for l in lengths]
    codes[3] = np.concatenate([np.zeros(5, dtype=np.uint8), codes[2]])                 # c3 = "AAAAA" + c2
    genome = [np.frombuffer(b"ACGT", dtype=np.uint8)[c] for c in codes]
    b = case["bcl"]
    forward = codes[b["contigId"]]
    reverse = COMPLEMENT[forward[::-1]]
    l0, l1 = gold["readLengths"]
    s0, s1 = (reverse if b["reverse0"] else forward), (reverse if b["reverse1"] else forward)
    bases = np.concatenate([s0[b["offset0"]:b["offset0"] + l0], s1[b["offset1"]:b["offset1"] + l1]])
    reads = ReadSet(((gold["quality"] << 2) | bases).astype(np.uint8)[None, :], (l0, l1))
    t = case["tls"]
    tls = Tls.make(t["min"], t["max"], t["median"], t["lowStdDev"], t["highStdDev"], t["bestModel"][0], t["bestModel"][1], t["mateDriftRange"])
    config = Config.default(tuple(gold["scores"]), max_read_length=l0 + l1)
    config.gappedMismatchesMax = gold["gappedMismatchesMax"]
    return genome, reads, tls, config


def request(orphan):
    q = np.zeros(1, dtype=RESCUE_REQUEST_DTYPE)
    q["orphanPosition"], q["orphanReadId"] = orphan["position"], orphan["readIndex"]
    q["orphanContigStrand"] = (orphan["contigId"] << 1) | (1 if orphan["reverse"] else 0)
    q["orphanObservedLength"] = orphan["observedLength"]
    return q


def replay(gold, rescue):
    """rescue(genome, reads, config, tls, requests) -> FlatFragments"""
    assert len(gold["cases"]) == 4
    for case in gold["cases"]:
        genome, reads, tls, config = fixture(gold, case)
        orphan = dict(case["orphan"])
        for step, want in enumerate(case["expect"]):
            flat = rescue(genome, reads, config, tls, request(orphan))
            what = "%s call %d" % (case["name"], step)
            assert flat.flags[0] == 1 and flat.begin[1] >= 1, what                     # CPPUNIT_ASSERT(rescueShadow(...))
            f = flat.fragments[0]                                                      # shadowList[0]
            assert int(f["contigId"]) == orphan["contigId"] and int(f["readIndex"]) == (orphan["readIndex"] + 1) % 2, what
            assert (int(f["position"]), bool(f["reverse"]), int(f["observedLength"]), int(f["mismatchCount"]), int(f["cigarLength"])) == \
                (want["position"], want["reverse"], want["observedLength"], want["mismatchCount"], want["cigarLength"]), what
            assert int(flat.cigar(0)[0]) == want["cigarWord"], what
            assert abs(float(f["logProbability"]) - want["logProbability"]) <= want["tolerance"], what
            orphan = {"readIndex": int(f["readIndex"]), "contigId": int(f["contigId"]), "position": int(f["position"]),
                      "reverse": bool(f["reverse"]), "observedLength": int(f["observedLength"])}


def test_restatement_reproduces_the_shadow_aligner_literals():
    chk = oracle_lib.port()
    replay(golden(), lambda genome, reads, config, tls, req: oracle_lib.rescue_shadows(chk, oracle_lib.GenomeHolder(genome), reads, config, tls, req))


def test_reference_build_reproduces_the_shadow_aligner_literals():
    if not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"):
        pytest.skip("the reference build of the checker did not travel to this box")
    chk = oracle_lib.reference()
    replay(golden(), lambda genome, reads, config, tls, req: oracle_lib.rescue_shadows(chk, oracle_lib.GenomeHolder(genome), reads, config, tls, req))


@pytest.mark.gpu
def test_cuda_reproduces_the_shadow_aligner_literals():
    from isaac_aligner_b200 import capi

    def rescue(genome, reads, config, tls, req):
        ctx = capi.Context(config)
        ctx.set_reference(genome)
        ctx.set_reads(reads)
        flat = ctx.rescue_shadows(tls, req)
        ctx.close()
        return flat

    replay(golden(), rescue)
