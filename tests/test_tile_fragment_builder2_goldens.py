"""The literal vectors of the reference's second FragmentBuilder unit test (tests/golden/fragment_builder2.json, transcribed from
testFragmentBuilder2.cpp by tests/golden/make_fragment_builder2_goldens.py): UngappedAligner / GappedAligner on one read at a
given position -- 30 mismatches, the mismatch cycle of a reverse read, a leading soft clip in front of the contig, a 1-base
deletion found by the gapped aligner, the same with 'n' bases.  Replayed through both CPU checkers and, on the GPU, through
isaac_ext_ungapped_batch / isaac_ext_gapped_batch."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200.types import CANDIDATE_DTYPE, Config, ReadSet, cigar_to_string

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fragment_builder2.json")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "n": "n"}


def golden():
    return json.load(open(GOLDEN))


def inputs(gold, case):
    """one single-read cluster whose strand sequence is the case's read string ('n' = BCL N), at the case's start position"""
    read, quals = case["read"], gold["qualities"]
    if case["reverse"]:
        read, quals = "".join(COMP[b] for b in reversed(read)), quals[::-1]
    bcl = np.array([0 if b == "n" else ((ord(q) - 33) << 2) | "ACGT".index(b) for b, q in zip(read, quals)], dtype=np.uint8)[None, :]
    reads = ReadSet(bcl, (len(read),), first_cycles=(gold["firstCycle"],))
    genome = [np.frombuffer(case["reference"].encode(), dtype=np.uint8)]
    cand = np.zeros(1, dtype=CANDIDATE_DTYPE)
    cand["position"] = case["startPosition"]
    cand["contigStrand"] = 1 if case["reverse"] else 0
    return genome, reads, cand


def first_mismatch_cycle(gold, case, mask):
    bits = [64 * w + b for w in range(len(mask)) for b in range(64) if (int(mask[w]) >> b) & 1]
    L = len(case["read"])
    cycles = [gold["firstCycle"] + L - 1 - i if case["reverse"] else gold["firstCycle"] + i for i in bits]   # AlignerBase.cpp:171
    return cycles[0], len(cycles)                                        # addMismatchCycle order = increasing strand index


def replay(gold, make):
    """make(genome, reads, config) -> object with ungapped(cand) / gapped(cand)"""
    assert len(gold["cases"]) == 5
    for case in gold["cases"]:
        genome, reads, cand = inputs(gold, case)
        aligner = make(genome, reads, Config.default(tuple(gold["scores"]), max_read_length=2 * len(case["read"])))
        frag, cigar, mask = aligner.ungapped(cand)
        if case["gapped"]:                                               # the harness' acceptance (testFragmentBuilder2.cpp:193-200)
            gfrag, gcigar, gmask = aligner.gapped(cand)
            assert int(gfrag["mismatchCount"][0]) <= 5 and int(frag["mismatchCount"][0]) > int(gfrag["mismatchCount"][0]), case["name"]
            assert float(frag["logProbability"][0]) < float(gfrag["logProbability"][0]), case["name"]
            assert int(gfrag["matchCount"][0]) + 16 > int(frag["observedLength"][0]), case["name"]
            frag, cigar, mask = gfrag, gcigar, gmask
        assert cigar_to_string(cigar[0][:frag["cigarLength"][0]]) == case["cigar"], case["name"]
        for key in ("mismatchCount", "editDistance", "observedLength", "position"):
            assert int(frag[key][0]) == case[key], (case["name"], key)
        if case["firstMismatchCycle"] is not None:
            assert first_mismatch_cycle(gold, case, mask[0]) == (case["firstMismatchCycle"], case["mismatchCount"]), case["name"]
        if hasattr(aligner, "close"):
            aligner.close()


class Checker:
    def __init__(self, chk, genome, reads, config):
        self.chk, self.g, self.reads, self.config = chk, oracle_lib.GenomeHolder(genome), reads, config

    def ungapped(self, cand):
        return self.chk.ungapped(self.g, self.reads, self.config, cand)

    def gapped(self, cand):
        return self.chk.gapped(self.g, self.reads, self.config, cand)


def test_restatement_reproduces_the_fragment_builder2_literals():
    chk = oracle_lib.port()
    replay(golden(), lambda genome, reads, config: Checker(chk, genome, reads, config))


def test_reference_build_reproduces_the_fragment_builder2_literals():
    if not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"):
        pytest.skip("the reference build of the checker did not travel to this box")
    chk = oracle_lib.reference()
    replay(golden(), lambda genome, reads, config: Checker(chk, genome, reads, config))


@pytest.mark.gpu
def test_cuda_reproduces_the_fragment_builder2_literals():
    from isaac_aligner_b200 import capi

    def make(genome, reads, config):
        ctx = capi.Context(config)
        ctx.set_reference(genome)
        ctx.set_reads(reads)
        return ctx

    replay(golden(), make)
