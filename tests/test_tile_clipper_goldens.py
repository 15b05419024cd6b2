"""The literal vectors of the reference's semialigned-clipper unit test (tests/golden/semialigned_clipper.json, transcribed from
testSemialignedClipper.cpp by tests/golden/make_clipper_goldens.py) replayed through the template call: a single-ended cluster
with one seed match at the test's position, FragmentBuilder::build without gaps, pickBestFragment, then
SemialignedEndsClipper::clip (MatchSelector.cpp:336-340).  CPU: the reference build of the checker must give the literals (pins
the checker's driver); GPU: isaac_ext_build_templates with ISAAC_EXT_CLIP_SEMIALIGNED must give them too."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200.batch import MatchBatch, Tls, TemplateOptions
from isaac_aligner_b200.synth import MATCH_DTYPE, SEED_DTYPE
from isaac_aligner_b200.types import Config, ReadSet, cigar_to_string

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "semialigned_clipper.json")
SEED_OFFSET, SEED_LENGTH = 32, 32


def golden():
    return json.load(open(GOLDEN))


def inputs(case, qualities):
    """genome, reads, matches of one case: the read starts 'blanks' bases in front of the contig"""
    assert not case["reverse"]
    reference = case["reference"].lstrip(" ")
    position = -(len(case["reference"]) - len(reference))
    read = case["read"]
    bcl = np.array([((ord(q) - 33) << 2) | "ACGT".index(b) for b, q in zip(read, qualities)], dtype=np.uint8)[None, :]
    reads = ReadSet(bcl, (len(read),))
    genome = [np.frombuffer(reference.encode(), dtype=np.uint8)]
    seeds = np.zeros(1, dtype=SEED_DTYPE)
    seeds[0] = (SEED_OFFSET, SEED_LENGTH, 0)
    matches = np.zeros(1, dtype=MATCH_DTYPE)
    matches["seedId"] = 0                                                # cluster 0, seed 0, forward
    matches["location"] = ((1 << 40) | (position + SEED_OFFSET)) << 1    # contig 0, no neighbours: the read at 'position'
    return genome, reads, MatchBatch(matches, np.array([0, 1], dtype=np.uint64), seeds, with_gaps=False)


def check(case, templates):
    f = templates.fragments[0]
    assert templates.templates["built"][0] == 1, case["name"]
    assert cigar_to_string(templates.cigar(0)) == case["cigar"], case["name"]
    assert (int(f["contigId"]), int(f["position"])) == (0, case["position"]), case["name"]


def test_reference_build_reproduces_the_semialigned_clipper_literals():
    if not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"):
        pytest.skip("the reference build of the checker did not travel to this box")
    ref, gold = oracle_lib.reference(), golden()
    assert len(gold["cases"]) == 4
    for case in gold["cases"]:
        genome, reads, mb = inputs(case, gold["qualities"])
        cfg = Config.default(tuple(gold["scores"]), max_read_length=len(case["read"]))
        want = oracle_lib.build_templates(ref, oracle_lib.GenomeHolder(genome), reads, cfg, mb, Tls.make(),
                                          TemplateOptions.make(clip_semialigned=True))
        check(case, want)
        unclipped = oracle_lib.build_templates(ref, oracle_lib.GenomeHolder(genome), reads, cfg, mb, Tls.make(), TemplateOptions.make())
        assert cigar_to_string(unclipped.cigar(0)) != case["cigar"], case["name"]       # it is the clipper that makes the literal


@pytest.mark.gpu
def test_cuda_reproduces_the_semialigned_clipper_literals():
    from isaac_aligner_b200 import capi
    gold = golden()
    for case in gold["cases"]:
        genome, reads, mb = inputs(case, gold["qualities"])
        ctx = capi.Context(Config.default(tuple(gold["scores"]), max_read_length=len(case["read"])))
        ctx.set_reference(genome)
        ctx.set_reads(reads)
        check(case, ctx.build_templates(mb, Tls.make(), TemplateOptions.make(clip_semialigned=True)))
        ctx.close()
