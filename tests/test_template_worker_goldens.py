"""The reference's TemplateBuilder unit test (lib/alignment/cppunit/testTemplateBuilder.cpp:113-343) replayed on the host half of
isaac_ext_build_templates (csrc/template_worker.cuh) on the CPU: the test hands hand-made candidate lists (f0_0: read 1 at 2,
100 bases, log probability -8, three unique seeds; f0_1: read 2 at 107 reverse, 99 bases, -12, one seed; :60-61) to buildTemplate
and asserts the mapping scores 1136 / 534 / 569, 1119 / 517 / 569, 1084 / 534 / 517 and 2 / 2 / 3.  The flat build result is made
by hand the same way, the plan / finish passes run through tests/cpp/test_template_worker.cu, and the one real rescueShadow of
testOrphan is answered by the checker on the fixture's cluster (BuilderInit.hh: reads cut from contig 0 at 2 and 107)."""
import ctypes

import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200.batch import (DODGY_ALIGNMENT_SCORE_UNALIGNED, RESCUE_REQUEST_DTYPE, TEMPLATE_DTYPE, BuildResult, FlatFragments,
                                      RescueResult, Templates, Tls, TemplateOptions)
from isaac_aligner_b200.types import ELAND_SCORES, FRAGMENT_DTYPE, Config, ReadSet
from test_template_worker import flat_view, p, worker_lib          # noqa: F401  (fixture)
from test_tile_fragment_builder_scenarios import clusters, contigs

NO_MATCH_CONTIG = 0x7FFFFF


def fragment(contig, position, observed, read_index, reverse, cigar_offset, mismatches, log_probability, unique_seeds):
    """getFragmentMetadata (testTemplateBuilder.cpp:21-50); everything else as FragmentMetadata() leaves it"""
    f = np.zeros(1, dtype=FRAGMENT_DTYPE)[0]
    f["contigId"], f["position"], f["observedLength"], f["readIndex"], f["reverse"] = contig, position, observed, read_index, reverse
    f["cigarOffset"], f["cigarLength"], f["mismatchCount"], f["logProbability"], f["uniqueSeedCount"] = cigar_offset, 1, mismatches, log_probability, unique_seeds
    f["readId"], f["firstSeedIndex"], f["nonUniqueSeedOffsetFirst"] = read_index, -1, 0xFFFF
    return f


F0_0 = fragment(0, 2, 100, 0, 0, 0, 0, -8.0, 3)                      # :60
F0_1 = fragment(0, 107, 99, 1, 1, 1, 2, -12.0, 1)                    # :61


def build_template(lib, lists):
    """buildTemplate(contigList, restOfGenomeCorrection, readMetadataList, adapters, fragments, cluster0, tls) -> Templates"""
    codes = contigs()
    genome = [np.frombuffer(b"ACGT", dtype=np.uint8)[c] for c in codes]
    reads = ReadSet(clusters(codes)["cluster0"][None, :], (100, 100))
    config = Config.default(ELAND_SCORES, max_read_length=200)           # TemplateBuilder(flowcells, 10, 4, false, 8, false, ELAND..., 20000, UNALIGNED)
    config.repeatThreshold, config.maxSeedsPerRead, config.gappedMismatchesMax, config.semialignedGapLimit = 10, 4, 8, 20000
    tls = Tls.make(150, 250, 190, 20, 30)                                # DummyTemplateLengthStatistics (:345-356): FR+ / RF-
    options = TemplateOptions.make(dodgy=DODGY_ALIGNMENT_SCORE_UNALIGNED)
    frags = np.array(list(lists[0]) + list(lists[1]), dtype=FRAGMENT_DTYPE) if lists[0] or lists[1] else np.zeros(0, dtype=FRAGMENT_DTYPE)
    begin = np.array([0, len(lists[0]), len(lists[0]) + len(lists[1])], dtype=np.uint64)
    built = FlatFragments(frags, begin, np.full(1000, 1600, dtype=np.uint32), np.array([1 if len(frags) else 0], dtype=np.uint8))   # cigarBuffer(1000, 1600)
    built_c = flat_view(built, BuildResult)
    read_length = np.array([100, 100], dtype=np.uint32)
    contig_length = np.array([len(c) for c in genome], dtype=np.uint64)
    requests = np.zeros(64, dtype=RESCUE_REQUEST_DTYPE)
    request_begin = np.zeros(2, dtype=np.uint64)
    head = [ctypes.c_uint32(1), ctypes.c_uint32(2), p(read_length), ctypes.c_uint32(len(contig_length)), p(contig_length),
            ctypes.byref(tls), ctypes.byref(options), ctypes.byref(built_c)]
    assert lib.template_worker_plan(*head, ctypes.c_uint64(requests.size), p(requests), p(request_begin), ctypes.c_uint(1)) == 0
    requests = requests[:int(request_begin[-1])].copy()
    rescued = oracle_lib.rescue_shadows(oracle_lib.reference(), oracle_lib.GenomeHolder(genome), reads, config, tls, requests)
    rescued_c = flat_view(rescued, RescueResult)
    templates, fragments = np.zeros(1, dtype=TEMPLATE_DTYPE), np.zeros(2, dtype=FRAGMENT_DTYPE)
    cigars, words = np.zeros(1024, dtype=np.uint32), ctypes.c_uint64()
    assert lib.template_worker_finish(*head, ctypes.byref(rescued_c), p(request_begin), p(templates), p(fragments),
                                      ctypes.c_uint64(cigars.size), p(cigars), ctypes.byref(words), ctypes.c_uint(1)) == 0
    return Templates(templates, fragments, cigars[:words.value].copy(), len(requests))


def check_fragment(t, i, want, score):
    f = t.fragments[i]
    for name in ("contigId", "position", "observedLength", "readIndex", "reverse", "cigarLength", "mismatchCount", "uniqueSeedCount"):
        assert int(f[name]) == int(want[name]), (i, name)
    assert float(f["logProbability"]) == float(want["logProbability"]), i
    assert int(t.templates["fragmentAlignmentScore"][0][i]) == score, i


def check_unaligned(t):
    """checkUnalignedTemplate (:90-111)"""
    assert int(t.templates["alignmentScore"][0]) == 0
    for i in range(2):
        f = t.fragments[i]
        assert int(f["contigId"]) == NO_MATCH_CONTIG and int(f["readIndex"]) == i
        assert (int(f["observedLength"]), int(f["reverse"]), int(f["cigarLength"]), int(f["mismatchCount"]), int(f["uniqueSeedCount"])) == (0, 0, 0, 0, 0)
        assert float(f["logProbability"]) == 0.0 and int(t.templates["fragmentAlignmentScore"][0][i]) == 0xFFFFFFFF


pytestmark = pytest.mark.skipif(oracle_lib.reference() is None, reason="the rescue of testOrphan is answered by the reference build")


def test_empty_match_list(worker_lib):                                   # :120-146
    check_unaligned(build_template(worker_lib, ([], [])))
    build_template(worker_lib, ([F0_0], [F0_1]))
    check_unaligned(build_template(worker_lib, ([], [])))


def test_orphan(worker_lib):                                             # :148-207
    t = build_template(worker_lib, ([F0_0], []))
    assert int(t.templates["alignmentScore"][0]) == 1136
    check_fragment(t, 0, F0_0, 534)
    assert int(t.templates["fragmentAlignmentScore"][0][1]) == 569
    t = build_template(worker_lib, ([], [F0_1]))
    assert int(t.templates["alignmentScore"][0]) == 1119
    check_fragment(t, 1, F0_1, 517)
    assert int(t.templates["fragmentAlignmentScore"][0][0]) == 569


def test_unique(worker_lib):                                             # :209-250
    t = build_template(worker_lib, ([F0_0], [F0_1]))
    assert int(t.templates["alignmentScore"][0]) == 1084
    check_fragment(t, 0, F0_0, 534)
    check_fragment(t, 1, F0_1, 517)


def test_multiple(worker_lib):                                           # :257-343: the pair with the best log probability wins
    lists = ([], [])
    t0, t1 = F0_0.copy(), F0_1.copy()
    for _ in range(2):
        lists[0].append(t0.copy()); t0["position"] += 56
        lists[0].append(t0.copy()); t0["position"] += 65
        lists[1].append(t1.copy()); t1["position"] += 300
    t0, t1 = F0_0.copy(), F0_1.copy()
    t0["contigId"] = t1["contigId"] = 1
    for _ in range(2):
        t0["position"] += 56; lists[0].append(t0.copy())
        t0["position"] += 65; lists[0].append(t0.copy())
        t1["position"] += 401; lists[1].append(t1.copy())
    t0, t1 = F0_0.copy(), F0_1.copy()
    t0["contigId"] = t1["contigId"] = 1
    t0["logProbability"] += 2; t1["logProbability"] += 2
    lists[0].append(t0.copy()); lists[1].append(t1.copy())
    best0, best1 = t0.copy(), t1.copy()
    t0["logProbability"] -= 2; t1["logProbability"] -= 2
    for _ in range(2):
        t0["position"] += 36; lists[0].append(t0.copy())
        t0["position"] += 45; lists[0].append(t0.copy())
        t1["position"] += 402; lists[1].append(t1.copy())
    t = build_template(worker_lib, lists)
    assert int(t.templates["alignmentScore"][0]) == 2
    check_fragment(t, 0, best0, 2)
    check_fragment(t, 1, best1, 3)
