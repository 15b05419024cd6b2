#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the reference's own unit tests and the reference's own code.

Run HERE (where /root/reference is mounted and oracle/_ref/libisaac_ref.so can be built); the JSON files it writes are
committed and are what the test-suite reads -- nothing under tests/ touches /root/reference at run time.

  banded_sw.json      the scenarios of lib/alignment/cppunit/testBandedSmithWaterman.cpp (testUngapped :79-103,
                      testSingleDeletion :105-131, testSingleInsertion :133-154, testMultipleIndels :156-212, testCustom
                      :60-77) rebuilt on a seeded genome, with the CIGAR the unit test asserts (structural, independent of
                      the genome content) next to the output of the reference's BandedSmithWaterman::align
  simple_indels.json  the literal (read, reference) pairs of lib/alignment/cppunit/testSimpleIndelAligner.cpp:264-615 that
                      use the default two seeds, parsed from the test source, with the CIGAR / mismatch count / edit
                      distance the unit test asserts next to the output of the reference's FragmentBuilder::build on the
                      equivalent two-match batch
  templates.json      the reference's TemplateBuilder (buildFragments + buildTemplate per cluster, verbatim sources) on a
                      seeded 200-pair workload of tests/common_build.py for three option sets: template / fragment mapping
                      scores, proper-pair flags, placements and CIGARs of both reads; the inputs are regenerated from the
                      seeds at test time and checked against the recorded SHA-256
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from isaac_aligner_b200.batch import MatchBatch  # noqa: E402
from isaac_aligner_b200.synth import MATCH_DTYPE, SEED_DTYPE  # noqa: E402
from isaac_aligner_b200.types import Config, ReadSet, cigar_to_string  # noqa: E402

CPPUNIT = "/root/reference/src/c++/lib/alignment/cppunit"
SW_SCORES = (2, -1, 15, 3)                     # testBandedSmithWaterman.cpp:45-50
INDEL_SCORES = (0, -1, -2, -1, -5)             # testSimpleIndelAligner.cpp:138-142


def sw_cases():
    rng = np.random.default_rng(20141504)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, size=1000))
    g = lambda a, n: genome[a:a + n]
    cases = []

    def add(name, query, database, expected):
        assert len(database) == len(query) + 15
        cases.append({"name": name, "query": query, "database": database, "expected": expected})

    database = g(100, 115)                                                   # testUngapped
    for i in range(16):
        add("ungapped_offset_%d" % i, database[i:i + 100], database, "100M")
    deletion = "AGAGCAGCGAGCGACAGCAGCAGCAAA"                                 # testSingleDeletion
    for dlen in range(1, 14):
        dl = 7 - dlen // 2
        left_s = g(100 + dl, 39) + "T"
        right_s = g(100 + dl + 40, 40)
        db = g(100, dl) + left_s + deletion[:dlen] + right_s + g(100 + dl + 80, 15 - dl - dlen)
        add("single_deletion_%d" % dlen, left_s + right_s, db, "40M%dD40M" % dlen)
    database = g(100, 220)                                                   # testSingleInsertion
    qlen = len(database) - 15
    for ilen in range(1, 10):
        left, dl = 100, 9
        right = qlen - left - ilen
        add("single_insertion_%d" % ilen, database[dl:dl + left] + "T" * ilen + database[left + dl:left + dl + right],
            database, "%dM%dI%dM" % (left, ilen, right))
    dl = 6                                                                   # testMultipleIndels
    dl_s = g(100, dl)
    left_s = g(100 + dl, 19) + "T"
    center_s = g(100 + dl + 20, 19) + "T"
    right_s = g(100 + dl + 40, 20)
    tail = lambda n: g(100 + dl + 60, n)
    add("insertion_and_deletion", left_s + "A" + center_s + right_s, dl_s + left_s + center_s + "ACAG" + right_s + tail(15 - dl + 1 - 4),
        "20M1I20M4D20M")
    add("two_insertions", left_s + "A" + center_s + "CG" + right_s, dl_s + left_s + center_s + right_s + tail(15 - dl + 1 + 2),
        "20M1I20M2I20M")
    add("two_deletions", left_s + center_s + right_s, dl_s + left_s + "AAG" + center_s + "ACAG" + right_s + tail(15 - dl - 3 - 4),
        "20M3D20M4D20M")
    # testCustom is disabled in the reference (testBandedSmithWaterman.hh:27) and its expectation is stale; the strings
    # are kept as a known-answer vector of the current code
    add("custom_disabled_in_reference",
        "CTAAGACCCCACACTCTGGGACACCAAGGTGGGAGGATCGCTGGAGCTCAGGAGTTTGAGACCAGCCTGGACAACATGGTGTGACCCTGTCTACAGAAAA",
        "AATGCCTCTGGCCTGGGCGTGGGAGTTCATGCTTGTAATCGCATATCGCTAGAGCCCAGGAGTTTGAGACCAGCCTGGACAACATGGTGAAAACCCTCGTTGCTACTAAAAATAC",
        None)
    return cases


def parse_simple_indel_vectors():
    text = open(os.path.join(CPPUNIT, "testSimpleIndelAligner.cpp")).read()
    body = text[text.index("void TestSimpleIndelAligner::testEverything()"):]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    # innermost { } blocks that contain an align( call
    blocks, stack = [], []
    for i, ch in enumerate(body):
        if ch == "{":
            stack.append(i)
        elif ch == "}" and stack:
            start = stack.pop()
            inner = body[start + 1:i]
            if "align(" in inner and "{" not in inner.replace("{", "", 0)[0:0] and inner.count("align(") == 1 and "{" not in inner:
                blocks.append(inner)
    vectors = []
    for block in blocks:
        before, call = block.split("align(", 1)
        if "leftClipped()" in before or "rightClipped()" in before:
            continue                                    # alignment-independent clipping preset by the test: no build() equivalent
        depth, args, cur, in_str = 1, [], "", False
        rest = ""
        for j, ch in enumerate(call):
            if ch == '"':
                in_str = not in_str
            if not in_str:
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                    if depth == 0:
                        args.append(cur)
                        rest = call[j + 1:]
                        break
                elif ch == "," and depth == 1:
                    args.append(cur)
                    cur = ""
                    continue
            cur += ch
        if len(args) < 3:
            continue
        read = "".join(re.findall(r"\"([^\"]*)\"", args[0]))
        ref = "".join(re.findall(r"\"([^\"]*)\"", args[1]))
        seeds = None
        if len(args) == 4:
            seeds = [(int(a), int(b)) for a, b in re.findall(r"SeedMetadata\(\s*(\d+),\s*(\d+),\s*0,\s*\d+\)", before)]
            if len(seeds) != 2:
                continue
        cigar = re.search(r"std::string\(\"([^\"]+)\"\), fragmentMetadataList\[0\]\.getCigarString\(\)", rest)
        mm = re.search(r"\((\d+)U, fragmentMetadataList\[0\]\.getMismatchCount\(\)", rest)
        ed = re.search(r"\((\d+)U, fragmentMetadataList\[0\]\.getEditDistance\(\)", rest)
        pos = re.search(r"\((\d+)UL, fragmentMetadataList\[0\]\.getFStrandReferencePosition\(\)\.getPosition\(\)", rest)
        if not cigar:
            continue
        vectors.append({"read": read, "reference": ref, "cigar": cigar.group(1), "seeds": seeds,
                        "mismatches": int(mm.group(1)) if mm else None, "editDistance": int(ed.group(1)) if ed else None,
                        "position": int(pos.group(1)) if pos else None})
    return vectors


def indel_batch(v):
    """the two-candidate setup of TestSimpleIndelAligner::align (:185-215) as a FragmentBuilder::build batch"""
    read, reference = v["read"], v["reference"]
    ref_off = len(reference) - len(reference.lstrip(" "))
    ref_ns = reference[ref_off:]
    pos = len(read) - len(read.lstrip(" "))
    read_ns = read[pos:]
    L = len(read_ns)
    head = pos - ref_off
    tail = len(reference) - L - ref_off
    seed_len = min(L, len(reference) - pos)
    if v.get("seeds"):
        seeds = np.array([(v["seeds"][0][0], v["seeds"][0][1], 0), (v["seeds"][1][0], v["seeds"][1][1], 0)], dtype=SEED_DTYPE)
    else:
        seeds = np.array([(0, 32, 0), (seed_len - 32 - 1, 32, 0)], dtype=SEED_DTYPE)     # getSeedMetadataList (:47-55)
    locs = [(0, head + int(seeds[0]["offset"])), (1, tail + int(seeds[1]["offset"]))]
    if min(l for _, l in locs) < 0:
        return None
    bcl = np.array([(35 << 2) | "ACGT".index(c) for c in read_ns], dtype=np.uint8)[None, :]
    reads = ReadSet(bcl, (L,))
    m = np.zeros(2, dtype=MATCH_DTYPE)
    for k, (seed, loc) in enumerate(sorted(locs, key=lambda x: (x[1], x[0]))):
        m[k] = ((seed << 1), ((((0 + 1) << 40) | loc) << 1))
    batch = MatchBatch(m, np.array([0, 2], dtype=np.uint64), seeds, with_gaps=False)
    genome = [np.frombuffer(ref_ns.encode(), dtype=np.uint8)]
    cfg = Config.default(INDEL_SCORES, max_read_length=L)
    cfg.semialignedGapLimit = 20000                                                   # :158-159
    return genome, reads, batch, cfg


def main():
    ref = oracle_lib.reference()
    assert ref is not None, "needs /root/reference to build oracle/_ref/libisaac_ref.so"
    out = []
    for c in sw_cases():
        cig, n, off = ref.banded_sw([c["query"].encode()], [c["database"].encode()], SW_SCORES, max_read_length=300)
        words = [int(w) for w in cig[0][:n[0]]]
        got = cigar_to_string(words)
        if c["expected"] is not None:
            assert got == c["expected"], (c["name"], got, c["expected"])
        c.update({"scores": SW_SCORES, "cigar": words, "cigarString": got, "offset": int(off[0])})
        out.append(c)
    json.dump({"source": "lib/alignment/cppunit/testBandedSmithWaterman.cpp", "cases": out},
              open(os.path.join(HERE, "banded_sw.json"), "w"), indent=1)
    print("banded_sw.json: %d cases, all literal expectations reproduced by the reference" % len(out))

    vectors = parse_simple_indel_vectors()
    kept, reproduced = [], 0
    for v in vectors:
        setup = indel_batch(v)
        if setup is None:
            continue
        genome, reads, batch, cfg = setup
        r = oracle_lib.build_fragments(ref, oracle_lib.GenomeHolder(genome), reads, cfg, batch)
        frs = []
        for i in range(r.fragments.size):
            f = r.fragments[i]
            frs.append({k: (float(f[k]) if k == "logProbability" else int(f[k])) for k in f.dtype.names if k != "cigarOffset"})
            frs[-1]["cigar"] = [int(w) for w in r.cigar(i)]
            frs[-1]["cigarString"] = cigar_to_string(r.cigar(i))
            frs[-1]["logProbabilityBits"] = int(np.float64(f["logProbability"]).view(np.uint64))
        hit = [f for f in frs if f["cigarString"] == v["cigar"] and (v["mismatches"] is None or f["mismatchCount"] == v["mismatches"])
               and (v["editDistance"] is None or f["editDistance"] == v["editDistance"])
               and (v["position"] is None or f["position"] == v["position"])]
        v["reproducedThroughBuild"] = bool(hit)
        reproduced += bool(hit)
        v["referenceFragments"] = frs
        kept.append(v)
    json.dump({"source": "lib/alignment/cppunit/testSimpleIndelAligner.cpp", "scores": INDEL_SCORES, "vectors": kept},
              open(os.path.join(HERE, "simple_indels.json"), "w"), indent=1)
    print("simple_indels.json: %d vectors parsed, %d usable, %d reproduce the unit test's literal expectation through build()"
          % (len(vectors), len(kept), reproduced))


TEMPLATE_CASES = [("default", dict(scatter_repeats=False, dodgy=0, mapq_threshold=0)),
                  ("scatter_unknown", dict(scatter_repeats=True, dodgy=255, mapq_threshold=0)),
                  ("unaligned_mapq10", dict(scatter_repeats=False, dodgy=-1, mapq_threshold=10))]


def template_workload():
    """shared with tests/test_reference_goldens.py"""
    import hashlib
    from common_build import build_workload
    genome, sim, reads, mb = build_workload(n_pairs=200, L=100, seed=901, genome_bases=60_000, n_contigs=2, indel_rate=6e-3,
                                            neighbor_rate=0.4, repeat_rate=0.03)
    h = hashlib.sha256()
    for c in genome:
        h.update(np.ascontiguousarray(c).tobytes())
    h.update(np.ascontiguousarray(reads.bcl).tobytes())
    h.update(mb.matches.tobytes()); h.update(mb.begin.tobytes()); h.update(mb.seeds.tobytes())
    return genome, reads, mb, h.hexdigest()


TEMPLATE_FIELDS = ["alignmentScore", "fragmentAlignmentScore0", "fragmentAlignmentScore1", "properPair", "built", "hadFragments"]
TEMPLATE_READ_FIELDS = ["contigId", "position", "reverse", "observedLength", "mismatchCount", "editDistance", "smithWatermanScore",
                        "logProbabilityBits", "cigar"]


def templates_as_json(t):
    """one row per cluster: TEMPLATE_FIELDS then TEMPLATE_READ_FIELDS of read 1 and of read 2"""
    out = []
    for c in range(len(t.templates)):
        T = t.templates[c]
        row = [int(T["alignmentScore"]), int(T["fragmentAlignmentScore"][0]), int(T["fragmentAlignmentScore"][1]),
               int(T["properPair"]), int(T["built"]), int(T["hadFragments"])]
        for r in range(2):
            f = t.fragments[2 * c + r]
            row += [int(f["contigId"]), int(f["position"]), int(f["reverse"]), int(f["observedLength"]), int(f["mismatchCount"]),
                    int(f["editDistance"]), int(f["smithWatermanScore"]), int(np.float64(f["logProbability"]).view(np.uint64)),
                    cigar_to_string(t.cigar(2 * c + r))]
        out.append(row)
    return out


def make_templates(ref):
    from isaac_aligner_b200.batch import Tls, TemplateOptions
    genome, reads, mb, digest = template_workload()
    cfg = Config.default((0, -3, -11, -4, -20), max_read_length=200)
    cases = []
    for name, kw in TEMPLATE_CASES:
        t = oracle_lib.build_templates(ref, oracle_lib.GenomeHolder(genome), reads, cfg, mb, Tls.make(), TemplateOptions.make(**kw))
        cases.append({"name": name, "options": kw, "templates": templates_as_json(t)})
    json.dump({"source": "lib/alignment/TemplateBuilder.cpp via oracle/ref_capi.cpp:oracle_build_templates", "inputSha256": digest,
               "templateFields": TEMPLATE_FIELDS, "readFields": TEMPLATE_READ_FIELDS, "cases": cases},
              open(os.path.join(HERE, "templates.json"), "w"), separators=(",", ":"))
    built = sum(x[4] for x in cases[0]["templates"])
    print("templates.json: %d clusters x %d option sets, %d templates built in the default case" % (len(cases[0]["templates"]), len(cases), built))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "templates":
        make_templates(oracle_lib.reference())
    else:
        main()
        make_templates(oracle_lib.reference())
