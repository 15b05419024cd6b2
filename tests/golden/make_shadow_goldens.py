"""Transcribes the literal vectors of the reference's ShadowAligner unit test
(/root/reference/src/c++/lib/alignment/cppunit/testShadowAligner.cpp:55-300) into tests/golden/shadow_aligner.json: per block the
template length statistics, the getBcl() arguments that cut the two reads (81 / 92 bases, Q40) out of a contig
(BuilderInit.hh:143-172), the orphan the test hands to rescueShadow and the values it asserts on shadowList[0]; the second call
of a block takes the first call's shadow as its orphan.  The fixture's contigs are rand() noise of lengths 190, 300, 230, 235
(= "AAAAA" + c2), 422 (getContigList(190, 300, 422), BuilderInit.hh:121-131): the asserted values depend on the geometry only, the
tests generate noise of the same lengths from a fixed seed.  Run in the build container only (it reads /root/reference)."""
import json
import os
import re

SRC = "/root/reference/src/c++/lib/alignment/cppunit/testShadowAligner.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shadow_aligner.json")
MODELS = {"FFp": 0, "FRp": 1, "RFp": 2, "RRp": 3, "FFm": 4, "FRm": 5, "RFm": 6, "RRm": 7}


def main():
    text = open(SRC).read()
    cases = []
    for m in re.finditer(r"void TestShadowAligner::(testRescueShadow\w+)\(\)\s*\{(.*?)\n\}", text, re.S):
        name, body = m.group(1), m.group(2)
        blocks = body.split("const TemplateLengthStatistics tls(")[1:]
        for k, block in enumerate(blocks):
            tls = re.match(r"(\d+), (\d+), (\d+), (\d+), (\d+), TemplateLengthStatistics::(\w+), TemplateLengthStatistics::(\w+), (-?\d+)\)", block)
            bcl = re.search(r"getBcl\(readMetadataList, contigList, (\d+), (\d+), (\d+), (true|false), (true|false)\)", block)
            orphan = {"readIndex": int(re.search(r"fragment0\.readIndex = (\d+);", block).group(1)),
                      "contigId": int(re.search(r"fragment0\.contigId = (\d+);", block).group(1)),
                      "position": int(re.search(r"fragment0\.position = (\d+);", block).group(1)),
                      "reverse": re.search(r"fragment0\.reverse = (true|false);", block).group(1) == "true",
                      "observedLength": 0}                               # a default-constructed FragmentMetadata
            expects = []
            for call in block.split("shadowAligner.rescueShadow(")[1:]:
                who = re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)L, (fragment\d)\.position\)", call)
                f = who.group(2)
                expects.append({
                    "position": int(who.group(1)),
                    "reverse": re.search(r"CPPUNIT_ASSERT_EQUAL\((true|false), %s\.reverse\)" % f, call).group(1) == "true",
                    "observedLength": int(re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U, %s\.observedLength\)" % f, call).group(1)),
                    "mismatchCount": int(re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U, %s\.mismatchCount\)" % f, call).group(1)),
                    "cigarLength": int(re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U, %s\.cigarLength\)" % f, call).group(1)),
                    "cigarWord": int(re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U << 4, shadowAligner", call).group(1)) << 4,
                    "logProbability": float(re.search(r"CPPUNIT_ASSERT_DOUBLES_EQUAL\((-[\d.]+), %s\.logProbability, ([\d.]+)\)" % f, call).group(1)),
                    "tolerance": float(re.search(r"CPPUNIT_ASSERT_DOUBLES_EQUAL\((-[\d.]+), %s\.logProbability, ([\d.]+)\)" % f, call).group(2)),
                })
            assert len(expects) == 2, (name, k, len(expects))
            cases.append({"name": "%s[%d]" % (name, k),
                          "tls": {"min": int(tls.group(1)), "max": int(tls.group(2)), "median": int(tls.group(3)), "lowStdDev": int(tls.group(4)),
                                  "highStdDev": int(tls.group(5)), "bestModel": [MODELS[tls.group(6)], MODELS[tls.group(7)]],
                                  "mateDriftRange": int(tls.group(8))},
                          "bcl": {"contigId": int(bcl.group(1)), "offset0": int(bcl.group(2)), "offset1": int(bcl.group(3)),
                                  "reverse0": bcl.group(4) == "true", "reverse1": bcl.group(5) == "true"},
                          "orphan": orphan, "expect": expects})
    json.dump({"source": "testShadowAligner.cpp:55-300, BuilderInit.hh:121-172", "scores": [2, -1, -15, -3, -25], "readLengths": [81, 92],
               "contigLengths": [190, 300, 230, 235, 422], "quality": 40, "gappedMismatchesMax": 8, "cases": cases}, open(OUT, "w"), indent=1)
    print("%s: %d blocks" % (OUT, len(cases)))


if __name__ == "__main__":
    main()
