"""Transcribes the literal vectors of the reference's second FragmentBuilder unit test
(/root/reference/src/c++/lib/alignment/cppunit/testFragmentBuilder2.cpp:205-307: UngappedAligner / GappedAligner called directly,
ELAND scores) into tests/golden/fragment_builder2.json: strand, start position, read and reference strings, whether the gapped
aligner runs, and the values the test asserts (CIGAR string, mismatch count, edit distance, observed length, position, first
mismatch cycle).  Reverse-strand reads are given in strand order by the harness (:110-123 reverses, never complements); the tests
feed the reverse complement as the sequenced read so that the strand sequence is the same string.  Run in the build container only."""
import json
import os
import re

SRC = "/root/reference/src/c++/lib/alignment/cppunit/testFragmentBuilder2.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fragment_builder2.json")


def main():
    text = open(SRC).read()
    qualities = re.search(r'irrelevantQualities\("([^"]+)"\)', text).group(1)
    cases = []
    for m in re.finditer(r"void TestFragmentBuilder2::(test\w+)\(\)\s*\{(.*?)\n\}", text, re.S):
        name, body = m.group(1), m.group(2)
        if name == "testEverything":
            continue
        code = "\n".join(line for line in body.split("\n") if not line.strip().startswith("//"))
        call = re.search(r'align\(\s*"([ACGTNn]+)"\s*,\s*"([ACGTNn]+)"\s*,\s*noAdapters,\s*fragmentMetadata(, true)?\)', code, re.S)
        start = re.search(r"fragmentMetadata\.position = (-?\d+);", code)
        first = re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U, unsigned\(\*fragmentMetadata.getMismatchCyclesBegin", code)
        cases.append({
            "name": name, "reverse": re.search(r"fragmentMetadata\.reverse\s*=\s*(true|false)", code).group(1) == "true",
            "startPosition": int(start.group(1)) if start else 0, "read": call.group(1), "reference": call.group(2),
            "gapped": call.group(3) is not None,
            "cigar": re.search(r'std::string\("(\w+)"\), fragmentMetadata.getCigarString', code).group(1),
            "mismatchCount": int(re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U, fragmentMetadata.getMismatchCount", code).group(1)),
            "editDistance": int(re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U, fragmentMetadata.getEditDistance", code).group(1)),
            "observedLength": int(re.search(r"CPPUNIT_ASSERT_EQUAL\((\d+)U, fragmentMetadata.getObservedLength", code).group(1)),
            "position": int(re.search(r"ReferencePosition\(0, (\d+)U?\), fragmentMetadata.get\w*StrandReferencePosition", code).group(1)),
            "firstMismatchCycle": int(first.group(1)) if first else None})
    json.dump({"source": "testFragmentBuilder2.cpp:205-307", "scores": [2, -1, -15, -3, -25], "qualities": qualities, "firstCycle": 1,
               "cases": cases}, open(OUT, "w"), indent=1)
    print("%s: %d cases" % (OUT, len(cases)))


if __name__ == "__main__":
    main()
