"""Transcribes the literal vectors of the reference's own semialigned-clipper unit test
(/root/reference/src/c++/lib/alignment/cppunit/testSemialignedClipper.cpp:207-265) into tests/golden/semialigned_clipper.json:
for every test method the read and reference strings passed to align() (leading blanks of the reference = the read starts that
many bases in front of the contig, makeContig :162-171), the qualities of the harness (:95), and the values the test asserts
(CIGAR string and strand position after UngappedAligner::alignUngapped + SemialignedEndsClipper::clip, ELAND scores).
Run in the build container only (it reads /root/reference); the JSON is what the tests use."""
import json
import os
import re

SRC = "/root/reference/src/c++/lib/alignment/cppunit/testSemialignedClipper.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "semialigned_clipper.json")


def main():
    text = open(SRC).read()
    qualities = re.search(r'irrelevantQualities\("([^"]+)"\)', text).group(1)
    cases = []
    for m in re.finditer(r"void TestSemialignedClipper::(test\w+)\(\)\s*\{(.*?)\n\}", text, re.S):
        name, body = m.group(1), m.group(2)
        if name == "testEverything":
            continue
        code = "\n".join(line for line in body.split("\n") if not line.strip().startswith("//"))
        call = re.search(r'align\(\s*"([ACGTNn]+)"\s*,\s*"( *[ACGTNn]+)"', code, re.S)
        cigar = re.search(r'std::string\("(\w+)"\), fragmentMetadata.getCigarString', code).group(1)
        position = int(re.search(r"ReferencePosition\(0, (\d+)U?\), fragmentMetadata.getStrandReferencePosition", code).group(1))
        reverse = re.search(r"fragmentMetadata\.reverse\s*=\s*(true|false)", code).group(1) == "true"
        cases.append({"name": name, "reverse": reverse, "read": call.group(1), "reference": call.group(2), "cigar": cigar, "position": position})
    json.dump({"source": "testSemialignedClipper.cpp:207-265", "scores": [2, -1, -15, -3, -25], "qualities": qualities, "cases": cases},
              open(OUT, "w"), indent=1)
    print("%s: %d cases" % (OUT, len(cases)))


if __name__ == "__main__":
    main()
