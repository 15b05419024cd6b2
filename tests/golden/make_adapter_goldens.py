"""Transcribes the literal vectors of the reference's own adapter unit test
(/root/reference/src/c++/lib/alignment/cppunit/testSequencingAdapter.cpp) into tests/golden/adapters.json: for every test
method the strand (fragmentMetadata.reverse), the read and reference strings passed to align(), the adapter list, and the
values the test asserts (CIGAR string, mismatch count, edit distance, observed length, position).  Run in the build
container only (it reads /root/reference); the JSON is what the tests use.

The reference's harness (testSequencingAdapter.cpp:159-182) aligns the read at position 0 of a one-contig reference with a
fresh FragmentSequencingAdapterClipper: checkInitStrand + UngappedAligner::alignUngapped (ELAND scores).  Its reverse-strand
reads are given in strand order (the harness reverses, never complements); the tests here feed the reverse complement as
the sequenced read so that the strand sequence is the same string."""
import json
import os
import re

SRC = "/root/reference/src/c++/lib/alignment/cppunit/testSequencingAdapter.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adapters.json")

ADAPTERS = {
    # testSequencingAdapter.cpp:59-77,96-99
    "matePairAdapters": [["CTGTCTCTTATACACATCT", False, 19], ["AGATGTGTATAAGAGACAG", True, 19]],
    "standardAdapters": [["CTGTCTCTTATACACATCT", False, 0], ["AGATGTGTATAAGAGACAG", True, 0]],
}


def main():
    text = open(SRC).read()
    cases = []
    for m in re.finditer(r"void TestSequencingAdapter::(test\w+)\(\)\s*\{(.*?)\n\}", text, re.S):
        name, body = m.group(1), m.group(2)
        if name == "testEverything":
            continue
        code = "\n".join(line for line in body.split("\n") if not line.strip().startswith("//"))
        code = re.sub(r"//[^\n]*", "", code)
        rev = re.search(r"fragmentMetadata\.reverse\s*=\s*(true|false)", code).group(1) == "true"
        call = re.search(r'align\(\s*"([ACGTNn]+)"\s*,\s*"([ACGTNn]+)"\s*,\s*(\w+)\s*,', code, re.S)
        read, reference, adapters = call.group(1), call.group(2), call.group(3)
        case = {"name": name, "reverse": rev, "read": read, "reference": reference, "adapters": ADAPTERS[adapters],
                "adapterList": adapters}
        c = re.search(r'std::string\("(\w+)"\), fragmentMetadata.getCigarString', code)
        case["cigar"] = c.group(1)
        for key, pat in (("mismatchCount", r"(\d+)U, fragmentMetadata.getMismatchCount"),
                         ("editDistance", r"(\d+)U, fragmentMetadata.getEditDistance"),
                         ("observedLength", r"(\d+)U, fragmentMetadata.getObservedLength"),
                         ("position", r"ReferencePosition\(0, (\d+)U\), fragmentMetadata.get\w*StrandReferencePosition")):
            v = re.search(pat, code)
            if v:
                case[key] = int(v.group(1))
        cases.append(case)
    assert len(cases) == 15, len(cases)
    json.dump({"source": "testSequencingAdapter.cpp (iSAAC-01.15.04.01)", "scores": [2, -1, -15, -3, -25], "cases": cases},
              open(OUT, "w"), indent=1)
    print("wrote %d cases to %s" % (len(cases), OUT))


if __name__ == "__main__":
    main()
