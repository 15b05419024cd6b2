"""Golden vectors of isaac_ext_pack_fragments: the records the reference's own io::FragmentHeader constructors
(/root/reference/src/c++/include/io/Fragment.hh:100-186) + FragmentCollector::storeBclAndCigar leave for small tiles of random
template records, written to tests/golden/pack_fragments.json together with their inputs.  Run in the build container only (it
needs oracle/_ref, i.e. /root/reference at build time); the JSON is what the tests use where the reference build is absent.

    python tests/golden/make_pack_goldens.py
"""
import base64
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib                                                       # noqa: E402
import test_tile_write_bin_records as T                                    # noqa: E402
from isaac_aligner_b200.types import ReadSet                            # noqa: E402

OUT = os.path.join(HERE, "pack_fragments.json")
CASES = [((40, 40), True, True), ((40, 40), False, True), ((57, 36), True, False), ((64,), True, True), ((33,), False, False)]


def b64(a):
    return base64.b64encode(np.ascontiguousarray(a).tobytes()).decode()


def main():
    ref = oracle_lib.reference()
    cases = []
    for k, (read_lengths, keep, arrays) in enumerate(CASES):
        rng = np.random.default_rng(7000 + k)
        n = 24
        reads = ReadSet(T.random_bcl(rng, n, sum(read_lengths)), read_lengths)
        templates = T.random_templates(rng, n, read_lengths)
        options, barcode = T.make_options(rng, n, keep, with_arrays=arrays)
        want, mask = oracle_lib.pack_fragments(ref, reads, templates, options, barcode_bytes=barcode)
        for offset in want.read_offset[:len(read_lengths)]:          # the padding bytes of the reference's struct are stack garbage
            want.records[:, offset:offset + want.header_length] &= mask
        case = {"readLengths": list(read_lengths), "clusters": n, "keepUnaligned": keep, "tile": int(options.c.tile),
                "barcodeIdx": int(options.c.barcodeIdx), "bcl": b64(reads.bcl), "templates": b64(templates.templates),
                "fragments": b64(templates.fragments), "cigars": b64(templates.cigars),
                "pf": b64(options.pf) if options.pf is not None else None, "xy": b64(options.xy) if options.xy is not None else None,
                "barcodeSequence": b64(options.barcode_sequence) if options.barcode_sequence is not None else None,
                "distributionBinSize": int(options.c.distributionBinSize),
                "binIndex": [b.tolist() for b in T.bin_index()] if options.c.distributionBinSize else None,
                "recordLength": want.record_length, "readOffset": list(want.read_offset), "headerLength": want.header_length,
                "headerMask": b64(mask), "records": b64(want.records), "fStrandPos": b64(want.f_strand_pos),
                "initialized": b64(want.initialized)}
        cases.append(case)
    json.dump({"source": "io::FragmentHeader / FragmentCollector::add of the reference through oracle/_ref (oracle_pack_fragments)",
               "cases": cases}, open(OUT, "w"), indent=0)
    print("%s: %d cases, %d bytes" % (OUT, len(cases), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
