"""Golden vectors of build::GapRealigner: small synthetic bins (isaac_aligner_b200/bins.py) and what the reference's own classes make of
them (oracle/_ref, i.e. /root/reference compiled unmodified): Index::pos_ and CIGAR of every index entry, the records afterwards, the
two gap lists.  Needs /root/reference at build time of the checker; run from the repo root:
    python tests/golden/make_realign_goldens.py
The replay (tests/test_realign_goldens.py) needs neither the reference nor a GPU for the CPU part."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib                                    # noqa: E402
from isaac_aligner_b200 import bins                  # noqa: E402
from isaac_aligner_b200.batch import Tls             # noqa: E402
from test_realign_host import make_contigs           # noqa: E402

CASES = [  # name, seed, read length, variant spacing, vigorous, dodgy, clip, groups
    ("plain", 7, 100, 220, False, False, False, None),
    ("clip_groups", 8, 75, 90, False, True, True, [0, 1]),
    ("vigorous_dense", 9, 150, 35, True, False, True, None),
]


def main():
    ref = oracle_lib.reference()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    for name, seed, L, spacing, vigorous, dodgy, clip, groups in CASES:
        contigs = make_contigs(seed, lengths=(2500, 16000), n_runs=True)
        b = bins.simulate_bin(contigs, contig=1, region=(1200, 12000), n_pairs=450, read_length=L, seed=seed, variant_spacing=spacing,
                              barcodes=2 if groups else 1, template_mean=int(2.6 * L) + 60, clip_fraction=0.15)
        tls = [Tls.make(mn=int(2.0 * L), mx=int(3.4 * L) + 120, median=int(2.6 * L) + 60) for _ in range(2 if groups else 1)]
        o = bins.RealignOptions(b.bin_start, b.bin_end, tls, vigorous=vigorous, dodgy=dodgy, clip_semialigned=clip, gap_groups=groups)
        want = oracle_lib.realign_bin(ref, oracle_lib.GenomeHolder(contigs), b, o)
        # the realigned CIGARs per entry, back to back in index order, so that the comparison does not depend on pool offsets
        realigned = np.flatnonzero(want.cigar_offset != bins.OWN_CIGAR)
        cigars = np.concatenate([want.cigar(i, b) for i in realigned]) if realigned.size else np.zeros(0, np.uint32)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "realign_%s.npz" % name),
                            contig0=contigs[0], contig1=contigs[1], data=b.data, record_offset=b.record_offset, index=b.index,
                            bin=np.array([b.bin_start, b.bin_end], dtype=np.uint64), tls=np.frombuffer(bytes(o.tls), dtype=np.uint32),
                            flags=np.array([vigorous, dodgy, clip], dtype=np.uint8), groups=np.array(groups if groups else [], dtype=np.uint32),
                            position=want.position, realigned=realigned, cigar_length=want.cigar_length, cigars=cigars,
                            data_after=want.data, gaps=want.gaps, deletions=want.deletions)
        print(name, "entries", len(b.index), "realigned", realigned.size, "gaps", want.gaps.size)


if __name__ == "__main__":
    main()
