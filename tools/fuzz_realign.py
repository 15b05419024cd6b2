"""Fuzz campaign for the gap realigner restatement (csrc/realign_device.cuh on the CPU, tests/cpp/test_realign_host.cu) against the
reference's own build::GapRealigner (oracle/_ref): random bins over random option sets.  usage: fuzz_realign.py [rounds] [seed0]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib                                    # noqa: E402
import test_realign_host as t                        # noqa: E402
from isaac_aligner_b200 import bins                  # noqa: E402
from isaac_aligner_b200.batch import Tls             # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    ref = oracle_lib.reference()
    lib = ctypes.CDLL(os.path.join(ROOT, "build", "libtest_realign_host.so"))
    total = 0
    for r in range(rounds):
        seed = seed0 + r
        rng = np.random.default_rng(seed)
        contigs = t.make_contigs(seed, lengths=(int(rng.integers(500, 4000)), int(rng.integers(20000, 50000))))
        contig = int(rng.integers(0, 2)) if contigs[0].size > 2500 else 1
        n = contigs[contig].size
        a = int(rng.integers(0, max(1, n // 4)))
        b = int(rng.integers(n // 2, n + 200))                                  # may lie behind the end of the contig
        spacing = int(rng.choice([25, 60, 120, 220, 400]))
        L = int(rng.choice([36, 75, 100, 150, 250, 400, 500]))
        barcodes = int(rng.integers(1, 4))
        bin_ = bins.simulate_bin(contigs, contig=contig, region=(a, min(b, n)), n_pairs=int(rng.integers(300, 1500)), read_length=L, seed=seed,
                                 variant_spacing=spacing, gapped_fraction=float(rng.choice([0.3, 0.6, 0.9])), barcodes=barcodes,
                                 clip_fraction=float(rng.choice([0.0, 0.1, 0.4, 0.9])), singleton_fraction=float(rng.choice([0.03, 0.3])),
                                 single_ended_fraction=float(rng.choice([0.03, 0.5])), dodgy_fraction=float(rng.choice([0.03, 0.4])), n_fraction=float(rng.choice([0.002, 0.05])), edge_fraction=float(rng.choice([0.0, 0.05])),
                                 max_indel=int(rng.choice([3, 14, 40])), template_mean=int(2.6 * L) + 60, error_rate=float(rng.choice([0.0, 0.004, 0.02])))
        bin_.bin_end = bins.reference_position(contig, b)
        tls = [Tls.make(mn=int(2.0 * L), mx=int(3.4 * L) + 120, median=int(2.6 * L) + 60) for _ in range(barcodes)]
        options = bins.RealignOptions(bin_.bin_start, bin_.bin_end, tls, vigorous=bool(rng.integers(0, 2)), dodgy=bool(rng.integers(0, 2)),
                                      clip_semialigned=bool(rng.integers(0, 2)), gap_groups=list(rng.integers(0, 2, size=barcodes)) if rng.integers(0, 2) else None,
                                      mismatch_cost=int(rng.choice([3, 3, 1, 5])), gap_open_cost=int(rng.choice([4, 4, 2, 8])), gap_extend_cost=int(rng.choice([0, 0, 1])))
        genome = oracle_lib.GenomeHolder(contigs)
        want = oracle_lib.realign_bin(ref, genome, bin_, options)
        got, counts = t.host_realign(lib, contigs, bin_, options)
        assert int(counts[4]) == 0, (seed, "error flags", int(counts[4]))
        total += t.compare(bin_, got, want)
        if (r + 1) % 10 == 0:
            print("%d rounds clean, %d realigned fragments so far" % (r + 1, total), flush=True)
    print("campaign %d: %d rounds, %d realigned fragments, no difference" % (seed0, rounds, total))


if __name__ == "__main__":
    main()
