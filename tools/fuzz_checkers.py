"""Differential campaign between the two CPU checkers (the reference's own code in oracle/_ref and the scalar restatement):
random seeds, read lengths, score presets, adapter sets, --avoid-smith-waterman, quality masking.  Development aid, CPU only:
  python tools/fuzz_checkers.py [rounds [campaign seed]]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from common import assert_fragments_equal, small_workload  # noqa: E402
from common_build import assert_flat_equal, build_workload, rescue_requests  # noqa: E402
from isaac_aligner_b200 import synth  # noqa: E402
from isaac_aligner_b200.batch import FRm, FRp, RFm, RFp, Tls  # noqa: E402
from isaac_aligner_b200.types import (BWA_SCORES, ELAND_SCORES, NEXTERA_MATEPAIR_ADAPTERS, NEXTERA_STANDARD_ADAPTERS,  # noqa: E402
                                      STANDARD_ADAPTERS, Config, ReadSet)

SETS = [((), None, True), (STANDARD_ADAPTERS, ["AGATCGGAAGAGC"], True), (NEXTERA_STANDARD_ADAPTERS, ["CTGTCTCTTATACACATCT"], True),
        (NEXTERA_MATEPAIR_ADAPTERS, ["CTGTCTCTTATACACATCT", "AGATGTGTATAAGAGACAG", "CTGTCTCTTATACACATCTAGATGTGTATAAGAGACAG"], False)]


def main(rounds, campaign=20261017):
    ref, port = oracle_lib.reference(), oracle_lib.port()
    rng = np.random.default_rng(campaign)
    for k in range(rounds):
        seed = int(rng.integers(1, 1 << 30))
        L = int(rng.choice([36, 50, 75, 100, 150, 250]))
        scores = BWA_SCORES if rng.random() < 0.6 else ELAND_SCORES
        adapters, inserted, read_through = SETS[int(rng.integers(0, len(SETS)))]
        avoid = bool(rng.random() < 0.4)
        indel = float(rng.choice([5e-4, 4e-3, 1e-2]))
        cfg = Config.default(scores, max_read_length=2 * L, avoid_smith_waterman=avoid)
        cfg.repeatThreshold = int(rng.choice([2, 10, 10, 100]))
        cfg.gappedMismatchesMax = int(rng.choice([3, 5, 5, 8]))
        cfg.semialignedGapLimit = int(rng.choice([0, 100, 100, 20000]))
        what = "seed %d L %d scores %s adapters %d avoid %s indel %g repeat %d gmm %d sgl %d" % (
            seed, L, scores[0], len(adapters), avoid, indel, cfg.repeatThreshold, cfg.gappedMismatchesMax, cfg.semialignedGapLimit)
        genome, sim, reads, mb = build_workload(n_pairs=500, L=L, seed=seed, indel_rate=indel)
        if inserted and L >= 75:
            synth.insert_adapters(sim, inserted, fraction=0.4, seed=seed + 5, read_through=read_through, min_keep=min(40, L // 2))
            reads = ReadSet(sim.bcl, (L, L), end_cycles_masked=reads.end_cycles_masked)
        else:
            adapters = ()
        g = oracle_lib.GenomeHolder(genome)
        try:
            for chk in (ref, port):
                chk.set_adapters(adapters)
            assert_flat_equal(oracle_lib.build_fragments(ref, g, reads, cfg, mb), oracle_lib.build_fragments(port, g, reads, cfg, mb), "build " + what)
            models = [(FRp, RFm), (RFp, FRm)][int(rng.integers(0, 2))]
            tls = Tls.make(m0=models[0], m1=models[1])
            req = rescue_requests(sim, seed=seed + 9)
            assert_flat_equal(oracle_lib.rescue_shadows(ref, g, reads, cfg, tls, req), oracle_lib.rescue_shadows(port, g, reads, cfg, tls, req), "rescue " + what)
            r = oracle_lib.determine_template_length(ref, g, reads, cfg, mb, None, -1)
            p = oracle_lib.determine_template_length(port, g, reads, cfg, mb, None, -1)
            assert bytes(r[0]) == bytes(p[0]) and r[1] == p[1], "tls " + what
        finally:
            for chk in (ref, port):
                chk.set_adapters(())
        print("ok", k, what, flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 20261017)
