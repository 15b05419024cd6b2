"""K3, the warp-wavefront banded Smith-Waterman (csrc/sw_wide.cuh) through isaac_ext_banded_sw_wide_batch, bit-exact with its checker:
at 16 lanes the reference's own BandedSmithWaterman (oracle/_ref) and the restatement, at 32 lanes (BASELINE configs[4]: 2x250 bp with
the widened band, no reference behaviour) the band-width-parametrised model of the restatement, whose 16-lane instance is pinned
against the reference in tests/test_wide_band_oracle.py."""
import numpy as np
import pytest

import oracle_lib
from common import random_sw_cases
from isaac_aligner_b200.types import Config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


@pytest.mark.parametrize("scores", [(0, -3, 11, 4), (2, -1, 15, 3)])
def test_sixteen_lanes_equal_the_reference(capi, scores):
    queries, dbs = random_sw_cases(20001, seed=701 + scores[0])           # odd: the last warp holds one alignment, not two
    ctx = capi.Context(Config.default(max_read_length=300))
    cg, lg, og = ctx.banded_sw_wide(16, queries, dbs, scores)
    for chk in oracle_lib.gpu_checkers():
        cr, lr, orf = chk.banded_sw(queries, dbs, scores, max_read_length=300, threads=8)
        assert np.array_equal(lg, lr) and np.array_equal(og, orf) and np.array_equal(cg, cr), chk.kind
    # ... and the thread-per-alignment kernel of the main path
    c2, l2, o2 = ctx.banded_sw(queries, dbs, scores)
    assert np.array_equal(lg, l2) and np.array_equal(og, o2) and np.array_equal(cg, c2)
    ctx.close()


@pytest.mark.parametrize("scores,lmin,lmax,seed", [((0, -3, 11, 4), 30, 250, 711), ((2, -1, 15, 3), 30, 250, 712), ((0, -3, 11, 4), 240, 500, 713)])
def test_thirty_two_lanes_equal_the_wide_band_model(capi, scores, lmin, lmax, seed):
    queries, dbs = random_sw_cases(12000, seed=seed, lmin=lmin, lmax=lmax, band=32, max_indel=20)
    ctx = capi.Context(Config.default(max_read_length=600))
    cg, lg, og = ctx.banded_sw_wide(32, queries, dbs, scores)
    cr, lr, orf = oracle_lib.port().banded_sw(queries, dbs, scores, max_read_length=600, threads=8, band=32)
    assert np.array_equal(lg, lr) and np.array_equal(og, orf) and np.array_equal(cg, cr)
    gaps = ((cg & 0xF) == 2) & (np.arange(cg.shape[1])[None, :] < lg[:, None])
    assert ((cg >> 4) * gaps).max() > 16                                   # deletions no 16-lane band can hold were found
    ctx.close()


def test_wide_band_errors_are_loud(capi):
    ctx = capi.Context(Config.default(max_read_length=300))
    queries, dbs = random_sw_cases(10, seed=3, band=32)
    with pytest.raises(capi.ExtError) as e:
        ctx.banded_sw_wide(24, queries, dbs, (0, -3, 11, 4))
    assert e.value.code == 4                                               # ISAAC_EXT_E_UNSUPPORTED
    with pytest.raises(capi.ExtError):
        ctx.banded_sw_wide(32, [q.replace(b"A", b"a") for q in queries], dbs, (0, -3, 11, 4))
    ctx.close()
