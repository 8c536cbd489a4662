"""BN enumeration depth sweep: pedigrees of 12..17 members (3^17 = 1.3e8 configurations per variant) exercise 3..8
rolled odometer levels, dependent and independent unrolled blocks.  The exact ES posterior of the same (loop-free)
pedigree, itself bit-identical to the reference, is the yardstick: BN must agree within 1e-9 relative."""
import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O
from tests.util import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def grow(n):
    """ped14 plus extra members: grandchildren marry founders and have children (keeps the pedigree loop-free)."""
    ped = synth.ped14()
    rows = list(zip(ped.ids, ped.mids, ped.fids, ped.genders))
    extra = [(15, 0, 0, 2), (16, 15, 11, 1), (17, 15, 11, 2)]  # 11 (male grandchild) x founder 15 -> 16, 17
    rows += extra[: max(0, n - 14)]
    return synth._mk(rows[:n])


@pytest.mark.parametrize("n,V", [(12, 24), (13, 24), (15, 12), (16, 8), (17, 4)])
def test_bn_matches_es(n, V):
    ped = grow(n)
    lk, fl = synth.synth_likelihoods(ped, V, seed=400 + n, x_fraction=0.25)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=0) as e:
        info = e.info()
        es = e.run(fs.ES, lk, fl)
        bn = e.run(fs.BN, lk, fl)
    assert info["bn_levels"] == n and info["has_loop"] == 0
    want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.ES)
    assert np.array_equal(es.post, want["post"])
    assert np.array_equal(bn.status, es.status) and np.array_equal(bn.gt, es.gt)
    assert rel_err(bn.post, es.post) < REL_TOL and np.array_equal(bn.single, es.single)


def test_bn_generic_block_equals_register_block(monkeypatch):
    """The dependent-rows unrolled block and the register-resident one must agree on a pedigree that allows both."""
    ped = synth.ped14()
    lk, fl = synth.synth_likelihoods(ped, 40, seed=77, x_fraction=0.25)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=0) as e:
        a = e.run(fs.BN, lk, fl)
    monkeypatch.setenv("FAMSEQ_BN_GENERIC", "1")
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=0) as e:
        b = e.run(fs.BN, lk, fl)
    assert rel_err(a.post, b.post) < 1e-12 and np.array_equal(a.gt, b.gt)
