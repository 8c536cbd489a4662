"""The oracle (oracle/famseq_oracle.c) is pinned against the reference.

* against every committed golden vector (tests/golden/*.npz, produced by the unmodified reference engine
  through oracle/_ref/ref_harness): bit-identical doubles for BN, ES and MCMC (libc rand stream);
* when oracle/_ref exists (build container, or shipped to the GPU box): fresh randomised comparisons.
"""
import os
import tempfile

import numpy as np
import pytest

from famseq_b200 import synth
from oracle import oracle as O
from tests.util import CasePed, golden_cases, load_case


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_golden_bit_for_bit(name):
    c = load_case(name)
    ped = CasePed(c)
    method = int(c["method"])
    r = O.run(ped, c["cols"], c["lk"], c["flags"], method=method, mrate=float(c["mrate"]), lc=float(c["lc"]),
              priors=c["priors"], burn=int(c["burn"]), rep=int(c["rep"]), rng=O.RNG_LIBC, seed=int(c["seed"]))
    assert np.array_equal(r["status"], c["status"])
    ok = c["status"] == 0
    for k in ("post", "single", "post_full", "single_full"):
        assert np.array_equal(r[k][ok], c[k][ok]), f"{name}: {k} is not bit-identical to the reference"
    assert np.array_equal(r["gt"][ok], c["gt"][ok])


def test_tables_match_closed_forms():
    a, xf, xm = O.tables(1e-7)
    t = a.reshape(3, 3, 3)
    assert np.allclose(t.sum(0), 1.0, atol=1e-15)
    assert abs(t[1, 1, 1] - 0.5) < 1e-15 and abs(t[0, 1, 1] - 0.25) < 1e-15
    assert np.allclose(xf.reshape(3, 3, 3).sum(0)[:, [0, 2]], 1.0, atol=1e-15)
    assert np.all(xf.reshape(3, 3, 3)[:, :, 1] == 0)  # a het father does not exist on X
    assert np.all(xm.reshape(3, 3, 3)[1] == 0)        # a het son does not exist on X
    a0, _, _ = O.tables(0.0)
    assert set(np.unique(a0)) <= {0.0, 0.25, 0.5, 1.0}


def test_loop_is_detected():
    p = synth.cousins_loop()
    lk, fl = synth.synth_likelihoods(p, 4, 1)
    with pytest.raises(RuntimeError):
        O.run(p, p.sequenced_cols(), lk, fl, method=O.ES)


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    import ctypes
    def px(ctr, key):
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); o = np.zeros(4, np.uint32)
        O.lib().fso_philox4x32(c.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p),
                               o.ctypes.data_as(ctypes.c_void_p))
        return [int(x) for x in o]
    assert px([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert px([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert px([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("pedname,method,V", [("trio", 2, 300), ("trio", 1, 300), ("half_sibs", 2, 100),
                                               ("three_wives", 2, 100), ("ped14", 2, 100), ("half_sibs", 3, 6),
                                               ("ped40", 3, 3), ("cousins_loop", 1, 20)])
def test_oracle_vs_live_reference(pedname, method, V):
    p = synth.PEDIGREES[pedname]()
    lk, fl = synth.synth_likelihoods(p, V, seed=1000 + V + method, x_fraction=0.3)
    kw = dict(method=method, burn=50, rep=400, seed=5)
    with tempfile.TemporaryDirectory() as td:
        pp = os.path.join(td, "p.ped")
        p.write(pp)
        ref = O.run_ref(pp, p.sequenced_cols(), lk, fl, **kw)
    got = O.run(p, p.sequenced_cols(), lk, fl, rng=O.RNG_LIBC, **kw)
    assert np.array_equal(got["status"], ref["status"])
    ok = ref["status"] == 0
    for k in ("post", "single", "gt", "post_full"):
        assert np.array_equal(got[k][ok], ref[k][ok]), k
