"""GPU tests of the compact OUTPUT (output half): posteriors Phred-encoded on the device as
the six decimal digits + exponent of the reference's own text (file.cpp:702-761: "%g" of fabs(-10*log10(p)), 99999 for
+inf).  Every code the device decides must give exactly the text the reference's formatting gives for the FP64 value
(oracle restatement: libm log10 + printf %g); whatever it does not decide must come back as an exception with the exact
double; and the exception rate must stay tiny."""
import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def engine_for(ped, device=0):
    return fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=device)


def check_codes(codes, p, fixes, n_fixes, what, single_fix_flag=False):
    """Every element: decided codes give the reference text; exception codes are listed in `fixes` with the exact double."""
    codes, p = codes.reshape(-1), p.reshape(-1)
    assert n_fixes == len(fixes), f"{what}: exception list overflowed ({n_fixes})"
    kind = codes >> 30
    is_fix = kind == 3
    idx = fixes["index"] & ~fs.engine.PHRED_FIX_SINGLE if single_fix_flag else fixes["index"]
    assert sorted(idx.tolist()) == np.nonzero(is_fix)[0].tolist(), f"{what}: exception list and exception codes disagree"
    for i, x in zip(idx, fixes["p"]):
        assert (np.isnan(x) and np.isnan(p[i])) or x == p[i], f"{what}: exception {i} carries {x!r}, value is {p[i]!r}"
    decided = np.nonzero(~is_fix)[0]
    # text of a code depends on the code only: format every distinct code once, every distinct probability once
    uc, inv_c = np.unique(codes[decided], return_inverse=True)
    text_c = np.array([fs.phred_text(int(c)) for c in uc], dtype=object)[inv_c]
    up, inv_p = np.unique(p[decided], return_inverse=True)
    text_p = np.array([O.phred_text(float(x)) for x in up], dtype=object)[inv_p]
    bad = np.nonzero(text_c != text_p)[0]
    assert bad.size == 0, f"{what}: {bad.size} codes print differently, e.g. p={p[decided][bad[0]]!r}: {text_c[bad[0]]} vs {text_p[bad[0]]}"
    return int(is_fix.sum())


def test_encoder_on_adversarial_probabilities():
    rng = np.random.default_rng(2026)
    n = 400_000
    six = rng.integers(100000, 1000000, n).astype(np.float64)
    sh = rng.integers(-21, -1, n)  # Phred values 1e-16 .. 1e4 with six given digits ...
    ties = (six + 0.5) * 10.0 ** sh  # ... followed by a 5: exact ties of the seventh digit and their neighbours
    near = np.concatenate([ties, np.nextafter(ties, 0), np.nextafter(ties, 1e9), six * 10.0 ** sh])
    near = near[(near > 5e-16) & (near < 3200)]
    p_near = 10.0 ** (-near / 10.0)  # probabilities whose Phred value lands next to a rounding boundary
    p = np.concatenate([rng.random(n), 10.0 ** (-320 * rng.random(n)), 1 - 10.0 ** (-16 * rng.random(n)), p_near[:n],
                        rng.random(1000) * 5e-320, [0.0, 1.0, -0.0, 5e-324, 2.2250738585072014e-308, 0.5, 0.1, 1e-5, np.nan, -0.5, 1.5, np.inf]])
    with engine_for(synth.trio()) as e:
        codes, fixes, n_fixes = e.phred_encode(p, fix_capacity=1 << 20)
    n_fix = check_codes(codes, p, fixes, n_fixes, "adversarial")
    assert (codes[-12:-4] >> 30).tolist()[:3] == [2, 1, 2]  # 0 -> 99999, 1 -> 0, -0 -> 99999
    assert ((codes[-4:] >> 30) == 3).all()                  # NaN, negative, > 1, inf: the host decides
    ordinary = codes[:3 * n]
    assert ((ordinary >> 30) == 3).mean() < 1e-5, "too many ordinary probabilities go to the host"
    assert n_fix < 0.2 * p.size


@pytest.mark.parametrize("pedname,method,V,kw", [("trio", fs.ES, 300_007, {}), ("ped14", fs.ES, 20_011, {}), ("half_sibs", fs.BN, 2001, {}),
                                                ("half_sibs", fs.MCMC, 3000, dict(burn=10, rep=60, seed=5, v_offset=11))])
def test_phred_entry_prints_what_the_fp64_entry_prints(pedname, method, V, kw):
    ped = synth.PEDIGREES[pedname]()
    pl, fl = synth.synth_pl(ped, V, seed=77 + V, x_fraction=0.25)
    pl = pl.astype(np.uint16)
    pl[::97, 0] = 65535  # impossible samples: failing variants
    with engine_for(ped) as e:
        want = e.run_pl(method, pl, fl, **kw)
        got = e.run_pl_phred(method, pl, fl, fix_capacity=1 << 16, **kw)
        lean = e.run_pl_phred(method, pl, fl, want_single=False, fix_capacity=1 << 16, **kw)
    assert np.array_equal(got.gt, want.gt) and np.array_equal(got.status, want.status) and want.status.sum() > 0
    post_fix = got.fixes[(got.fixes["index"] & fs.engine.PHRED_FIX_SINGLE) == 0]
    single_fix = got.fixes[(got.fixes["index"] & fs.engine.PHRED_FIX_SINGLE) != 0]
    assert got.n_fixes == len(got.fixes)
    n1 = check_codes(got.post, want.post, post_fix, len(post_fix), f"{pedname} post")
    n2 = check_codes(got.single, want.single, single_fix, len(single_fix), f"{pedname} single", single_fix_flag=True)
    assert n1 + n2 <= 1e-5 * got.post.size + 4
    assert lean.single is None and np.array_equal(lean.post, got.post) and np.array_equal(lean.gt, got.gt)
    assert lean.n_fixes == len(post_fix)


def test_exception_capacity_is_reported_not_exceeded():
    p = np.full(1000, np.nan)
    with engine_for(synth.trio()) as e:
        codes, fixes, n = e.phred_encode(p, fix_capacity=10)
        assert n == 1000 and len(fixes) == 10 and ((codes >> 30) == 3).all()
        codes, fixes, n = e.phred_encode(p[:0])
        assert n == 0 and codes.size == 0


def test_phred_entry_on_a_multi_device_engine():
    devices = list(range(fs.device_count())) if fs.device_count() > 1 else [0, 0]
    ped = synth.trio()
    pl, fl = synth.synth_pl(ped, 50_003, seed=5)
    pl = pl.astype(np.uint16)
    bad = np.array([7, 20_000, 49_999])
    with engine_for(ped) as one, engine_for(ped, device=devices) as many:
        a = one.run_pl_phred(fs.ES, pl, fl, fix_capacity=4096)
        b = many.run_pl_phred(fs.ES, pl, fl, fix_capacity=4096)
    assert np.array_equal(a.post, b.post) and np.array_equal(a.single, b.single) and np.array_equal(a.gt, b.gt)
    assert a.n_fixes == b.n_fixes and sorted(a.fixes["index"].tolist()) == sorted(b.fixes["index"].tolist())
    assert bad.size == 3
