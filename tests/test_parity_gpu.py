"""GPU parity tests: the CUDA kernels, called through the C ABI, against
  (a) the committed golden vectors produced by the unmodified reference engine,
  (b) the C oracle on fresh seeded inputs,
  (c) size-independent properties at large batch sizes (ES == BN on loop-free pedigrees, shard invariance,
      rows summing to one, run-to-run determinism).
Tolerances: BN / ES posteriors 1e-9 relative (north_star), called genotypes and status bit-exact; ES is in
fact required to be bit-identical to the reference (the kernel mirrors its operation order without FMA).
MCMC: identical Philox stream in oracle and kernel => 1e-9 relative; plus a Monte-Carlo z-test against the
exact BN posterior."""
import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O
from tests.util import REL_TOL, CasePed, assert_parity, assert_same_chains, golden_cases, load_case, rel_err

pytestmark = pytest.mark.gpu

# the first n rows form a valid pedigree for every n (parents precede children; married-in founders 5, 8, 11)
NESTED = [(1, 0, 0, 1), (2, 0, 0, 2), (3, 2, 1, 1), (4, 2, 1, 2), (5, 0, 0, 2), (6, 5, 3, 1), (7, 5, 3, 2), (8, 0, 0, 1),
          (9, 4, 8, 2), (10, 4, 8, 1), (11, 0, 0, 2)]


def engine_for(ped, cols=None, params=None):
    cols = ped.sequenced_cols() if cols is None else cols
    return fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, params=params, device=0)


def params_from(c):
    prm = fs.Params.default()
    prm.mrate = float(c["mrate"])
    prm.lrc = float(c["lc"])
    for k, row in zip(("geno_prob_n", "geno_prob_k", "geno_prob_xn", "geno_prob_xk"), c["priors"]):
        for g in range(3):
            getattr(prm, k)[g] = float(row[g])
    return prm


# ------------------------------------------------------------------------------------------------------
# (a) golden vectors from the reference
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in golden_cases() if not n.endswith("_mcmc")])
def test_golden_bn_es(name):
    c = load_case(name)
    method = int(c["method"])
    with engine_for(CasePed(c), c["cols"].tolist(), params_from(c)) as e:
        got = e.run(method, c["lk"], c["flags"])
    assert_parity(got, c, REL_TOL, name)
    if method == fs.ES:  # stronger than the contract: same doubles as the reference
        ok = c["status"] == 0
        assert np.array_equal(got.post[ok], c["post"][ok]), f"{name}: ES is not bit-identical to the reference"
        assert np.array_equal(got.single[ok], c["single"][ok])


@pytest.mark.parametrize("name", golden_cases("*_mcmc"))
def test_golden_mcmc_statistical(name):
    """The golden MCMC vectors use libc rand(); only statistical agreement is meaningful.  Compare both the
    reference's single run and ours against each other with the spread of 16 of our seeds."""
    c = load_case(name)
    burn, rep = int(c["burn"]), int(c["rep"])
    with engine_for(CasePed(c), c["cols"].tolist(), params_from(c)) as e:
        runs = np.stack([e.run(fs.MCMC, c["lk"], c["flags"], burn=burn, rep=rep, seed=100 + k).post for k in range(16)])
        single = e.run(fs.MCMC, c["lk"], c["flags"], burn=burn, rep=rep, seed=1).single
    assert np.array_equal(single, c["single"])
    mean, sd = runs.mean(0), runs.std(0, ddof=1)
    # the reference's one run must look like a draw from our seed-to-seed distribution (mixing at mu=1e-7 is
    # slow, so the spread is dominated by the random start; z = 6 on every entry, absolute floor 1e-12)
    z = np.abs(c["post"] - mean) / (sd * np.sqrt(1 + 1 / 16) + 1e-12)
    assert np.quantile(z, 0.99) < 6.0, f"{name}: z99={np.quantile(z, 0.99):.2f}"


# ------------------------------------------------------------------------------------------------------
# (b) fresh inputs against the C oracle
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pedname,method,V,xf", [
    ("trio", fs.ES, 100003, 0.2), ("trio", fs.BN, 50001, 0.2), ("ped14", fs.ES, 5000, 0.2),
    ("half_sibs", fs.ES, 3000, 0.3), ("three_wives", fs.ES, 3000, 0.3), ("half_sibs", fs.BN, 40, 0.3),
    ("three_wives", fs.BN, 40, 0.3), ("cousins_loop", fs.BN, 100, 0.3), ("ped14", fs.BN, 6, 0.5),
    ("ped100", fs.ES, 600, 0.2),
])
def test_random_vs_oracle(pedname, method, V, xf):
    ped = synth.PEDIGREES[pedname]()
    lk, fl = synth.synth_likelihoods(ped, V, seed=777 + V, x_fraction=xf)
    want = O.run(ped, ped.sequenced_cols(), lk, fl, method=method)
    with engine_for(ped) as e:
        got = e.run(method, lk, fl)
    e1, e2 = assert_parity(got, want, REL_TOL, f"{pedname}/{method}")
    if method == fs.ES:
        ok = want["status"] == 0
        assert np.array_equal(got.post[ok], want["post"][ok])


@pytest.mark.parametrize("sizes", [(1, 2, 3), (4,), (5,), (6,), (7,), (8,), (9,), (10,), (11,)])
def test_bn_every_group_shape(sizes):
    """BN thread-group shapes change with N (1, 3, 9, 27, 81, 243 threads per variant, 0-2 rolled levels):
    cover every pedigree size from 1 to 11 with a chain of nuclear families."""
    for n in sizes:
        rows = NESTED[:n]
        ped = synth._mk(rows)
        V = 64 if n <= 9 else 12
        lk, fl = synth.synth_likelihoods(ped, V, seed=31 + n, x_fraction=0.3)
        want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.BN)
        with engine_for(ped) as e:
            got = e.run(fs.BN, lk, fl)
        assert_parity(got, want, REL_TOL, f"BN n={n}")


@pytest.mark.parametrize("pedname,V", [("ped14", 8), ("half_sibs", 60), ("three_wives", 60), ("cousins_loop", 60), ("trio", 500)])
def test_bn_with_analytically_summed_leaves(pedname, V, monkeypatch):
    """FAMSEQ_BN_FACTOR=1: the innermost block of childless members is summed in closed form instead of being enumerated
    (bn_kernel.cu, bn_block_factored).  Same marginals as the exhaustive enumeration and as the oracle, to 1e-9."""
    ped = synth.PEDIGREES[pedname]()
    lk, fl = synth.synth_likelihoods(ped, V, seed=41, x_fraction=0.3)
    want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.BN)
    with engine_for(ped) as e:
        plain = e.run(fs.BN, lk, fl)
    monkeypatch.setenv("FAMSEQ_BN_FACTOR", "1")
    with engine_for(ped) as e:
        got = e.run(fs.BN, lk, fl)
    assert_parity(got, want, REL_TOL, f"BN factored/{pedname}")
    assert np.array_equal(got.gt, plain.gt) and np.array_equal(got.status, plain.status)
    ok = plain.status == 0
    assert rel_err(got.post[ok], plain.post[ok]) <= REL_TOL


@pytest.mark.parametrize("pedname,cols,V", [("ped14", None, 20011), ("half_sibs", None, 5003), ("three_wives", None, 5000),
                                          ("ped14", [13, 2, 7, 0, 10, 5], 4001)])
def test_es_generated_kernel_returns_the_same_doubles(pedname, cols, V, monkeypatch):
    """FAMSEQ_ES_JIT=1: the message program as straight-line code (es_jit.cu) against the interpreter (es_kernel.cu) and
    the oracle: same doubles (autosomes, chrX, Known, failing and LRC-gated variants, a ragged last tile)."""
    ped = synth.PEDIGREES[pedname]()
    cols = ped.sequenced_cols() if cols is None else cols
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, len(cols) + 1)]), V, seed=91, x_fraction=0.3)
    lk[::7] = np.round(lk[::7])  # certain variants: the LRC gate keeps the pedigree out
    lk[5::11] = 0.0              # a failing variant every now and then
    monkeypatch.setenv("FAMSEQ_ES_JIT", "0")
    with engine_for(ped, cols) as e:
        interp = e.run(fs.ES, lk, fl)
        assert e.info()["jit_launches"] == 0
    monkeypatch.setenv("FAMSEQ_ES_JIT", "1")
    with engine_for(ped, cols) as e:
        got = e.run(fs.ES, lk, fl)
        assert e.info()["jit_launches"] >= 1
    assert np.array_equal(got.status, interp.status) and np.array_equal(got.gt, interp.gt)
    assert np.array_equal(got.single, interp.single, equal_nan=True) and np.array_equal(got.post, interp.post, equal_nan=True)
    want = O.run(ped, cols, lk, fl, method=O.ES)
    ok = want["status"] == 0
    assert np.array_equal(got.status, want["status"]) and np.array_equal(got.post[ok], want["post"][ok])


@pytest.mark.parametrize("V", [4096 + 17, 29])
def test_es_generated_kernel_when_a_row_sum_vanishes_inside_the_pedigree(V, monkeypatch):
    """-mRate 0 and Mendel-impossible certain genotypes: the individual-only posteriors are fine, the peeling hits a zero sum
    (family.cpp:1376-1384) and the variant fails AFTER the generated kernel has sent its `single` rows off -- the late clean-up
    of es_jit.cu, in full tiles and in the ragged one.  Same bytes as the interpreter and the oracle."""
    ped = synth.ped14()
    cols = ped.sequenced_cols()
    lk, fl = synth.synth_likelihoods(ped, V, seed=77, x_fraction=0.25)
    kid = next(i for i in range(ped.n) if ped.mids[i] != 0)
    mother, father = list(ped.ids).index(ped.mids[kid]), list(ped.ids).index(ped.fids[kid])
    bad = np.arange(V) % 5 == 3
    lk[np.ix_(bad, [mother, father])] = np.array([1.0, 0.0, 0.0])
    lk[bad, kid] = np.array([0.0, 0.0, 1.0])
    prm = fs.Params.default()
    prm.mrate = 0.0
    want = O.run(ped, cols, lk, fl, method=O.ES, mrate=0.0)
    assert want["status"][bad].all() and not want["status"][~bad].all()
    out = {}
    for jit in ("0", "1"):
        monkeypatch.setenv("FAMSEQ_ES_JIT", jit)
        with engine_for(ped, cols, prm) as e:
            out[jit] = e.run(fs.ES, lk, fl)
            assert (e.info()["jit_launches"] >= 1) == (jit == "1")
    for jit, got in out.items():
        assert np.array_equal(got.status, want["status"]), jit
        assert not got.post[got.status != 0].any() and not got.single[got.status != 0].any(), jit
        assert (got.gt[got.status != 0] == 255).all(), jit
        ok = want["status"] == 0
        assert np.array_equal(got.post[ok], want["post"][ok]) and np.array_equal(got.single[ok], want["single"][ok]), jit
    assert np.array_equal(out["0"].post, out["1"].post) and np.array_equal(out["0"].single, out["1"].single)


def test_partial_sequencing_and_column_order():
    """Input columns in a different order than the ped rows, some members unsequenced."""
    ped = synth.ped14()
    cols = [13, 2, 7, 0, 10, 5]
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, 7)]), 2000, seed=5, x_fraction=0.2)
    for method, V in ((fs.ES, 2000), (fs.BN, 6)):
        want = O.run(ped, cols, lk[:V], fl[:V], method=method)
        with engine_for(ped, cols) as e:
            got = e.run(method, lk[:V], fl[:V])
        assert_parity(got, want, REL_TOL, f"partial/{method}")


def test_lrc_and_priors_variants():
    ped = synth.half_sibs()
    lk, fl = synth.synth_likelihoods(ped, 1500, seed=9, x_fraction=0.3)
    pri = np.array([[0.98, 0.015, 0.005], [0.3, 0.4, 0.3], [0.99, 0.0, 0.01], [0.6, 0.0, 0.4]])
    for lc, mrate in ((0.0, 1e-7), (0.9999, 1e-7), (1.0, 0.0), (5.0, 1e-3)):
        prm = fs.Params.default()
        prm.lrc, prm.mrate = lc, mrate
        for k, row in zip(("geno_prob_n", "geno_prob_k", "geno_prob_xn", "geno_prob_xk"), pri):
            for g in range(3):
                getattr(prm, k)[g] = float(row[g])
        want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.ES, mrate=mrate, lc=lc, priors=pri)
        with engine_for(ped, params=prm) as e:
            got = e.run(fs.ES, lk, fl)
            gbn = e.run(fs.BN, lk[:30], fl[:30])
        assert_parity(got, want, REL_TOL, f"lrc={lc}")
        wbn = O.run(ped, ped.sequenced_cols(), lk[:30], fl[:30], method=O.BN, mrate=mrate, lc=lc, priors=pri)
        assert_parity(gbn, wbn, REL_TOL, f"bn lrc={lc}")


def test_es_refuses_loops_bn_and_mcmc_accept():
    ped = synth.cousins_loop()
    lk, fl = synth.synth_likelihoods(ped, 16, seed=2)
    with engine_for(ped) as e:
        with pytest.raises(fs.FamSeqError) as ei:
            e.run(fs.ES, lk, fl)
        assert ei.value.code == -4
        assert e.run(fs.BN, lk, fl).status.sum() == 0
        assert e.run(fs.MCMC, lk, fl, burn=10, rep=100).status.sum() == 0


def test_empty_and_ragged_batches():
    ped = synth.trio()
    lk, fl = synth.synth_likelihoods(ped, 1000, seed=4)
    want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.ES)
    with engine_for(ped) as e:
        r0 = e.run(fs.ES, lk[:0], fl[:0])
        assert r0.post.shape == (0, 3, 3)
        for V in (1, 2, 127, 128, 129, 999):
            got = e.run(fs.ES, lk[:V], fl[:V])
            assert np.array_equal(got.post, want["post"][:V]) and np.array_equal(got.gt, want["gt"][:V].astype(np.uint8))
            gb = e.run(fs.BN, lk[:V], fl[:V])
            assert rel_err(gb.post, want["post"][:V]) < REL_TOL
        # flags == NULL means "novel autosomal" for every variant (the LK driver, file.cpp:1751)
        w0 = O.run(ped, ped.sequenced_cols(), lk[:100], None, method=O.ES)
        assert np.array_equal(e.run(fs.ES, lk[:100], None).post, w0["post"])


# ------------------------------------------------------------------------------------------------------
# MCMC
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pedname,V,burn,rep", [("trio", 300, 50, 500), ("half_sibs", 100, 50, 400), ("ped40", 40, 30, 300),
                                                ("ped100", 40, 20, 150)])
def test_mcmc_same_stream_as_oracle(pedname, V, burn, rep):
    """Oracle and kernel draw from the same Philox stream: the chains visit the same states, so the
    Rao-Blackwellised posteriors agree to rounding (the kernel multiplies by 1/sum instead of dividing)."""
    ped = synth.PEDIGREES[pedname]()
    lk, fl = synth.synth_likelihoods(ped, V, seed=21, x_fraction=0.25)
    want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=12345, v_offset=1000)
    with engine_for(ped) as e:
        got = e.run(fs.MCMC, lk, fl, burn=burn, rep=rep, seed=12345, v_offset=1000)
        # shard invariance: the second half computed on its own, with its global offset, gives the same bytes
        half = e.run(fs.MCMC, lk[V // 2:], fl[V // 2:], burn=burn, rep=rep, seed=12345, v_offset=1000 + V // 2)
    assert_parity(got, want, 1e-9, f"mcmc/{pedname}")
    assert np.array_equal(half.post, got.post[V // 2:]) and np.array_equal(half.gt, got.gt[V // 2:])


@pytest.mark.parametrize("pedname,cols,V,burn,rep", [("trio", None, 700, 20, 300), ("half_sibs", None, 500, 20, 200),
                                                     ("three_wives", None, 300, 10, 100), ("ped14", [13, 2, 7, 0, 10, 5], 300, 10, 100),
                                                     ("ped40", None, 600, 10, 150), ("ped100", None, 300, 10, 60)])
def test_mcmc_specialised_kernel_returns_the_same_bytes(pedname, cols, V, burn, rep, monkeypatch):
    """The Gibbs kernel the engine generates and compiles for one pedigree (gibbs_jit.cu) against the table-driven
    kernel: same Philox stream, same weights, hence the same chains -- autosomes and chrX, Known or not, partially
    sequenced pedigrees, -LRC gating and the oracle on top.  (Round 2: the generated kernel caches the conditional
    weights and adds n * P per run of n unchanged sweeps, so the posteriors agree to a few ulps, no longer bit for bit.)"""
    ped = synth.PEDIGREES[pedname]()
    cols = ped.sequenced_cols() if cols is None else cols
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, len(cols) + 1)]), V, seed=77, x_fraction=0.3)
    lk[::7] = np.round(lk[::7])  # some fully certain variants: the LRC gate takes the individual-only branch
    monkeypatch.setenv("FAMSEQ_MCMC_JIT", "0")
    with engine_for(ped, cols) as e:
        generic = e.run(fs.MCMC, lk, fl, burn=burn, rep=rep, seed=99, v_offset=5)
        assert e.info()["jit_launches"] == 0
    monkeypatch.setenv("FAMSEQ_MCMC_JIT", "1")
    with engine_for(ped, cols) as e:
        jit = e.run(fs.MCMC, lk, fl, burn=burn, rep=rep, seed=99, v_offset=5)
        assert e.info()["jit_launches"] >= 1
    assert_same_chains(jit, generic, f"mcmc-jit/{pedname}")
    want = O.run(ped, cols, lk, fl, method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=99, v_offset=5)
    assert_parity(jit, want, 1e-9, f"mcmc-jit/{pedname}")


def test_mcmc_background_compile_switches_kernels_without_changing_results(monkeypatch):
    """Default mode: a large batch starts the compile on a worker thread, batches run on the table-driven kernel until
    the cubin is ready, later ones on the specialised kernel; the results do not depend on which kernel ran (to a few ulps)."""
    import time
    ped = synth.half_sibs()
    lk, fl = synth.synth_likelihoods(ped, 4000, seed=5, x_fraction=0.2)
    monkeypatch.delenv("FAMSEQ_MCMC_JIT", raising=False)
    monkeypatch.setenv("FAMSEQ_JIT_MIN_WORK", "1e6")
    with engine_for(ped) as e:
        first = e.run(fs.MCMC, lk, fl, burn=10, rep=100, seed=3)
        deadline = time.time() + 120
        while e.info()["jit_launches"] == 0 and time.time() < deadline:
            time.sleep(0.5)
            again = e.run(fs.MCMC, lk, fl, burn=10, rep=100, seed=3)
        assert e.info()["jit_launches"] >= 1, "the specialised kernel never became ready"
    assert_same_chains(first, again, "background compile")


def test_mcmc_converges_to_exact_bn():
    """Monte-Carlo check against the exact posterior on a pedigree where the chain mixes (mu = 0.02)."""
    ped = synth.half_sibs()
    prm = fs.Params.default()
    prm.mrate = 0.02
    lk, fl = synth.synth_likelihoods(ped, 64, seed=8)
    lk = np.sqrt(np.sqrt(lk))  # flatten the likelihoods so that the posterior is not degenerate
    fl[:] = fl & 1
    with engine_for(ped, params=prm) as e:
        exact = e.run(fs.BN, lk, fl).post
        runs = np.stack([e.run(fs.MCMC, lk, fl, burn=200, rep=4000, seed=k).post for k in range(16)])
    mean, se = runs.mean(0), runs.std(0, ddof=1) / 4.0
    # entries whose conditional never changes along the chain have a zero seed-to-seed spread and a tiny bias
    # from the states the chain never visits, hence the 1e-4 floor on the standard error
    z = np.abs(mean - exact) / (se + 1e-4)
    assert np.quantile(z, 0.99) < 5.0 and np.abs(mean - exact).max() < 0.01


# ------------------------------------------------------------------------------------------------------
# (c) properties at scale
# ------------------------------------------------------------------------------------------------------
def test_large_trio_es_equals_bn_and_is_shard_invariant():
    ped = synth.trio()
    V = 2_000_000
    lk, fl = synth.synth_likelihoods(ped, V, seed=20261018, x_fraction=0.0)
    with engine_for(ped) as e:
        es = e.run(fs.ES, lk, fl)
        bn = e.run(fs.BN, lk, fl)
        part = e.run(fs.ES, lk[777_777:1_234_567], fl[777_777:1_234_567])
    assert es.status.sum() == 0 and bn.status.sum() == 0
    assert rel_err(bn.post, es.post) < REL_TOL and np.array_equal(bn.single, es.single)
    assert (bn.gt != es.gt).mean() < 1e-6  # exact ties may break differently only through BN's summation order
    assert np.array_equal(part.post, es.post[777_777:1_234_567])
    assert np.abs(es.post.sum(-1) - 1).max() < 1e-12
    sub = slice(0, 20000)
    want = O.run(ped, ped.sequenced_cols(), lk[sub], fl[sub], method=O.ES)
    assert np.array_equal(es.post[sub], want["post"]) and np.array_equal(es.gt[sub], want["gt"].astype(np.uint8))


def test_ped14_es_equals_bn_and_bn_is_deterministic():
    ped = synth.ped14()
    lk, fl = synth.synth_likelihoods(ped, 600, seed=20261021)
    with engine_for(ped) as e:
        es = e.run(fs.ES, lk, fl)
        bn = e.run(fs.BN, lk, fl)
        bn2 = e.run(fs.BN, lk[300:], fl[300:])
    assert rel_err(bn.post, es.post) < REL_TOL
    assert np.array_equal(bn.gt, es.gt)
    assert np.array_equal(bn2.post, bn.post[300:])  # bit-identical whatever the batch split


def test_device_resident_path_matches_host_path():
    torch = pytest.importorskip("torch")
    ped = synth.ped14()
    V = 50_000
    lk, fl = synth.synth_likelihoods(ped, V, seed=3)
    S = lk.shape[1]
    with engine_for(ped) as e:
        host = e.run(fs.ES, lk, fl)
        d_lk = torch.from_numpy(lk).cuda()
        d_fl = torch.from_numpy(fl).cuda()
        d_post = torch.empty((V, S, 3), dtype=torch.float64, device="cuda")
        d_single = torch.empty_like(d_post)
        d_gt = torch.empty((V, S), dtype=torch.uint8, device="cuda")
        d_st = torch.empty(V, dtype=torch.uint8, device="cuda")
        e.run_device(fs.ES, V, d_lk.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), d_single.data_ptr(), d_gt.data_ptr(),
                     d_st.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert e.info()["kernel_launches"] >= 2
    assert np.array_equal(d_post.cpu().numpy(), host.post) and np.array_equal(d_gt.cpu().numpy(), host.gt)
    assert np.array_equal(d_single.cpu().numpy(), host.single) and np.array_equal(d_st.cpu().numpy(), host.status)


# ------------------------------------------------------------------------------------------------------
# nuclear-family fast path (es_nuclear_kernel.cu)
# ------------------------------------------------------------------------------------------------------
def _wide_likelihoods(V, S, seed):
    """Random mantissas over a wide exponent range (incl. subnormal and zero entries): stresses the shared-reciprocal
    division of the fast path, which must return the IEEE quotient bit for bit."""
    rng = np.random.default_rng(seed)
    lk = rng.random((V, S, 3)) * np.exp2(rng.integers(-1040, 1, (V, S, 3)).astype(np.float64))
    lk[rng.random((V, S, 3)) < 0.02] = 0.0
    lk[rng.random((V, S)) < 0.05] = 1.0
    return lk


@pytest.mark.parametrize("n_children", [1, 2, 3, 4, 5])
def test_nuclear_fast_path_is_bit_identical(n_children, monkeypatch):
    rows = [(1, 0, 0, 1), (2, 0, 0, 2)] + [(3 + k, 2, 1, 1 + k % 2) for k in range(n_children)]
    ped = synth._mk(rows)
    V = 400_000 if n_children == 1 else 100_000
    lk, fl = synth.synth_likelihoods(ped, V, seed=50 + n_children, x_fraction=0.3)
    wide = _wide_likelihoods(V, ped.n, 60 + n_children)
    lk = np.concatenate([lk, wide])
    fl = np.concatenate([fl, fl])
    want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.ES)
    with engine_for(ped) as e:
        got = e.run(fs.ES, lk, fl)
    assert_parity(got, want, 0.0, f"nuclear C={n_children}")
    ok = want["status"] == 0
    assert np.array_equal(got.post[ok], want["post"][ok]) and np.array_equal(got.single[ok], want["single"][ok])
    assert 0 < (~ok).sum() < len(ok)
    # the message-program interpreter must give the same bytes for the same pedigree
    monkeypatch.setenv("FAMSEQ_ES_GENERIC", "1")
    with engine_for(ped) as e:
        gen = e.run(fs.ES, lk[:50_000], fl[:50_000])
    assert np.array_equal(gen.post, got.post[:50_000]) and np.array_equal(gen.gt, got.gt[:50_000])
    assert np.array_equal(gen.status, got.status[:50_000])


def test_nuclear_partial_sequencing_and_permuted_rows():
    # children listed before their parents, the father unsequenced, input columns in another order
    rows = [(7, 5, 9, 2), (9, 0, 0, 1), (4, 5, 9, 1), (5, 0, 0, 2)]
    ped = synth._mk(rows)
    cols = [3, 0, 2]
    lk = _wide_likelihoods(20_000, 3, 77)
    lk2, fl = synth.synth_likelihoods(synth.trio(), 20_000, seed=78, x_fraction=0.5)
    for data in (lk, lk2):
        want = O.run(ped, cols, data, fl, method=O.ES)
        with engine_for(ped, cols) as e:
            got = e.run(fs.ES, data, fl)
        assert_parity(got, want, 0.0, "nuclear/partial")
        ok = want["status"] == 0
        assert np.array_equal(got.post[ok], want["post"][ok])


def test_lrc_gate_shortcut_matches_division():
    """-LRC 1 takes the big < sum shortcut in the nuclear kernel; other values divide.  Likelihood rows whose largest
    entry is within a few ulps of the row sum are the cases where the two could differ."""
    ped = synth.trio()
    rng = np.random.default_rng(5)
    V = 50_000
    lk = np.zeros((V, 3, 3))
    lk[:, :, 0] = rng.random((V, 3))
    tiny = np.exp2(rng.integers(-1074, -40, (V, 3)).astype(np.float64))
    lk[:, :, 1] = lk[:, :, 0] * tiny * (rng.random((V, 3)) < 0.7)
    lk[:, :, 2] = 0.0
    for lc in (1.0, 0.999999999):
        prm = fs.Params.default()
        prm.lrc = lc
        want = O.run(ped, ped.sequenced_cols(), lk, None, method=O.ES, lc=lc)
        with engine_for(ped, params=prm) as e:
            got = e.run(fs.ES, lk, None)
        assert_parity(got, want, 0.0, f"lrc gate {lc}")
        assert np.array_equal(got.post, want["post"])


# ------------------------------------------------------------------------------------------------------
# random pedigrees (synth.random_pedigree): several spouses, childless married-in founders, shuffled ped rows,
# unsequenced members, loops -- every method and every kernel against the oracle
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(8))
def test_random_pedigree_all_methods(seed, monkeypatch):
    loops = seed % 4 == 3
    ped = synth.random_pedigree(300 + seed, 5 + seed, loops=loops, shuffle=seed % 2 == 1, unsequenced=0.2 if seed % 3 == 0 else 0.0)
    cols = ped.sequenced_cols()
    S = len(cols)
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, S + 1)]), 300, seed=500 + seed, x_fraction=0.3)
    w_bn = O.run(ped, cols, lk[:24], fl[:24], method=O.BN)
    w_mc = O.run(ped, cols, lk[:60], fl[:60], method=O.MCMC, burn=10, rep=80, rng=O.RNG_PHILOX, seed=9, v_offset=3)
    results = {}
    for jit in ("0", "1"):
        monkeypatch.setenv("FAMSEQ_MCMC_JIT", jit)
        monkeypatch.setenv("FAMSEQ_ES_JIT", jit)
        with engine_for(ped, cols) as e:
            has_loop = e.info()["has_loop"]
            bn = e.run(fs.BN, lk[:24], fl[:24])
            mc = e.run(fs.MCMC, lk[:60], fl[:60], burn=10, rep=80, seed=9, v_offset=3)
            es = None if has_loop else e.run(fs.ES, lk, fl)
        assert_parity(bn, w_bn, REL_TOL, f"random {seed} BN")
        assert_parity(mc, w_mc, 1e-9, f"random {seed} MCMC jit={jit}")
        if es is not None:
            w_es = O.run(ped, cols, lk, fl, method=O.ES)
            ok = w_es["status"] == 0
            assert np.array_equal(es.status, w_es["status"]) and np.array_equal(es.post[ok], w_es["post"][ok]), f"random {seed} ES jit={jit}"
        results[jit] = (mc, es)
    assert_same_chains(results["0"][0], results["1"][0], f"random {seed} MCMC kernels")
    if results["0"][1] is not None:
        assert np.array_equal(results["0"][1].post, results["1"][1].post, equal_nan=True)
