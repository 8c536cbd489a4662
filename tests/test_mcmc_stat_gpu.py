"""Statistical MCMC parity as SURVEY.md section 8(c) specifies it (reference: family::calPostProbMCMC + estGenoProb,
family.cpp:1932-2096, :2101-2297).

tests/golden/mcmc16/*.npz hold, per case, the mean and standard error over R = 16 runs of the UNMODIFIED reference
(libc rand(), srand(1..16); tests/golden/make_mcmc_golden.py) at C5's own setting: 1 000 burn-in + 10 000 sampling
sweeps, on TestData/loftest.txt x fam01..fam06, on the synthetic 40-member looped pedigree and on a chrX case.  The
CUDA kernels run the same inputs with 16 Philox seeds; every posterior entry must satisfy

        |mean_ours - mean_ref| <= 4 sqrt(se_ours^2 + se_ref^2)                                     (z = 4)

BN: the section also asks for agreement with the exact posterior on pedigrees with N <= 11.  At the default mutation
rate (1e-7) the single-site sampler does not reach that: the REFERENCE's own 16-run mean differs from the REFERENCE's
own BN posterior by more than 4 standard errors on a third of the entries (up to 0.07 absolute on loftest x fam04;
the goldens carry both, test_reference_sampler_itself_misses_bn_at_the_default_mutation_rate pins the fact).  What is
tested instead is that the kernels are no farther from BN than the reference is, within the same Monte-Carlo bound;
convergence to BN proper is tested where the chain mixes (test_parity_gpu.py::test_mcmc_converges_to_exact_bn)."""
import numpy as np
import pytest

import famseq_b200 as fs
from tests.util import MCMC_Z, CasePed, load_mcmc16, mcmc16_cases, mcmc_z_scores

pytestmark = pytest.mark.gpu
SEEDS = [1000 + k for k in range(16)]


def _runs(c, jit, monkeypatch):
    monkeypatch.setenv("FAMSEQ_MCMC_JIT", jit)
    ped = CasePed(c)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, c["cols"].tolist(), device=0) as e:
        runs = [e.run(fs.MCMC, c["lk"], c["flags"], burn=int(c["burn"]), rep=int(c["rep"]), seed=s) for s in SEEDS]
        info = e.info()
    assert (info["jit_launches"] > 0) == (jit == "1")
    return runs


@pytest.mark.parametrize("jit", ["1", "0"])
@pytest.mark.parametrize("name", mcmc16_cases())
def test_sixteen_seeds_against_sixteen_reference_runs(name, jit, monkeypatch):
    if jit == "0" and name not in ("loftest_fam01", "syn_halfsibs_x"):
        pytest.skip("the table-driven kernel returns the bytes of the generated one (test_parity_gpu.py); two cases suffice")
    c = load_mcmc16(name)
    runs = _runs(c, jit, monkeypatch)
    assert np.array_equal(runs[0].single, c["single"]) or name == "syn_halfsibs_x"
    ok = c["status"] == 0  # variants on which all 16 reference runs succeeded
    for r in runs:  # ... must succeed here too (a chrX chain can get stuck; none of the chosen seeds does on these)
        assert not r.status[ok].any(), f"{name}: a chain failed on a variant the reference computes"
    z, mean, se = mcmc_z_scores(np.stack([r.post for r in runs]), c["mean"], c["se"], ok)
    assert z.max() <= MCMC_Z, f"{name}: max z = {z.max():.2f} over {z.size} entries"
    if c["bn"].size:  # no farther from the exact posterior than the reference's own sampler, within the same bound
        bound = MCMC_Z * np.sqrt(se ** 2 + c["se"][ok] ** 2) + np.abs(c["mean"][ok] - c["bn"][ok]) + 1e-9 * c["bn"][ok]
        assert (np.abs(mean - c["bn"][ok]) <= bound).all(), f"{name}: farther from BN than the reference sampler"


def test_reference_sampler_itself_misses_bn_at_the_default_mutation_rate():
    """Documents why the BN clause above is relative: computed from the goldens alone (reference MCMC vs reference BN)."""
    far = total = 0
    for name in mcmc16_cases():
        c = load_mcmc16(name)
        if not c["bn"].size or not name.startswith("loftest"):
            continue
        d, se = np.abs(c["mean"] - c["bn"]), c["se"]
        far += int((d > MCMC_Z * se).sum())
        total += d.size
    assert far > total // 5, (far, total)


def _wide_likelihoods(V, S, seed):
    rng = np.random.default_rng(seed)
    lk = rng.random((V, S, 3)) * np.exp2(rng.integers(-330, 1, (V, S, 3)).astype(np.float64))
    lk[rng.random((V, S, 3)) < 0.05] = 0.0
    return lk


@pytest.mark.parametrize("pedname", ["half_sibs", "ped40"])
def test_fixup_path_of_the_generated_kernel(pedname, monkeypatch):
    """Likelihoods over a wide exponent range drive some weight sums out of the generated kernel's fast range: it marks
    those chains (status 2) and the table-driven kernel, launched right behind in fix-up mode, redoes them
    (engine.cu dispatch, gibbs_jit.cu).  The counter fs_info.mcmc_fixups must show that this happened, and the results
    must be those of the table-driven kernel alone (same chains) and agree with the oracle."""
    from famseq_b200 import synth
    from oracle import oracle as O
    from tests.util import assert_parity, assert_same_chains

    ped = synth.PEDIGREES[pedname]()
    cols = ped.sequenced_cols()
    V, burn, rep = 600, 5, 40
    lk = _wide_likelihoods(V, len(cols), 3)
    fl = (np.arange(V) % 4).astype(np.uint8)
    monkeypatch.setenv("FAMSEQ_MCMC_JIT", "0")
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=0) as e:
        table = e.run(fs.MCMC, lk, fl, burn=burn, rep=rep, seed=99, v_offset=7)
        assert e.info()["mcmc_fixups"] == 0
    monkeypatch.setenv("FAMSEQ_MCMC_JIT", "1")
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=0) as e:
        jit = e.run(fs.MCMC, lk, fl, burn=burn, rep=rep, seed=99, v_offset=7)
        info = e.info()
    assert info["jit_launches"] >= 1 and 0 < info["mcmc_fixups"] < V, info
    assert not (jit.status == 2).any()
    assert_same_chains(jit, table, f"fix-up/{pedname}")
    want = O.run(ped, cols, lk, fl, method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=99, v_offset=7)
    assert_parity(jit, want, 1e-9, f"fix-up/{pedname}")


def test_pilot_picks_the_generator_that_suits_the_data(monkeypatch):
    """Two generated Gibbs kernels (gibbs_jit.cu): cached conditionals, fast where the chains sit still, and dense sweeps,
    whose speed does not depend on the data.  On the first large batch the engine runs a short pilot of the cached kernel and
    counts how often its warps had to redo a group member by member: pedigree-consistent sequencing data keep the cached
    kernel (fs_info.gibbs_generator == 2), Mendel-inconsistent likelihoods -- chains that keep moving -- switch the engine
    to the dense one (1).  Either way the results are the oracle's."""
    from famseq_b200 import synth
    from oracle import oracle as O
    from tests.util import assert_parity

    monkeypatch.setenv("FAMSEQ_MCMC_JIT", "1")
    monkeypatch.delenv("FAMSEQ_JIT_CACHED", raising=False)
    ped = synth.ped14()
    cols = ped.sequenced_cols()
    V, burn, rep = 4096, 200, 1000
    consistent, fl = synth.synth_likelihoods(ped, V, seed=31)
    unrelated, _ = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, len(cols) + 1)]), V, seed=32)
    fl[:] = fl & 1
    for lk, want_generator in ((consistent, 2), (unrelated, 1)):
        with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=0) as e:
            got = e.run(fs.MCMC, lk, fl, burn=burn, rep=rep, seed=4, v_offset=9)
            assert e.info()["gibbs_generator"] == want_generator, e.info()
        want = O.run(ped, cols, lk[:256], fl[:256], method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=4, v_offset=9)
        head = fs.Result(got.post[:256], got.single[:256], got.gt[:256], got.status[:256])
        assert_parity(head, want, 1e-9, f"pilot/{want_generator}")
