"""The statistical MCMC reference (tests/golden/mcmc16, 16 runs of the unmodified reference engine with libc rand())
against the oracle drawing from the Philox stream the CUDA kernels use: pins the counter-based variant of the sampler to
the reference in the SURVEY 8(c) sense (z = 4 per entry, 16 seeds on both sides) without a GPU.  The GPU twin of this
test (tests/test_mcmc_stat_gpu.py) runs the kernels on the same cases and seeds."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import MCMC_Z, CasePed, load_mcmc16, mcmc16_cases, mcmc_z_scores


def test_goldens_are_present_and_have_the_c5_setting():
    names = mcmc16_cases()
    assert {f"loftest_fam0{k}" for k in range(1, 7)} <= set(names) and "syn_ped40" in names
    for n in names:
        c = load_mcmc16(n)
        assert int(c["burn"]) == 1000 and int(c["rep"]) == 10000 and int(c["runs"]) == 16 and c["lk"].shape[0] >= 32


@pytest.mark.parametrize("name", mcmc16_cases())
def test_philox_oracle_against_sixteen_reference_runs(name):
    c = load_mcmc16(name)
    ped = CasePed(c)
    runs = np.stack([O.run(ped, c["cols"].tolist(), c["lk"], c["flags"], method=O.MCMC, burn=int(c["burn"]), rep=int(c["rep"]),
                           rng=O.RNG_PHILOX, seed=1000 + k)["post"] for k in range(16)])
    z, _, _ = mcmc_z_scores(runs, c["mean"], c["se"], c["status"] == 0)
    assert z.max() <= MCMC_Z, f"{name}: max z = {z.max():.2f}"
