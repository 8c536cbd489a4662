"""Shared helpers for the parity tests."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-9  # north_star: BN and ES posteriors within 1e-9 relative of the reference CPU path


def golden_cases(pattern="*"):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, pattern + ".npz")))


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


class CasePed:
    """Duck-typed pedigree (same attribute names as oracle.Pedigree / synth.PedFile)."""

    def __init__(self, c):
        self.ids, self.mids, self.fids, self.genders = (c[k].tolist() for k in ("ids", "mids", "fids", "genders"))
        self.names = ["x"] * len(self.ids)
        self.n = len(self.ids)


def rel_err(a, b):
    """max |a-b| / |b| with exact zeros required to match exactly."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    zero_mismatch = (a == 0) != (b == 0)
    if zero_mismatch.any():
        return np.inf
    nz = b != 0
    if not nz.any():
        return 0.0
    return float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])))


def assert_parity(got, want, tol=REL_TOL, what=""):
    """got: object with post/single/gt/status; want: dict with the same keys (reference or oracle)."""
    ws = np.asarray(want["status"]).astype(np.uint8)
    gs = np.asarray(got.status).astype(np.uint8)
    assert np.array_equal(gs, ws), f"{what}: status differs at {np.nonzero(gs != ws)[0][:10]}"
    ok = ws == 0
    wgt = np.asarray(want["gt"]).astype(np.int64) & 0xff
    ggt = np.asarray(got.gt).astype(np.int64) & 0xff
    assert np.array_equal(ggt[ok], wgt[ok]), f"{what}: called genotypes differ ({int((ggt[ok] != wgt[ok]).sum())} entries)"
    e1 = rel_err(np.asarray(got.single)[ok], np.asarray(want["single"])[ok])
    e2 = rel_err(np.asarray(got.post)[ok], np.asarray(want["post"])[ok])
    assert e1 <= tol, f"{what}: single posterior rel err {e1:.3e} > {tol}"
    assert e2 <= tol, f"{what}: pedigree posterior rel err {e2:.3e} > {tol}"
    return e1, e2


# ---- statistical MCMC parity ------------------------------------------------------
MCMC16_DIR = os.path.join(GOLDEN_DIR, "mcmc16")
MCMC_Z = 4.0


def mcmc16_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(MCMC16_DIR, "*.npz")))


def load_mcmc16(name):
    return dict(np.load(os.path.join(MCMC16_DIR, name + ".npz")))


def mcmc_z_scores(ours_runs, ref_mean, ref_se, ok):
    """Per-entry z of SURVEY 8(c): |mean_ours - mean_ref| / sqrt(se_ours^2 + se_ref^2), standard errors from the R
    independent seeds of each side.  A difference below the floating-point parity tolerance (1e-9 relative, REL_TOL)
    is no difference: many entries are (nearly) deterministic functions of the likelihoods -- the chains never leave
    one state -- so both standard errors are ~1e-17 or exactly zero and rounding alone would dominate z."""
    ours_runs = np.asarray(ours_runs)[:, ok]
    R = ours_runs.shape[0]
    mean = ours_runs.mean(0)
    se = ours_runs.std(0, ddof=1) / np.sqrt(R)
    rm, rs = ref_mean[ok], ref_se[ok]
    den = np.sqrt(se * se + rs * rs)
    diff = np.maximum(0.0, np.abs(mean - rm) - REL_TOL * np.abs(rm))
    z = np.where(diff == 0, 0.0, diff / np.where(den == 0, 1e-300, den))
    return z, mean, se


def assert_same_chains(a, b, what="", tol=1e-12):
    """Two Gibbs kernels on the same Philox stream visit the same states; the generated kernel adds n * P at the end of a
    run of n sweeps with unchanged weights where the table-driven one adds P n times, so the posteriors agree to a few
    ulps (tol), not bit for bit.  Called genotypes must agree wherever the two largest posteriors are not tied to tol."""
    assert np.array_equal(a.status, b.status), f"{what}: status differs"
    if a.single is not None and b.single is not None:
        assert np.array_equal(a.single, b.single, equal_nan=True), f"{what}: individual-only posteriors differ"
    ok = a.status == 0
    pa, pb = np.asarray(a.post)[ok], np.asarray(b.post)[ok]
    assert rel_err(pa, pb) <= tol, f"{what}: posteriors differ by {rel_err(pa, pb):.3e} relative"
    top = np.sort(pb, axis=-1)
    clear = (top[..., 2] - top[..., 1]) > 1e-9 * top[..., 2]
    assert np.array_equal(np.asarray(a.gt)[ok][clear], np.asarray(b.gt)[ok][clear]), f"{what}: called genotypes differ"
