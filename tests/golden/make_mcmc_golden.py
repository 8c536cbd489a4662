"""Generates tests/golden/mcmc16/*.npz: the statistical MCMC reference of SURVEY.md section 8(c).

For every case the UNMODIFIED reference engine (oracle/_ref/ref_harness, which calls srand(seed) before the batch)
is run R = 16 times with seeds 1..16 -- family::calPostProbMCMC, family.cpp:1932-2096 -- and the per-entry mean and
standard error over the 16 runs are stored next to the inputs.  tests/test_mcmc_stat_*.py compare the CUDA kernels
(16 Philox seeds) against them entry by entry: |mean_ours - mean_ref| <= 4 sqrt(se_ours^2 + se_ref^2).

Run it in the build container (the only place /root/reference exists):
    python tests/golden/make_mcmc_golden.py
Cases: TestData/loftest.txt x fam01..fam06 and the synthetic 40-member looped pedigree (C5), all at C5's own setting
of 1 000 burn-in + 10 000 sampling sweeps, 32 variants each; plus the exact BN posterior of the reference for the
pedigrees with N <= 11.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from famseq_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.golden.make_golden import match_columns, parse_lk_file  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mcmc16")
TD = "/root/reference/TestData"
R, BURN, REP, V = 16, 1000, 10000, 32


def run_case(name, ped, cols, lk, flags, with_bn):
    with tempfile.TemporaryDirectory() as td:
        pp = os.path.join(td, "p.ped")
        synth.PedFile(ped.ids, ped.mids, ped.fids, ped.genders, ped.names).write(pp)
        runs = [O.run_ref(pp, cols, lk, flags=flags, method=3, burn=BURN, rep=REP, seed=k) for k in range(1, R + 1)]
        bn = O.run_ref(pp, cols, lk, flags=flags, method=1)["post"] if with_bn else np.zeros((0,))
    post = np.stack([r["post"] for r in runs])
    # a chrX chain can get stuck where a member's weights are all zero for the whole run (row sum 0 -> the reference
    # returns false), which depends on the random start: `status` counts the failing runs per variant, and the
    # statistics of such variants (excluded by the tests) are over the successful runs only
    status = np.sum([r["status"] for r in runs], axis=0).astype(np.uint8)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        ids=np.array(ped.ids, np.int32), mids=np.array(ped.mids, np.int32), fids=np.array(ped.fids, np.int32),
        genders=np.array(ped.genders, np.int32), cols=np.array(cols, np.int32), lk=lk,
        flags=np.zeros(lk.shape[0], np.uint8) if flags is None else flags, burn=BURN, rep=REP, runs=R,
        mean=post.mean(0), se=post.std(0, ddof=1) / np.sqrt(R), single=runs[0]["single"], status=status, bn=bn)
    print(f"{name}: V={lk.shape[0]} S={lk.shape[1]} N={ped.n} max se={post.std(0, ddof=1).max() / 4:.3g} variants with failing runs={int((status > 0).sum())}")


def main():
    os.makedirs(OUT, exist_ok=True)
    names, rows = parse_lk_file(f"{TD}/loftest.txt")
    for k in range(1, 7):
        ped = O.Pedigree.read(f"{TD}/fam0{k}.ped")
        cols, idx = match_columns(names, ped)
        lk = np.array([[[float(x) for x in r.split("\t")[ci].split(",")[:3]] for ci in idx] for r in rows[:V]])
        run_case(f"loftest_fam0{k}", ped, cols, lk, None, with_bn=True)
    p = synth.ped40()
    lk, fl = synth.synth_likelihoods(p, V, 20261018 + 3)
    run_case("syn_ped40", p, p.sequenced_cols(), lk, fl, with_bn=False)
    p = synth.half_sibs()  # chrX and Known flags on a small pedigree with several spouses
    lk, fl = synth.synth_likelihoods(p, V, 23, x_fraction=0.4)
    run_case("syn_halfsibs_x", p, p.sequenced_cols(), lk, fl, with_bn=True)


if __name__ == "__main__":
    main()
