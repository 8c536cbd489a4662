"""Generates tests/golden/*.npz: input/output vectors produced by the UNMODIFIED reference engine
(oracle/_ref/ref_harness, compiled from /root/reference/src by oracle/Makefile).

Run it in the build container (the only place /root/reference exists):
    python tests/golden/make_golden.py
Each case stores the ped rows, the raw likelihood batch, the flags, the parameters and the
reference's raw doubles (post, single, gt, status, and the full N x 3 matrices).
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

OUT = os.path.dirname(os.path.abspath(__file__))
TD = "/root/reference/TestData"


def parse_lk_file(path):
    lines = open(path).read().split("\n")
    names = lines[0].split("\t")
    rows = [l for l in lines[1:] if len(l) >= 2]
    return names, rows


def match_columns(names, ped):
    cols, idx = [], []
    for ci, nm in enumerate(names):
        for j, pn in enumerate(ped.names):
            if nm == pn:
                cols.append(j)
                idx.append(ci)
                break
    return cols, idx


def run_case(name, ped, cols, lk, flags, method, **kw):
    with tempfile.TemporaryDirectory() as td:
        pp = os.path.join(td, "p.ped")
        synth.PedFile(ped.ids, ped.mids, ped.fids, ped.genders, ped.names).write(pp)
        r = O.run_ref(pp, cols, lk, flags=flags, method=method, **kw)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        ids=np.array(ped.ids, np.int32), mids=np.array(ped.mids, np.int32), fids=np.array(ped.fids, np.int32),
        genders=np.array(ped.genders, np.int32), cols=np.array(cols, np.int32), lk=lk,
        flags=np.zeros(lk.shape[0], np.uint8) if flags is None else flags, method=method,
        mrate=kw.get("mrate", 1e-7), lc=kw.get("lc", 1.0),
        priors=np.asarray(kw.get("priors", O.DEFAULT_PRIORS), np.float64), burn=kw.get("burn", 0), rep=kw.get("rep", 0),
        seed=kw.get("seed", -1), post=r["post"], single=r["single"], gt=r["gt"], status=r["status"],
        post_full=r["post_full"], single_full=r["single_full"])
    print(f"{name}: V={lk.shape[0]} S={lk.shape[1]} N={ped.n} failed={int(r['status'].sum())}")


def main():
    # ---- C2: TestData/loftest.txt x fam01..06, BN and ES (and MCMC with the libc stream, seed 1) ----------
    names, rows = parse_lk_file(f"{TD}/loftest.txt")
    for k in range(1, 7):
        ped = O.Pedigree.read(f"{TD}/fam0{k}.ped")
        cols, idx = match_columns(names, ped)
        lk = np.array([[[float(x) for x in r.split("\t")[ci].split(",")[:3]] for ci in idx] for r in rows])
        for method, tag in ((1, "bn"), (2, "es")):
            run_case(f"loftest_fam0{k}_{tag}", ped, cols, lk, None, method)
        run_case(f"loftest_fam0{k}_mcmc", ped, cols, lk[:20], None, 3, burn=200, rep=2000, seed=1)

    # ---- C1: TestData/test.vcf (the SNP records that reach the engine) x fam01, BN and ES -------------------
    ped = O.Pedigree.read(f"{TD}/fam01.ped")
    header = None
    lks, fls = [], []
    for line in open(f"{TD}/test.vcf"):
        if line.startswith("#CHROM"):
            header = line.rstrip("\n").split("\t")
        if line.startswith("#") or len(line) < 2:
            continue
        t = line.rstrip("\n").split("\t")
        fmt = t[8].split(":")
        if "PL" not in fmt or len(t[3]) != 1 or len(t[4]) != 1 or t[4] in ".-":
            continue
        cols, idx = match_columns(header[9:], ped)
        ipl = max(i for i, f in enumerate(fmt) if f in ("PL", "GL"))
        row = []
        for ci in idx:
            f = t[9 + ci].split(":")
            row.append([10.0 ** (-abs(float(x)) / 10.0) for x in f[ipl].split(",")[:3]])
        lks.append(row)
        fls.append((1 if t[2] != "." else 0) | (2 if t[0] in ("X", "chrX", "CHRX") else 0))
    lk = np.array(lks)
    # the reference decodes with libm pow(); numpy's ** may differ in the last bit, so decode through C
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.pow.restype = ctypes.c_double
    libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
    k = 0
    for line in open(f"{TD}/test.vcf"):
        if line.startswith("#") or len(line) < 2:
            continue
        t = line.rstrip("\n").split("\t")
        fmt = t[8].split(":")
        if "PL" not in fmt or len(t[3]) != 1 or len(t[4]) != 1 or t[4] in ".-":
            continue
        ipl = max(i for i, f in enumerate(fmt) if f in ("PL", "GL"))
        for s, ci in enumerate(idx):
            f = t[9 + ci].split(":")
            for g, x in enumerate(f[ipl].split(",")[:3]):
                lk[k, s, g] = libm.pow(10.0, -abs(float(x)) / 10.0)
        k += 1
    flags = np.array(fls, np.uint8)
    for method, tag in ((1, "bn"), (2, "es")):
        run_case(f"testvcf_fam01_{tag}", ped, cols, lk, flags, method)

    # ---- synthetic pedigrees: Known / chrX flags, several spouses, LRC and mutation-rate variants ----------
    def syn(name, pedf, V, seed, method, x_fraction=0.25, **kw):
        p = pedf()
        lk, fl = synth.synth_likelihoods(p, V, seed, x_fraction=x_fraction)
        run_case(name, p, p.sequenced_cols(), lk, fl, method, **kw)

    syn("syn_trio_es", synth.trio, 4096, 11, 2)
    syn("syn_trio_bn", synth.trio, 4096, 11, 1)
    syn("syn_trio_es_mu0", synth.trio, 512, 12, 2, mrate=0.0)
    syn("syn_trio_es_lrc", synth.trio, 512, 13, 2, lc=0.999999)
    syn("syn_ped14_es", synth.ped14, 1024, 14, 2)
    syn("syn_ped14_bn", synth.ped14, 12, 14, 1)
    syn("syn_halfsibs_es", synth.half_sibs, 512, 15, 2)
    syn("syn_halfsibs_bn", synth.half_sibs, 64, 15, 1)
    syn("syn_threewives_es", synth.three_wives, 512, 16, 2)
    syn("syn_threewives_bn", synth.three_wives, 64, 16, 1)
    syn("syn_cousins_bn", synth.cousins_loop, 128, 17, 1)
    pri = np.array([[0.98, 0.015, 0.005], [0.3, 0.4, 0.3], [0.99, 0.0, 0.01], [0.6, 0.0, 0.4]])
    syn("syn_ped14_es_priors", synth.ped14, 256, 18, 2, priors=pri, mrate=1e-4)

    # partly sequenced pedigree (members named NA keep likelihood 1,1,1)
    p = synth.ped14()
    for i in (0, 3, 6, 9, 12):
        p.names[i] = "NA"
    lk, fl = synth.synth_likelihoods(p, 512, 19, x_fraction=0.25)
    run_case("syn_ped14_partial_es", p, p.sequenced_cols(), lk, fl, 2)
    run_case("syn_ped14_partial_bn", p, p.sequenced_cols(), lk[:8], fl[:8], 1)

    # edge cases: zero likelihood rows (status 1), exact ties, extreme PLs (underflow to 0 / subnormal)
    p = synth.trio()
    lk = np.ones((8, 3, 3))
    lk[0] = 0.0                                  # every sample impossible -> row sum 0 -> false
    lk[1, 0] = [0.0, 0.0, 0.0]                   # one sample impossible
    lk[2] = [[1e-300, 1e-300, 1e-300]] * 3       # products underflow
    lk[3] = [[0.5, 0.5, 0.5]] * 3                # flat likelihoods: ties
    lk[4] = [[1.0, 0.0, 0.0]] * 3                # all certain -> LRC gate keeps the single posterior
    lk[5] = [[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]]  # Mendelian error, certain
    lk[6] = [[1e-320, 1e-310, 1e-315], [1.0, 1e-5, 1e-9], [1e-3, 1.0, 1e-3]]  # subnormal likelihoods
    lk[7] = [[0.2, 0.3, 0.5], [0.5, 0.3, 0.2], [0.1, 0.8, 0.1]]
    fl = np.array([0, 0, 0, 1, 0, 1, 0, 3], np.uint8)
    for method, tag in ((1, "bn"), (2, "es")):
        run_case(f"edge_trio_{tag}", p, p.sequenced_cols(), lk, fl, method)


if __name__ == "__main__":
    main()
