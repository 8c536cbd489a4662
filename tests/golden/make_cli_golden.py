"""Generates tests/golden/cli/: small input files and the output of the UNMODIFIED reference command line
(oracle/_ref/FamSeq, built from /root/reference/src by oracle/Makefile) for a set of flag combinations.

    python tests/golden/make_cli_golden.py        # build container only (needs /root/reference)

Inputs: the reference's TestData pedigrees and likelihood table, a subset of TestData/test.vcf (full header,
every record carrying PL, the first 150 other records) and synthetic VCFs that exercise what TestData does
not: chrX, missing samples (./.), a GL tag, Y / MT / non-SNP records, a location file.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from famseq_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "cli")
TD = "/root/reference/TestData"
REF = os.path.join(ROOT, "oracle", "_ref", "FamSeq")


def make_inputs():
    os.makedirs(OUT, exist_ok=True)
    for k in range(1, 7):
        shutil.copy(f"{TD}/fam0{k}.ped", f"{OUT}/fam0{k}.ped")
    shutil.copy(f"{TD}/loftest.txt", f"{OUT}/loftest.txt")
    others = 0
    with open(f"{OUT}/test_subset.vcf", "w") as fo:
        for line in open(f"{TD}/test.vcf"):
            if line.startswith("#"):
                fo.write(line)
                continue
            fmt = line.split("\t")[8] if line.count("\t") > 8 else ""
            if "PL" in fmt.split(":"):
                fo.write(line)
            elif others < 150:
                fo.write(line)
                others += 1
    # synthetic trio VCF with the record kinds TestData lacks
    ped = synth.trio()
    ped.write(f"{OUT}/trio.ped")
    pl, fl = synth.synth_pl(ped, 400, seed=5, x_fraction=0.2)
    synth.write_vcf(f"{OUT}/trio_syn.vcf", ped, pl, fl)
    lines = open(f"{OUT}/trio_syn.vcf").read().split("\n")
    body = [l for l in lines if l and not l.startswith("#")]
    head = [l for l in lines if l.startswith("#")]
    rng = np.random.default_rng(3)
    for i in range(0, len(body), 7):  # missing samples
        f = body[i].split("\t")
        f[9 + int(rng.integers(0, 3))] = "./."
        body[i] = "\t".join(f)
    f = body[3].split("\t"); f[9] = f[10] = f[11] = "./."; body[3] = "\t".join(f)      # all missing
    f = body[5].split("\t"); f[0] = "Y"; body[5] = "\t".join(f)                           # chrY
    f = body[6].split("\t"); f[0] = "MT"; body[6] = "\t".join(f)                          # mitochondrial
    f = body[8].split("\t"); f[0] = "chrX"; body[8] = "\t".join(f)
    f = body[9].split("\t"); f[0] = "chr7"; body[9] = "\t".join(f)
    f = body[10].split("\t"); f[0] = "GL000192.1"; body[10] = "\t".join(f)               # unplaced contig
    f = body[12].split("\t"); f[4] = "AT"; body[12] = "\t".join(f)                        # indel
    f = body[13].split("\t"); f[3] = "."; body[13] = "\t".join(f)                         # REF .
    f = body[15].split("\t"); f[4] = "."; body[15] = "\t".join(f)                         # ALT . (hom-ref block)
    f = body[16].split("\t"); f[8] = "GT:DP"; f[9:] = ["0/0:30"] * 3; body[16] = "\t".join(f)   # no PL in FORMAT
    f = body[17].split("\t"); f[8] = "GT:DP:GL"; body[17] = "\t".join(f)                  # GL decoded like PL
    f = body[18].split("\t"); f[10] = "0/0:30"; body[18] = "\t".join(f)                   # field count mismatch
    f = body[19].split("\t"); f[9] = "0/0:30:0,0,0"; f[10] = "0/0:30:0,9000,9000"; f[11] = "0/0:30:9000,0,9000"  # underflow to 0
    body[19] = "\t".join(f)
    f = body[20].split("\t"); f[9] = "0/0:9:9000,9000,9000"; body[20] = "\t".join(f)      # impossible sample -> NA
    with open(f"{OUT}/trio_syn.vcf", "w") as fo:
        fo.write("\n".join(head + body) + "\n")
    with open(f"{OUT}/trio_syn.loc", "w") as fo:
        for l in body[::3]:
            c = l.split("\t")
            chrom = c[0][3:] if c[0].startswith("chr") else c[0]
            fo.write(f"{chrom}\t{c[1]}\n")
    # likelihood tables in the other three encodings
    raw = open(f"{OUT}/loftest.txt").read().split("\n")
    for tag, fn in (("log10", np.log10), ("ln", np.log), ("PS", lambda x: -10 * np.log10(x))):
        with open(f"{OUT}/loftest_{tag}.txt", "w") as fo:
            fo.write(raw[0] + "\n")
            for r in raw[1:21]:
                cells = []
                for cell in r.split("\t"):
                    if not cell:
                        cells.append(cell)
                        continue
                    cells.append(",".join(repr(float(fn(float(x)))) for x in cell.split(",")))
                fo.write("\t".join(cells) + "\n")
    # 14-member pedigree, ES and BN from a VCF
    ped = synth.ped14()
    ped.write(f"{OUT}/ped14.ped")
    pl, fl = synth.synth_pl(ped, 40, seed=6, x_fraction=0.2)
    synth.write_vcf(f"{OUT}/ped14_syn.vcf", ped, pl, fl)
    ped = synth.cousins_loop()
    ped.write(f"{OUT}/cousins.ped")
    pl, fl = synth.synth_pl(ped, 30, seed=7)
    synth.write_vcf(f"{OUT}/cousins_syn.vcf", ped, pl, fl)


CASES = [
    # name, argv (input/output names are relative to tests/golden/cli)
    ("lk_fam01_bn", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam01.ped", "-method", "1"]),
    ("lk_fam01_es", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam01.ped", "-method", "2"]),
    ("lk_fam02_es", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam02.ped", "-method", "2"]),
    ("lk_fam03_es", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam03.ped", "-method", "2"]),
    ("lk_fam04_es", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam04.ped", "-method", "2"]),
    ("lk_fam05_es", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam05.ped", "-method", "2"]),
    ("lk_fam06_es", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam06.ped", "-method", "2"]),
    ("lk_fam04_bn", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam04.ped"]),
    ("lk_fam05_log10", ["LK", "-lkFile", "loftest_log10.txt", "-pedFile", "fam05.ped", "-method", "2", "-lkType", "log10"]),
    ("lk_fam05_ln", ["LK", "-lkFile", "loftest_ln.txt", "-pedFile", "fam05.ped", "-method", "2", "-lkType", "ln"]),
    ("lk_fam05_ps", ["LK", "-lkFile", "loftest_PS.txt", "-pedFile", "fam05.ped", "-method", "2", "-lkType", "PS"]),
    ("lk_fam06_opts", ["LK", "-lkFile", "loftest.txt", "-pedFile", "fam06.ped", "-method", "2", "-mRate", "1e-3", "-genoProbN",
                       "0.9", "0.08", "0.02", "-LRC", "0.99"]),
    ("vcf_fam01_bn", ["vcf", "-vcfFile", "test_subset.vcf", "-pedFile", "fam01.ped", "-method", "1"]),
    ("vcf_fam01_es", ["vcf", "-vcfFile", "test_subset.vcf", "-pedFile", "fam01.ped", "-method", "2"]),
    ("vcf_fam01_es_v", ["vcf", "-vcfFile", "test_subset.vcf", "-pedFile", "fam01.ped", "-method", "2", "-v"]),
    ("vcf_fam01_es_a", ["vcf", "-vcfFile", "test_subset.vcf", "-pedFile", "fam01.ped", "-method", "2", "-a"]),
    ("vcf_fam05_es", ["vcf", "-vcfFile", "test_subset.vcf", "-pedFile", "fam05.ped", "-method", "2", "-v"]),
    ("vcf_trio_es", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped", "-method", "2"]),
    ("vcf_trio_es_a", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped", "-method", "2", "-a"]),
    ("vcf_trio_es_d", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped", "-method", "2", "-d"]),
    ("vcf_trio_bn", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped"]),
    ("vcf_trio_es_loc", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped", "-method", "2", "-l", "trio_syn.loc"]),
    ("vcf_trio_es_priors", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped", "-method", "2", "-genoProbK", "0.3", "0.4", "0.3",
                            "-genoProbXN", "0.99", "0.01", "-genoProbXK", "0.6", "0.4", "-mRate", "0"]),
    ("vcf_ped14_es", ["vcf", "-vcfFile", "ped14_syn.vcf", "-pedFile", "ped14.ped", "-method", "2"]),
    ("vcf_ped14_bn", ["vcf", "-vcfFile", "ped14_syn.vcf", "-pedFile", "ped14.ped", "-method", "1"]),
    ("vcf_cousins_bn", ["vcf", "-vcfFile", "cousins_syn.vcf", "-pedFile", "cousins.ped", "-method", "1"]),
    ("vcf_trio_mcmc", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped", "-method", "3", "-numBurnIn", "100", "-numRep", "2000"]),
    ("vcf_badflags", ["vcf", "-vcfFile", "trio_syn.vcf", "-pedFile", "trio.ped", "-method", "7", "-mRate", "0.9", "-bogus", "stray", "-v", "-a",
                      "-LRC"]),
]


def main():
    make_inputs()
    manifest = {}
    for name, argv in CASES:
        out = f"{name}.expected"
        r = subprocess.run([REF] + argv + ["-output", out], cwd=OUT, capture_output=True, text=True)
        manifest[name] = {"argv": argv, "returncode": r.returncode, "stdout": r.stdout, "stderr": r.stderr}
        print(name, r.returncode, os.path.getsize(os.path.join(OUT, out)))
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
