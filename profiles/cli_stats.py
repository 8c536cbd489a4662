"""The command line on a synthetic trio VCF, FAMSEQ_STATS=1, a few runs: where the wall clock goes (tuning aid).
    python profiles/cli_stats.py [records] [runs]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from famseq_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ped = synth.trio()
exe = os.path.join(ROOT, "famseq_b200", "bin", "FamSeq")
with tempfile.TemporaryDirectory() as d:
    vcf, pp, out = os.path.join(d, "in.vcf"), os.path.join(d, "fam.ped"), os.path.join(d, "out.vcf")
    pl, fl = synth.synth_pl(ped, n, seed=99)
    synth.write_vcf(vcf, ped, pl, fl)
    ped.write(pp)
    for r in range(runs):
        t0 = time.perf_counter()
        p = subprocess.run([exe, "vcf", "-vcfFile", vcf, "-pedFile", pp, "-output", out, "-method", "2"], capture_output=True, text=True,
                           env=dict(os.environ, FAMSEQ_STATS="1"))
        wall = time.perf_counter() - t0
        line = [l for l in p.stderr.splitlines() if l.startswith("{")]
        print(f"run {r}: rc={p.returncode} wall={wall:.3f}s", line[-1] if line else p.stderr[-300:])
