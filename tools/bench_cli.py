#!/usr/bin/env python
"""End-to-end (file -> file) throughput of the command line: this repo's `FamSeq` against the reference binary
(oracle/_ref/FamSeq) on the same synthetic trio VCF.  Prints one JSON line.

    python tools/bench_cli.py [--variants 1000000] [--method 2]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from famseq_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=1_000_000)
    ap.add_argument("--method", default="2")
    ap.add_argument("--ref-variants", type=int, default=200_000, help="records given to the (slow) reference binary")
    args = ap.parse_args()
    ours = os.path.join(ROOT, "famseq_b200", "bin", "FamSeq")
    ref = os.path.join(ROOT, "oracle", "_ref", "FamSeq")
    ped = synth.trio()
    with tempfile.TemporaryDirectory() as td:
        pp = os.path.join(td, "trio.ped")
        ped.write(pp)
        pl, fl = synth.synth_pl(ped, args.variants, seed=99)
        big, small = os.path.join(td, "big.vcf"), os.path.join(td, "small.vcf")
        t0 = time.perf_counter()
        synth.write_vcf(big, ped, pl, fl)
        synth.write_vcf(small, ped, pl[:args.ref_variants], fl[:args.ref_variants])
        gen_s = time.perf_counter() - t0
        out = {"variants": args.variants, "method": args.method, "vcf_bytes": os.path.getsize(big), "generate_s": gen_s}
        env = dict(os.environ, FAMSEQ_STATS="1")
        for _ in range(2):  # second run: warm page cache, CUDA context creation still included
            t0 = time.perf_counter()
            r = subprocess.run([ours, "vcf", "-vcfFile", big, "-pedFile", pp, "-output", os.path.join(td, "o.vcf"), "-method", args.method],
                               capture_output=True, text=True, env=env)
            wall = time.perf_counter() - t0
        stats = json.loads(r.stderr.strip().splitlines()[-1]) if r.stderr.strip() else {}
        out["ours"] = {"wall_s": wall, "variants_per_s": args.variants / wall, "stats": stats, "rc": r.returncode}
        if stats.get("total_s"):  # rate without the wait for CUDA context creation (fs_create), which a long run amortises
            out["ours"]["pipeline_variants_per_s"] = args.variants / max(1e-9, stats["total_s"] - stats.get("start_wait_s", 0.0))
        if os.path.exists(ref):
            t0 = time.perf_counter()
            r2 = subprocess.run([ref, "vcf", "-vcfFile", small, "-pedFile", pp, "-output", os.path.join(td, "r.vcf"), "-method", args.method],
                                capture_output=True, text=True)
            wall2 = time.perf_counter() - t0
            out["reference"] = {"wall_s": wall2, "variants": args.ref_variants, "variants_per_s": args.ref_variants / wall2, "rc": r2.returncode}
            subprocess.run([ours, "vcf", "-vcfFile", small, "-pedFile", pp, "-output", os.path.join(td, "o2.vcf"), "-method", args.method],
                           capture_output=True)
            out["identical_output"] = open(os.path.join(td, "o2.vcf")).read() == open(os.path.join(td, "r.vcf")).read()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
