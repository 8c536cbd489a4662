#!/usr/bin/env python
"""The reference's own CUDA build (oracle/_ref/FamSeqCuda: src/family.cu compiled unmodified for sm_100a, BN only,
N <= 19) against this repo's FamSeq on the same synthetic VCF, both on the GPU of this box.  Prints one JSON line:
seconds per variant of each (the reference: slope between two input sizes, so that start-up and CUDA context creation
cancel; this repo's binary: its own timers) and
whether the outputs agree (Phred numbers within 2e-6 relative, genotypes and text exactly).

    python tools/bench_ref_gpu.py [--pedigree ped14] [--small 8] [--large 40]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from famseq_b200 import synth  # noqa: E402


def numeric_equal(a, b):
    try:
        x, y = float(a), float(b)
    except ValueError:
        return False
    return x == y or (abs(x) < 1e-9 and abs(y) < 1e-9) or abs(x - y) <= 2e-6 * max(abs(x), abs(y))


def same(got, want):
    gl, wl = got.split("\n"), want.split("\n")
    if len(gl) != len(wl):
        return False
    for g, w in zip(gl, wl):
        if g == w:
            continue
        a, b = re.split(r"([\t:,])", g), re.split(r"([\t:,])", w)
        if len(a) != len(b) or not all(x == y or numeric_equal(x, y) for x, y in zip(a, b)):
            return False
    return True


def timed(cmd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, FAMSEQ_STATS="1"))
    return time.perf_counter() - t0, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pedigree", default="ped14")
    ap.add_argument("--small", type=int, default=8)
    ap.add_argument("--large", type=int, default=108)
    ap.add_argument("--ours-large", type=int, default=200000, help="records for this repo's binary (the slope needs seconds of work)")
    args = ap.parse_args()
    ours, ref = os.path.join(ROOT, "famseq_b200", "bin", "FamSeq"), os.path.join(ROOT, "oracle", "_ref", "FamSeqCuda")
    if not os.path.exists(ref):
        print(json.dumps({"unavailable": "oracle/_ref/FamSeqCuda is not built (make -C oracle refgpu, needs /root/reference)"}))
        return
    ped = synth.PEDIGREES[args.pedigree]()
    out = {"pedigree": args.pedigree, "members": len(ped.ids), "method": "BN (-method 1)"}
    with tempfile.TemporaryDirectory() as td:
        pp = os.path.join(td, "fam.ped")
        ped.write(pp)
        sizes = sorted({args.small, args.large, args.ours_large})
        pl, fl = synth.synth_pl(ped, max(sizes), seed=20261018 + 3)
        fl[:] = fl & 1  # autosomes only
        files = {}
        for n in sizes:
            files[n] = os.path.join(td, f"in_{n}.vcf")
            synth.write_vcf(files[n], ped, pl[:n], fl[:n])

        def run(binary, n, tag):
            o = os.path.join(td, f"out_{tag}_{n}.vcf")
            dt, r = timed([binary, "vcf", "-vcfFile", files[n], "-pedFile", pp, "-method", "1", "-output", o])
            return dt, r, o

        run(ref, args.small, "warm")  # page cache, driver
        ts, rs, _ = run(ref, args.small, "ref")
        tl, rl, ref_out = run(ref, args.large, "ref")
        out["reference_cuda"] = {"seconds": {str(args.small): ts, str(args.large): tl}, "rc": rl.returncode,
                                 "s_per_variant": (tl - ts) / (args.large - args.small)}
        out["reference_cuda"]["variants_per_s"] = 1.0 / max(1e-12, out["reference_cuda"]["s_per_variant"])
        run(ours, args.small, "warm")
        ts, rs, _ = run(ours, args.large, "ours")
        tl, rl, _ = run(ours, args.ours_large, "ours")
        # this binary reports its own timers: everything except the wait for CUDA context creation, which varies by
        # seconds from one process to the next on a GPU without persistence mode and would swamp the slope
        st = json.loads(rl.stderr.strip().splitlines()[-1])
        out["ours"] = {"seconds": {str(args.large): ts, str(args.ours_large): tl}, "rc": rl.returncode, "stats": st,
                       "s_per_variant": (st["total_s"] - st["start_wait_s"]) / args.ours_large,
                       "kernel_s_per_variant": st["kernel_ms"] * 1e-3 / args.ours_large}
        out["ours"]["variants_per_s"] = 1.0 / max(1e-12, out["ours"]["s_per_variant"])
        _, _, ours_out = run(ours, args.large, "cmp")
        out["outputs_agree"] = same(open(ours_out).read(), open(ref_out).read())
        out["speedup"] = out["ours"]["variants_per_s"] / out["reference_cuda"]["variants_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
