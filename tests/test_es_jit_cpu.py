"""CPU-only check of the generated Elston-Stewart code (famseq_b200/csrc/cuda/es_jit.cu).

The two message-program functions of the generated CUDA C++ (peel_a / peel_x: straight-line SSA code, no memory
traffic except the result row) are cut out of the source the engine hands back through fs_get_es_kernel, compiled for
the HOST behind a small shim (the round-to-nearest intrinsics become plain double operations, -ffp-contract=off; the
shared-reciprocal division is replaced by the IEEE division it is proven equal to) and run against the oracle.  They
must produce the same doubles.  This is test infrastructure: the product has no CPU compute path."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O

SHIM = r"""
#include <cstring>
typedef unsigned int u32; typedef unsigned long long u64; typedef long long i64; typedef unsigned char u8;
#define __device__
#define __forceinline__ inline
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, 8); return d; }
static inline void div3(double x0, double x1, double x2, double s, double &q0, double &q1, double &q2, bool) { q0 = x0 / s; q1 = x1 / s; q2 = x2 / s; }
struct Pipe {};                                   // the tile pipeline of the kernel: nothing to do for one variant on the host
static inline void pipeline_point(Pipe &) {}
static inline u8 call_genotype(double p0, double p1, double p2) {
    double big = -1.0; int arg = -1;
    if (big < p0) { big = p0; arg = 0; }
    if (big < p1) { big = p1; arg = 1; }
    if (big < p2) { big = p2; arg = 2; }
    return (u8)arg;
}
"""


def build_host_peel(tmp_path, src: str, S: int):
    start = src.index("// the message program with the autosomal rules")
    end = src.index('extern "C" __global__')
    args = ", ".join(f"lk[{c * 3 + g}]" for c in range(S) for g in range(3))
    wrapper = f"""
extern "C" int peel_host(int chrx, const double *lk, const double *pa, const double *pm, double *row, u8 *gt_row) {{
    Pipe pipe;
    return chrx ? peel_x(pipe, true, row, gt_row, pa[0], pa[1], pa[2], pm[0], pm[1], pm[2], {args})
                : peel_a(pipe, true, row, gt_row, pa[0], pa[1], pa[2], pm[0], pm[1], pm[2], {args});
}}
"""
    cpp, so = str(tmp_path / "peel.cpp"), str(tmp_path / "peel.so")
    open(cpp, "w").write(SHIM + src[start:end] + wrapper)
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", so, cpp], check=True)
    lib = ctypes.CDLL(so)
    lib.peel_host.restype = ctypes.c_int
    lib.peel_host.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5
    return lib


@pytest.mark.parametrize("name,cols", [("half_sibs", None), ("three_wives", None), ("ped14", None), ("ped14", [13, 2, 7, 0, 10, 5])])
def test_generated_peeling_code_reproduces_the_oracle(name, cols, tmp_path):
    ped = synth.PEDIGREES[name]()
    cols = ped.sequenced_cols() if cols is None else cols
    S, V = len(cols), 60
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, S + 1)]), V, seed=314, x_fraction=0.4)
    want = O.run(ped, cols, lk, fl, method=O.ES, lc=5.0)  # -LRC 5: the pedigree is always used
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1) as e:
        src, _ = e.es_kernel()
        priors = e.params.priors()
    lib = build_host_peel(tmp_path, src, S)
    checked = 0
    for v in range(V):
        known, chrx = int(fl[v]) & 1, (int(fl[v]) >> 1) & 1
        pa = np.ascontiguousarray(priors[1 if known else 0])
        pm = np.ascontiguousarray(priors[3 if known else 2]) if chrx else pa
        row, gt_row = np.zeros(S * 3), np.zeros(S, np.uint8)
        row_lk = np.ascontiguousarray(lk[v].reshape(-1))
        failed = lib.peel_host(chrx, row_lk.ctypes.data, pa.ctypes.data, pm.ctypes.data, row.ctypes.data, gt_row.ctypes.data)
        if want["status"][v]:
            continue  # the individual-only posterior failed first, or a row sum was zero
        assert not failed
        assert np.array_equal(row.reshape(S, 3), want["post"][v]), f"{name} variant {v}"
        assert np.array_equal(gt_row, want["gt"][v].astype(np.uint8))
        checked += 1
    assert checked > V // 2
