"""CPU-only check of the generated Gibbs sampler (famseq_b200/csrc/cuda/gibbs_jit.cu).

The kernel the engine generates for a pedigree is compiled for the HOST behind a shim that plays one CUDA thread
(threadIdx = blockIdx = 0, shared memory is a static buffer, ld.shared / ld.global.cg / red.global are plain memory
operations, the round-to-nearest intrinsics plain double operations with -ffp-contract=off, the Newton reciprocal
1/x) and run variant by variant against the oracle drawing from the same Philox stream.  Same chain, so the posteriors
agree to rounding (1e-9).  This is test infrastructure: the product has no CPU compute path."""
import ctypes
import re
import subprocess

import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O

SHIM = r"""
#include <cmath>
#include <cstring>
#include <algorithm>
typedef unsigned int u32; typedef unsigned long long u64; typedef long long i64; typedef unsigned char u8;
#define __device__
#define __global__
#define __constant__ static const
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n)
#define __shared__
struct Dim { int x; };
static const Dim threadIdx = {0}, blockIdx = {0}, gridDim = {1};
static unsigned char smem_raw[232448] __attribute__((aligned(128)));
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)((const unsigned char *)p - smem_raw); }
static inline double lds64(u32 a) { double v; std::memcpy(&v, smem_raw + a, 8); return v; }
static inline double __ldcg(const double *p) { return *p; }
static inline void atomicAdd(double *p, double v) { *p += v; }
static inline void atomicAdd(u64 *p, u64 v) { *p += v; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline u32 __double2uint_ru(double x) { const double c = std::ceil(x); return c <= 0.0 ? 0u : (c >= 4294967295.0 ? 4294967295u : (u32)c); }
static inline int __double2int_rd(double x) { return (int)std::floor(x); }
static inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, 8); return d; }
static inline int __double2hiint(double d) { long long x; std::memcpy(&x, &d, 8); return (int)(x >> 32); }
static inline double __hiloint2double(int hi, int lo) { long long x = ((long long)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &x, 8); return d; }
static inline u32 __umulhi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
static inline double newton_reciprocal(double s) { return 1.0 / s; }
using std::max;
"""

PHILOX_AND_CALL = r"""
static inline void philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 &o0, u32 &o1, u32 &o2, u32 &o3) {
    for (int r = 0; r < 10; r++) {
        const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}
static inline u8 call_genotype(double p0, double p1, double p2) {
    double big = -1.0; int arg = -1;
    if (big < p0) { big = p0; arg = 0; }
    if (big < p1) { big = p1; arg = 1; }
    if (big < p2) { big = p2; arg = 2; }
    return (u8)arg;
}
"""


def chain_stays_in_fast_range(ped, cols, lk, fl, burn, rep, seed, v_offset) -> bool:
    """Runs the oracle's chain of ONE variant on the same Philox stream and reports whether every weight sum it met is
    a positive normal number in [2^-963, 2^963) -- the range the generated kernel's straight-line code covers."""
    O.run(ped, cols, lk, fl, method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=seed, v_offset=v_offset)
    lo, hi = O.mcmc_sum_range()
    if lo > hi:  # the sampler never ran (individual-only failure or LRC gate)
        return True
    return O.FAST_SUM_LO <= lo and hi < O.FAST_SUM_HI


def lrc_gated(lk_row) -> bool:
    """Every sample certain (largest likelihood == row sum): the default -LRC 1 gate skips the sampler."""
    return bool(np.all(lk_row.max(-1) / lk_row.sum(-1) >= 1.0))


def build_host_kernel(tmp_path, src: str):
    head = src[:src.index("typedef unsigned int u32;")]  # "// generated ..." comments and the TB / NCOL macros
    head = re.sub(r"#define TB \d+", "#define TB 1", head)  # one thread plays the block: cooperative loops cover everything
    body = src[src.index("__constant__ u64 TAB_BITS"):]
    body = body.replace('extern "C" __global__ void', 'extern "C" void').replace("extern __shared__ __align__(16) unsigned char smem_raw[];", "")
    cpp, so = str(tmp_path / "gibbs.cpp"), str(tmp_path / "gibbs.so")
    open(cpp, "w").write(head + SHIM + PHILOX_AND_CALL + body)
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", so, cpp], check=True)
    lib = ctypes.CDLL(so)
    P, I64, U64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64
    lib.famseq_gibbs.restype = None
    lib.famseq_gibbs.argtypes = [P, P, P, P, P, P, I64, ctypes.c_int, ctypes.c_int, U64, I64, P, ctypes.c_int, P]
    return lib


@pytest.mark.parametrize("name,cols,V", [("trio", None, 40), ("half_sibs", None, 30), ("ped14", [13, 2, 7, 0, 10, 5], 20), ("ped40", None, 12)])
def test_generated_gibbs_code_reproduces_the_oracle(name, cols, V, tmp_path):
    ped = synth.PEDIGREES[name]()
    cols = ped.sequenced_cols() if cols is None else cols
    S, burn, rep, seed, v_offset = len(cols), 15, 120, 4242, 1000
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, S + 1)]), V, seed=27, x_fraction=0.4)
    want = O.run(ped, cols, lk, fl, method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=seed, v_offset=v_offset)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1) as e:
        src, _ = e.gibbs_kernel()
    lib = build_host_kernel(tmp_path, src)
    scratch = np.zeros(6 * ped.n * 1024 + 1024)  # block-private rows of 3 x TB doubles, TB <= 1024
    checked = handed_back = 0
    for v in range(V):
        row_lk = np.ascontiguousarray(lk[v])
        flag = np.array([fl[v]], np.uint8)
        post, single = np.zeros((S, 3)), np.zeros((S, 3))
        gt, status = np.zeros(S, np.uint8), np.full(1, 9, np.uint8)
        lib.famseq_gibbs(row_lk.ctypes.data, flag.ctypes.data, post.ctypes.data, single.ctypes.data, gt.ctypes.data, status.ctypes.data,
                         1, burn, rep, seed, v_offset + v, scratch.ctypes.data, 1, None)
        if status[0] == 2:  # handed back to the table-driven kernel: legitimate only if a weight sum left the fast range
            assert not chain_stays_in_fast_range(ped, cols, lk[v:v + 1], fl[v:v + 1], burn, rep, seed, v_offset + v), f"{name} variant {v}"
            handed_back += 1
            continue
        assert status[0] == want["status"][v], f"{name} variant {v}"
        if status[0]:
            continue
        assert np.array_equal(single, want["single"][v])
        assert np.allclose(post, want["post"][v], rtol=1e-9, atol=0), f"{name} variant {v}"
        assert np.array_equal(gt, want["gt"][v].astype(np.uint8))
        checked += 1
    assert checked >= V // 2


def test_status_2_is_raised_exactly_when_a_weight_sum_leaves_the_fast_range(tmp_path):
    """Likelihoods over a wide exponent range: some chains meet weight sums that are zero, subnormal or tiny.  The
    generated code must flag exactly those (status 2, redone by the table-driven kernel) and agree with the oracle on
    every other chain."""
    ped = synth.half_sibs()
    cols = ped.sequenced_cols()
    S, V, burn, rep, seed, v_offset = len(cols), 60, 5, 40, 99, 7
    rng = np.random.default_rng(3)
    lk = rng.random((V, S, 3)) * np.exp2(rng.integers(-330, 1, (V, S, 3)).astype(np.float64))
    lk[rng.random((V, S, 3)) < 0.05] = 0.0
    fl = (rng.integers(0, 4, V)).astype(np.uint8)
    want = O.run(ped, cols, lk, fl, method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=seed, v_offset=v_offset)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1) as e:
        src, _ = e.gibbs_kernel()
    lib = build_host_kernel(tmp_path, src)
    scratch = np.zeros(6 * ped.n * 1024 + 1024)
    flagged = clean = 0
    for v in range(V):
        row_lk, flag = np.ascontiguousarray(lk[v]), np.array([fl[v]], np.uint8)
        post, single = np.zeros((S, 3)), np.zeros((S, 3))
        gt, status = np.zeros(S, np.uint8), np.full(1, 9, np.uint8)
        lib.famseq_gibbs(row_lk.ctypes.data, flag.ctypes.data, post.ctypes.data, single.ctypes.data, gt.ctypes.data, status.ctypes.data,
                         1, burn, rep, seed, v_offset + v, scratch.ctypes.data, 1, None)
        gated = want["status"][v] == 1 and not np.any(want["single"][v])  # failed before the sampler started
        in_range = chain_stays_in_fast_range(ped, cols, lk[v:v + 1], fl[v:v + 1], burn, rep, seed, v_offset + v)
        if status[0] == 2:
            assert not in_range, f"variant {v} was handed back although its sums stay in the fast range"
            flagged += 1
            continue
        assert gated or in_range or lrc_gated(lk[v]), f"variant {v} left the fast range without being handed back"
        assert status[0] == want["status"][v], f"variant {v}"
        if status[0] == 0:
            assert np.allclose(post, want["post"][v], rtol=1e-9, atol=0), f"variant {v}"
            clean += 1
    assert flagged > 0 and clean > 0, (flagged, clean)
