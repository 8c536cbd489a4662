// common.cuh -- pieces shared by the three method kernels: the per-run constant block, the
// individual-only posterior + LRC gate every method starts with (family.cpp:1405-1499, :767-789) and
// the arg-max of get_postRlt (family.cpp:636-665).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace famseq {

constexpr int FS_MAX_MEMBERS = 128;
enum : int { K_TAB_AUTO = 0, K_TAB_XF = 1, K_TAB_XM = 2 }; // same values as host TableKind

// Passed by value as a __grid_constant__ kernel parameter: lives in the constant bank, so table
// entries become direct constant operands of the FP64 instructions.
struct RunConstants {
    double tab[3][27];  // [TAB_AUTO|TAB_XF|TAB_XM][g*9 + mother*3 + father]
    double prior[4][3]; // genoProbN, genoProbK, genoProbXN, genoProbXK
    double lrc;         // -LRC
    int32_t n, s;
    // unseq_fail[flags & 3]: some UNSEQUENCED member's prior row sums to <= 0 for this (Known, chrX)
    // combination, i.e. calPostProbSingle would return false for every such variant.
    uint8_t unseq_fail[4];
    uint8_t col_male[FS_MAX_MEMBERS]; // gender == 1 of the member behind each input column
    uint8_t pad[4];
};

struct BatchPtrs {
    const double *lk;     // [V][S][3]
    const uint8_t *flags; // [V] or nullptr
    double *post;         // [V][S][3]
    double *single;       // [V][S][3]; nullptr = not wanted (only the nuclear-family kernel sees that: the engine gives
                          // every other kernel a scratch buffer)
    uint8_t *gt;          // [V][S]
    uint8_t *status;      // [V]
    int64_t V;
    // Compact input (fs_run_pl): integer Phred-scaled likelihoods [V][S][3], decoded as lut[pl] = pow(10, -pl/10) with
    // the 65 536-entry table the engine built on the host with libm (file.cpp:588-590).  The nuclear-family kernel
    // gathers from the table itself; for every other kernel the engine first expands pl into an FP64 scratch `lk`.
    const uint16_t *pl = nullptr;
    const double *lut = nullptr;
};

// Prior pair of one variant: `a` for females and for everybody on autosomes, `m` for males.
struct VariantPriors {
    double a[3], m[3];
};

__device__ __forceinline__ VariantPriors select_priors(const RunConstants &C, unsigned flag) {
    const int known = flag & 1, chrx = (flag >> 1) & 1;
    VariantPriors p;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        p.a[g] = known ? C.prior[1][g] : C.prior[0][g];
        p.m[g] = chrx ? (known ? C.prior[3][g] : C.prior[2][g]) : p.a[g];
    }
    return p;
}

// a[g] for a runtime g without turning the array into local memory
__device__ __forceinline__ double pick3(const double (&a)[3], int g) { return g == 0 ? a[0] : (g == 1 ? a[1] : a[2]); }

// ---- x[0..2] / s, correctly rounded -------------------------------------------------------------------
// r = RN(1/s); q0 = RN(x r); one residual correction q = RN(q0 + (x - s q0) r), the residual being exact in an
// FMA.  With a correctly rounded reciprocal this is RN(x/s) (Markstein's theorem) provided nothing over- or
// underflows, which the exponent guard ensures: s in [2^-900, 2^900], x = 0 or x in [2^-900, 2^900].  Checked
// against the IEEE divide on 2e9 adversarial operand pairs on the host and bit for bit by the GPU parity tests.
// The guard runs on the integer pipe (the FP64 pipe is the scarce one): a positive normal divisor with exponent in
// [-900, 900], every dividend either +-0 or of magnitude in the same range.
__device__ __forceinline__ bool safe_dividend(double x) {
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu, lo = (unsigned)__double2loint(x);
    return (hi - ((1023u - 900u) << 20)) <= (1800u << 20) || (hi | lo) == 0u;
}
__device__ __forceinline__ void div3(double x0, double x1, double x2, double s, double &q0, double &q1, double &q2) {
    const bool fast = ((unsigned)__double2hiint(s) - ((1023u - 900u) << 20)) <= (1800u << 20) // positive (sign bit clear)
                      & safe_dividend(x0) & safe_dividend(x1) & safe_dividend(x2);
    if (fast) {
        const double r = __drcp_rn(s);
        const double a = __dmul_rn(x0, r), b = __dmul_rn(x1, r), c = __dmul_rn(x2, r);
        q0 = __fma_rn(__fma_rn(-s, a, x0), r, a);
        q1 = __fma_rn(__fma_rn(-s, b, x1), r, b);
        q2 = __fma_rn(__fma_rn(-s, c, x2), r, c);
    } else {
        q0 = x0 / s;
        q1 = x1 / s;
        q2 = x2 / s;
    }
}

// LRC gate of one sample (family.cpp:767-789): true when big/sum < lrc.  For the default -LRC 1 and a row without
// negative entries (sign bits clear), big/ls < 1 <=> big < ls (the quotient of two distinct adjacent doubles already
// rounds below 1; 0/0, inf/inf and NaN rows compare false both ways), which spares the division; any other -LRC
// value divides.
__device__ __forceinline__ bool lrc_wants_pedigree(double lrc, double l0, double l1, double l2, double big, double ls) {
    const bool no_negative = (__double2hiint(l0) | __double2hiint(l1) | __double2hiint(l2)) >= 0;
    if (lrc == 1.0 && no_negative) return big < ls;
    return big / ls < lrc;
}

// get_postRlt: strict '<' starting from -1, so the first maximum wins and NaN rows give -1 (255).
__device__ __forceinline__ uint8_t call_genotype(double p0, double p1, double p2) {
    double big = -1.0;
    int arg = -1;
    if (big < p0) { big = p0; arg = 0; }
    if (big < p1) { big = p1; arg = 1; }
    if (big < p2) { big = p2; arg = 2; }
    return (uint8_t)arg;
}

} // namespace famseq
