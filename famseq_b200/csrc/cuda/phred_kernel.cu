// phred_kernel.cu -- Phred encoding of posterior probabilities on the device (output half).
//
// The reference's drivers print every posterior as fabs(-10 * log10(p)) with ostream's default formatting -- "%g", six
// significant digits -- or 99999 when the value is +inf (file.cpp:702-761, :938-997, :1814-1873).  A caller that only
// wants that text does not need the 8-byte double: this kernel turns p into the six decimal digits and the decimal
// exponent the text is made of, 4 bytes per value, halving what has to cross PCIe (fs_run_pl_phred).
//
// Exactness: the host text is made from glibc's log10, this one from CUDA's; both are within a few ulps of the true
// logarithm, so the two doubles v = |-10 log10 p| can differ in the last bits, and their six-digit roundings differ
// only when v sits next to a rounding boundary (..5 in the seventh digit).  The kernel measures the distance of v from
// the nearest boundary; closer than 3e-8 in units of the sixth digit -- more than a hundred ulps of v -- it does not
// decide: the value goes to the caller as an exception with its exact double (fs_phred_fix), and the host formats it
// the reference's way.  Non-finite, negative and > 1 inputs go the same way.  About 6 values in 10^8 take that path.
#include <cstdint>

#include "../../../include/famseq_b200.h"
#include "kernels.hpp"

namespace famseq {

namespace {

__constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22}; // all exact doubles

// Returns the packed code of p; *fix = true when the host has to format this value itself.
__device__ __forceinline__ uint32_t phred_pack(double p, bool *fix) {
    *fix = false;
    if (p == 0.0) return FS_PHRED_INF;  // -10 log10(0) = +inf: the reference prints 99999
    if (p == 1.0) return FS_PHRED_ZERO; // -10 log10(1) = -0: fabs -> 0
    if (!(p > 0.0 && p < 1.0)) {        // NaN, negative, above 1: no business here, the host decides
        *fix = true;
        return FS_PHRED_FIX;
    }
    const double v = fabs(-10.0 * log10(p)); // in (4.8e-16, 3240)
    int k = (int)floor(log10(v));            // decimal exponent of the leading digit, possibly off by one next to a power of ten
    k = max(-17, min(4, k));
    double s = v * kPow10[5 - k];            // six integer digits; the power of ten is exact, one rounding in the product
    if (s < 1e5) {
        k--;
        s = v * kPow10[5 - k];
    } else if (s >= 1e6) {
        k++;
        s = v * kPow10[5 - k];
    }
    const double whole = floor(s), frac = s - whole; // exact
    if (!(s >= 1e5 && s < 1e6) || fabs(frac - 0.5) < 3e-8) {
        *fix = true;
        return FS_PHRED_FIX;
    }
    uint32_t m = (uint32_t)whole + (frac > 0.5 ? 1u : 0u);
    if (m == 1000000u) {
        m = 100000u;
        k++;
    }
    return m | ((uint32_t)(k + 32) << 20);
}

__global__ void __launch_bounds__(256) phred_pack_kernel(const double *__restrict__ p, uint32_t *__restrict__ out, int64_t n, int64_t index0,
                                                          fs_phred_fix *__restrict__ fixes, int64_t capacity, unsigned long long *__restrict__ n_fixes) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const double x = p[k];
        bool fix;
        out[k] = phred_pack(x, &fix);
        if (fix) {
            const unsigned long long slot = atomicAdd(n_fixes, 1ull);
            if ((int64_t)slot < capacity) {
                fixes[slot].index = index0 + k;
                fixes[slot].p = x;
            }
        }
    }
}

} // namespace

cudaError_t launch_phred_pack(const double *p, uint32_t *out, int64_t n, int64_t index0, fs_phred_fix *fixes, int64_t capacity,
                              unsigned long long *n_fixes, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    phred_pack_kernel<<<grid, 256, 0, stream>>>(p, out, n, index0, fixes, capacity, n_fixes);
    return cudaGetLastError();
}

} // namespace famseq
