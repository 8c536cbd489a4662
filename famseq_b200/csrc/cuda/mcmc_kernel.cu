// mcmc_kernel.cu -- single-site Gibbs sampler with Rao-Blackwellised marginals, one chain (= one
// variant) per thread, for sm_100a.  Replaces family::calPostProbMCMC + estGenoProb
// (src/family.cpp:1932-2096, :2098-2299).
//
//   * sweep order, full conditionals, the 1e6 pre-scale, the `rd < w0` / `rd > 1 - w2` draw rule, the
//     accumulation of the normalised weights (not of indicator counts) and the chrX quirk (females get
//     no children factor, family.cpp:2230-2257) follow the reference;
//   * the random numbers do not: libc rand() (one global, never seeded stream) is replaced by a
//     counter-based Philox4x32-10 stream keyed by (seed, global variant index), so a variant's chain is
//     the same whatever the batch split or GPU count.  Draw t of a variant is word t%4 of Philox block
//     t/4; the first N draws initialise the genotypes (u % 3), then one draw per member per sweep;
//   * per-chain state: genotype vector packed 2 bits/member in two 64-bit registers; likelihood rows and
//     the 3N accumulators in shared memory as [member][g][thread] (conflict-free); the three 27-entry
//     transmission tables are replicated per lane in shared memory so that data-dependent look-ups never
//     bank-conflict.
// The kernel is bound by shared-memory bandwidth and FP64 issue, not HBM.
#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"

namespace famseq {

namespace {

struct Philox {
    uint32_t c0, c2, c3, k0, k1; // counter word 1 is always 0
    uint32_t b0, b1, b2, b3;
    int have;

    __device__ __forceinline__ void start(uint64_t seed, uint64_t gv) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        c0 = 0;
        c2 = (uint32_t)gv;
        c3 = (uint32_t)(gv >> 32);
        have = 0;
    }
    __device__ __forceinline__ void refill() {
        uint32_t x0 = c0, x1 = 0, x2 = c2, x3 = c3, ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t n0 = hi1 ^ x1 ^ ka, n2 = hi0 ^ x3 ^ kb;
            x0 = n0;
            x1 = lo1;
            x2 = n2;
            x3 = lo0;
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        b0 = x0;
        b1 = x1;
        b2 = x2;
        b3 = x3;
        c0++;
        have = 4;
    }
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) refill();
        const uint32_t r = b0;
        b0 = b1;
        b1 = b2;
        b2 = b3;
        have--;
        return r;
    }
};

struct Genotypes {
    uint64_t lo, hi;
    __device__ __forceinline__ int get(int i) const { return (int)(((i < 32 ? lo : hi) >> (2 * (i & 31))) & 3u); }
    __device__ __forceinline__ void set(int i, int g) {
        const uint64_t m = 3ull << (2 * (i & 31)), v = (uint64_t)g << (2 * (i & 31));
        if (i < 32)
            lo = (lo & ~m) | v;
        else
            hi = (hi & ~m) | v;
    }
};

template <int TB>
__global__ void __launch_bounds__(TB) mcmc_kernel(const __grid_constant__ McmcParams P, const BatchPtrs B, int burn, int rep,
                                                  uint64_t seed, int64_t v_offset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RunConstants &C = P.C;
    const McmcPlan &pl = P.plan;
    const int N = pl.n, S = C.s;
    double *s_tab = reinterpret_cast<double *>(smem_raw); // [3*27][32] lane-replicated transmission tables
    double *s_w = s_tab + 81 * 32;                         // [N][3][TB] own factor: 1e6*prior*lk or 1e6*lk
    double *s_acc = s_w + N * 3 * TB;                      // [N][3][TB]

    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < 81 * 32; e += TB) s_tab[e] = C.tab[(e >> 5) / 27][(e >> 5) % 27];
    __syncthreads();
    const double *tab = s_tab + lane;

    const int64_t v = (int64_t)blockIdx.x * TB + tid;
    if (v >= B.V) return;
    const unsigned flag = B.flags ? B.flags[v] : 0u;
    const bool chrx = (flag >> 1) & 1u;
    const VariantPriors pr = select_priors(C, flag);
    const double *lkv = B.lk + v * S * 3;
    double *gp = B.post + v * S * 3, *gs = B.single + v * S * 3;
    uint8_t *gg = B.gt + v * S;

    // individual-only posterior + LRC gate (family.cpp:1940-1971)
    bool failed = C.unseq_fail[flag & 3u] != 0;
    bool pedigree_needed = false;
    for (int c = 0; c < S; c++) {
        const double l0 = lkv[c * 3], l1 = lkv[c * 3 + 1], l2 = lkv[c * 3 + 2];
        const bool male = C.col_male[c] != 0;
        const double r0 = l0 * (male ? pr.m[0] : pr.a[0]);
        const double r1 = l1 * (male ? pr.m[1] : pr.a[1]);
        const double r2 = l2 * (male ? pr.m[2] : pr.a[2]);
        const double rs = __dadd_rn(__dadd_rn(r0, r1), r2);
        if (rs <= 0.0) failed = true;
        gs[c * 3] = r0 / rs;
        gs[c * 3 + 1] = r1 / rs;
        gs[c * 3 + 2] = r2 / rs;
        double big = 0.0;
        if (big < l0) big = l0;
        if (big < l1) big = l1;
        if (big < l2) big = l2;
        const double ls = __dadd_rn(__dadd_rn(l0, l1), l2);
        if (big / ls < C.lrc) pedigree_needed = true;
    }
    if (failed) {
        for (int k = 0; k < S * 3; k++) gp[k] = gs[k] = 0.0;
        for (int c = 0; c < S; c++) gg[c] = 255;
        B.status[v] = 1;
        return;
    }
    if (!pedigree_needed) { // family.cpp:1973-2058: FPP := individual-only posterior
        for (int c = 0; c < S; c++) {
            const double p0 = gs[c * 3], p1 = gs[c * 3 + 1], p2 = gs[c * 3 + 2];
            gp[c * 3] = p0;
            gp[c * 3 + 1] = p1;
            gp[c * 3 + 2] = p2;
            gg[c] = call_genotype(p0, p1, p2);
        }
        B.status[v] = 0;
        return;
    }

    // own factors: ((1e6 * prior) * lk) for founders, (1e6 * lk) for the others (family.cpp:2115-2126)
    for (int i = 0; i < N; i++) {
        const int col = pl.col[i];
        const bool founder = pl.mother[i] < 0, male = pl.male[i] != 0;
#pragma unroll
        for (int g = 0; g < 3; g++) {
            const double lk = col >= 0 ? lkv[col * 3 + g] : 1.0;
            const double base = founder ? 1000000.0 * (male ? pr.m[g] : pr.a[g]) : 1000000.0;
            s_w[(i * 3 + g) * TB + tid] = base * lk;
            s_acc[(i * 3 + g) * TB + tid] = 0.0;
        }
    }

    Philox rng;
    rng.start(seed, (uint64_t)(v_offset + v));
    Genotypes cur{0, 0};
    for (int i = 0; i < N; i++) cur.set(i, (int)(rng.next() % 3u)); // family.cpp:2063-2067

    const int total_sweeps = burn + rep;
    for (int sweep = 0; sweep < total_sweeps; sweep++) {
        const bool sampling = sweep >= burn;
        for (int i = 0; i < N; i++) {
            const int mo = pl.mother[i];
            const bool male = pl.male[i] != 0;
            double w0 = s_w[(i * 3 + 0) * TB + tid];
            double w1 = s_w[(i * 3 + 1) * TB + tid];
            double w2 = s_w[(i * 3 + 2) * TB + tid];
            if (mo >= 0) {
                const int kind = chrx ? (male ? K_TAB_XM : K_TAB_XF) : K_TAB_AUTO;
                const double *t = tab + (kind * 27 + cur.get(mo) * 3 + cur.get(pl.father[i])) * 32;
                w0 *= t[0];
                w1 *= t[9 * 32];
                w2 *= t[18 * 32];
            }
            if (!chrx || male) { // chrX: only males get the children factor (reference quirk)
                const int lb = pl.link_begin[i], le = pl.link_begin[i + 1];
                for (int k = lb; k < le; k++) {
                    const int c = pl.link_child[k];
                    const int kind = chrx ? (pl.male[c] ? K_TAB_XM : K_TAB_XF) : K_TAB_AUTO;
                    const int other = cur.get(pl.link_other[k]);
                    // Pr(child genotype | mother, father): this member sits in the father slot when male
                    const double *t = tab + (kind * 27 + cur.get(c) * 9 + (male ? other * 3 : other)) * 32;
                    const int step = male ? 32 : 96;
                    w0 *= t[0];
                    w1 *= t[step];
                    w2 *= t[2 * step];
                }
            }
            const double sum = (w0 + w1) + w2;
            if (sum <= 0.0) {
                w0 = w1 = w2 = 0.0;
            } else {
                const double inv = 1.0 / sum;
                w0 *= inv;
                w1 *= inv;
                w2 *= inv;
            }
            const double rd = ((double)rng.next() + 0.5) * (1.0 / 4294967296.0);
            const int g = (rd < w0) ? 0 : ((rd > (1.0 - w2)) ? 2 : 1);
            cur.set(i, g);
            if (sampling) {
                s_acc[(i * 3 + 0) * TB + tid] += w0;
                s_acc[(i * 3 + 1) * TB + tid] += w1;
                s_acc[(i * 3 + 2) * TB + tid] += w2;
            }
        }
    }

    // postProb = genoFry / numRep, not renormalised; a row summing to <= 0 fails the variant (family.cpp:2082-2092)
    const double nrep = (double)rep;
    for (int i = 0; i < N; i++) {
        const double p0 = s_acc[(i * 3 + 0) * TB + tid] / nrep;
        const double p1 = s_acc[(i * 3 + 1) * TB + tid] / nrep;
        const double p2 = s_acc[(i * 3 + 2) * TB + tid] / nrep;
        if (__dadd_rn(__dadd_rn(p0, p1), p2) <= 0.0) failed = true;
        const int c = pl.col[i];
        if (c >= 0) {
            gp[c * 3] = p0;
            gp[c * 3 + 1] = p1;
            gp[c * 3 + 2] = p2;
            gg[c] = call_genotype(p0, p1, p2);
        }
    }
    if (failed) {
        for (int k = 0; k < S * 3; k++) gp[k] = gs[k] = 0.0;
        for (int c = 0; c < S; c++) gg[c] = 255;
    }
    B.status[v] = failed ? 1 : 0;
}

} // namespace

size_t mcmc_smem_bytes(const McmcParams &P, int tb) {
    return (size_t)(81 * 32 + 2 * P.plan.n * 3 * tb) * sizeof(double);
}

// Block size that keeps the most chains resident per SM.
int mcmc_pick_block(const McmcParams &P, size_t smem_limit, size_t smem_per_sm) {
    const int candidates[] = {128, 96, 64, 32};
    int best = 0, best_threads = 0;
    for (int tb : candidates) {
        const size_t need = mcmc_smem_bytes(P, tb) + 1024; // 1 KB per-block reservation
        if (need > smem_limit + 1024) continue;
        const int blocks = (int)std::min<size_t>(32, smem_per_sm / need);
        const int threads = std::min(2048, blocks * tb);
        if (threads > best_threads) {
            best_threads = threads;
            best = tb;
        }
    }
    return best;
}

cudaError_t launch_mcmc(const McmcParams &P, const BatchPtrs &B, int tb, int burn, int rep, uint64_t seed,
                        int64_t v_offset, cudaStream_t stream) {
    if (B.V <= 0) return cudaSuccess;
    const size_t smem = mcmc_smem_bytes(P, tb);
    const unsigned grid = (unsigned)((B.V + tb - 1) / tb);
    cudaError_t rc;
#define FS_LAUNCH_MCMC(TBV)                                                                                       \
    case TBV:                                                                                                      \
        rc = cudaFuncSetAttribute(mcmc_kernel<TBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        if (rc != cudaSuccess) return rc;                                                                          \
        mcmc_kernel<TBV><<<grid, TBV, smem, stream>>>(P, B, burn, rep, seed, v_offset);                            \
        break;
    switch (tb) {
        FS_LAUNCH_MCMC(128)
        FS_LAUNCH_MCMC(96)
        FS_LAUNCH_MCMC(64)
        FS_LAUNCH_MCMC(32)
    default:
        return cudaErrorInvalidValue;
    }
#undef FS_LAUNCH_MCMC
    return cudaGetLastError();
}

} // namespace famseq
