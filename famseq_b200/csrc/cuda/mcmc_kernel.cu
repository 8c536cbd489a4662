// mcmc_kernel.cu -- single-site Gibbs sampler with Rao-Blackwellised marginals, one chain (= one variant)
// per thread, for sm_100a.  Replaces family::calPostProbMCMC + estGenoProb (src/family.cpp:1932-2096,
// :2098-2299).
//
// What follows the reference: the ped-order scan, the full conditionals, the 1e6 pre-scale, the
// `rd < w0` / `rd > 1 - w2` draw rule, accumulation of the normalised weights (not of indicator counts),
// `postProb = sum / numRep` without renormalisation, and the chrX quirk (females get no children factor,
// family.cpp:2230-2257).
// What does not: libc rand() (one global, never seeded stream) is replaced by counter-based Philox4x32-10.
// The random word of member i in sweep s of site v is word i%4 of Philox(counter = (s, i/4, v_lo, v_hi),
// key = seed); s = 0 initialises the genotypes (word % 3), s >= 1 are the sweeps, whose uniform is
// rd = ((word >> 1) + 0.5) * 2^-31 (31 bits like rand() / RAND_MAX; it lets the generated kernel of gibbs_jit.cu decide a
// draw by integer comparisons that are exactly equivalent to rd < w0, rd > 1 - w2).  A value depends on
// (seed, global site, sweep, member) only: any batch split or GPU count gives the same bytes.
//
// A chain is strictly sequential, so the kernel is bound by instruction issue of the few warps whose chain
// state fits on an SM, not by FP64 throughput or HBM (profiles/*mcmc*): the design minimises instructions per
// Gibbs step and on-chip bytes per chain.
//   * all threads of a warp work on the same member: control flow is uniform, pedigree metadata are one
//     uniform constant-bank word per member and per parent-child link (host/mcmc_plan.cpp);
//   * the 3N accumulators live in a global scratch ([block][member][g][thread], thread-private, coalesced) and
//     are updated with fire-and-forget red.global.add.f64 -- three instructions per step, off the dependency
//     chain; the 3N own factors live in shared memory as [member][g][thread] when a full SM's worth of chains
//     fits (small pedigrees), otherwise in the same scratch, fetched one member ahead through L2, so that the
//     number of chains in flight is set by registers (24 warps/SM) and not by shared memory (7 warps/SM for 40
//     members);
//   * the draw is decided on un-normalised weights (rd*sum < w0, rd*sum > sum - w2), so the reciprocal
//     (MUFU seed + two Newton steps) is off the critical path;
//   * the transmission tables are replicated 16 times in shared memory so that the data-dependent look-ups
//     of a half-warp hit 16 different bank pairs;
//   * persistent blocks loop over tiles of TB variants.
#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"

namespace famseq {

namespace {

constexpr int kCopies = 16;      // replicas of every table entry
constexpr int kEntries = 81 + 27; // 3 x 27 transmission entries followed by 27 ones
constexpr int kOnes = 81;

__device__ __forceinline__ double lds64(uint32_t shared_addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(shared_addr));
    return v;
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o0 = c0;
    o1 = c1;
    o2 = c2;
    o3 = c3;
}

// genotype vector, 2 bits per member in GW 64-bit words (GW = 2: up to 64 members, GW = 4: up to 128); the member index
// is warp-uniform
template <int GW> struct Genotypes {
    uint64_t w[GW];
    __device__ __forceinline__ uint64_t word(int i) const {
        if (GW == 2) return i < 32 ? w[0] : w[1];
        return i < 64 ? (i < 32 ? w[0] : w[1]) : (i < 96 ? w[2] : w[GW - 1]);
    }
    __device__ __forceinline__ int get(int i) const { return (int)((word(i) >> (2 * (i & 31))) & 3u); }
    __device__ __forceinline__ void set(int i, int g) {
        const uint64_t m = 3ull << (2 * (i & 31)), v = (uint64_t)g << (2 * (i & 31));
#pragma unroll
        for (int k = 0; k < GW; k++)
            if ((i >> 5) == k) w[k] = (w[k] & ~m) | v;
    }
};

// 1/s to ~1 ulp: hardware seed (MUFU.RCP64H) + two Newton steps; outside the safe range the IEEE divide.
__device__ __forceinline__ double fast_reciprocal(double s) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(s));
    double e = fma(-s, x, 1.0);
    x = fma(x, e, x);
    e = fma(-s, x, 1.0);
    x = fma(x, e, x);
    // rare: not a positive normal number with exponent in [-963, 963) (integer-pipe test on the high word)
    if (!(((unsigned)__double2hiint(s) - 0x03c00000u) < 0x78600000u)) x = 1.0 / s;
    return x;
}

// WGLOBAL = false: own factors in shared memory (small pedigrees: every chain of a full SM fits).
// WGLOBAL = true : own factors in the global scratch next to the accumulators, read one member ahead through L2;
//                  shared memory then only holds the tables and the number of chains per SM is set by registers.
template <int TB, bool WGLOBAL, int GW>
__global__ void __launch_bounds__(TB) mcmc_kernel(const __grid_constant__ McmcParams P, const BatchPtrs B, int burn, int rep,
                                                  uint64_t seed, int64_t v_offset, double *__restrict__ scratch, int n_tiles, int fixup,
                                                  unsigned long long *fixup_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RunConstants &C = P.C;
    const McmcPlan &pl = P.plan;
    const int N = pl.n, S = C.s;
    double *s_tab = reinterpret_cast<double *>(smem_raw); // [108][16]
    double *s_w = s_tab + kEntries * kCopies;              // [N][3][TB] own factor: (1e6*prior)*lk or 1e6*lk

    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < kEntries * kCopies; e += TB) {
        const int entry = e / kCopies;
        s_tab[e] = entry < 81 ? C.tab[entry / 27][entry % 27] : 1.0;
    }
    __syncthreads();
    const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(s_tab) + (uint32_t)(lane & (kCopies - 1)) * 8u;
    const uint32_t w_addr = (uint32_t)__cvta_generic_to_shared(s_w) + (uint32_t)tid * 8u;
    double *acc = scratch + (size_t)blockIdx.x * ((size_t)N * 3 * TB * (WGLOBAL ? 2 : 1)) + tid; // [member][g][TB], thread-private
    double *wg = acc + (size_t)N * 3 * TB;                                                       // WGLOBAL: own factors
    double *w = WGLOBAL ? wg : s_w + tid;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t v = (int64_t)tile * TB + tid;
        if (v >= B.V) continue;
        if (fixup) { // second pass after the specialised kernel: only what it left (status 2)
            if (B.status[v] != 2) continue;
            if (fixup_count) atomicAdd(fixup_count, 1ull);
        }
        const unsigned flag = B.flags ? B.flags[v] : 0u;
        const bool chrx = (flag >> 1) & 1u;
        const VariantPriors pr = select_priors(C, flag);
        const double *lkv = B.lk + v * S * 3;
        double *gp = B.post + v * S * 3, *gs = B.single + v * S * 3;
        uint8_t *gg = B.gt + v * S;

        // ---- individual-only posterior + LRC gate (family.cpp:1940-1971) ----------------------------------------
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
            continue;
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
            continue;
        }

        // ---- chain state: own factors ((1e6 * prior) * lk for founders, 1e6 * lk otherwise, family.cpp:2115-2126)
        for (int i = 0; i < N; i++) {
            const uint32_t d = pl.member[i];
            const int col = pl.col[i];
            const bool founder = (d >> 14) & 1u, male = (d >> 15) & 1u;
#pragma unroll
            for (int g = 0; g < 3; g++) {
                const double lk = col >= 0 ? lkv[col * 3 + g] : 1.0;
                const double base = founder ? 1000000.0 * (male ? pick3(pr.m, g) : pick3(pr.a, g)) : 1000000.0;
                w[(i * 3 + g) * TB] = base * lk;
                acc[(size_t)((i * 3 + g) * TB)] = 0.0;
            }
        }
        const uint64_t gv = (uint64_t)(v_offset + v);
        const uint32_t gv_lo = (uint32_t)gv, gv_hi = (uint32_t)(gv >> 32);
        uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        Genotypes<GW> cur;
#pragma unroll
        for (int k = 0; k < GW; k++) cur.w[k] = 0;
        for (int i = 0; i < N; i++) { // family.cpp:2063-2067
            if ((i & 3) == 0) philox4x32_10(0u, (uint32_t)(i >> 2), gv_lo, gv_hi, k0, k1, r0, r1, r2, r3);
            const int q = i & 3;
            const uint32_t u = q == 0 ? r0 : (q == 1 ? r1 : (q == 2 ? r2 : r3));
            cur.set(i, (int)(u % 3u));
        }

        // ---- sweeps ---------------------------------------------------------------------------------------------
        // The step is written branch-light: founders and missing children multiply by a row of ones, the first two
        // children are handled in line, shared memory is addressed with 32-bit shared-window addresses.
        asm volatile("" ::: "memory");
        double nw0 = 0.0, nw1 = 0.0, nw2 = 0.0; // WGLOBAL: own factors of the next member, fetched one step ahead
        if constexpr (WGLOBAL) {
            nw0 = __ldcg(wg);
            nw1 = __ldcg(wg + TB);
            nw2 = __ldcg(wg + 2 * TB);
        }
        const int total_sweeps = burn + rep;
        for (int sweep = 1; sweep <= total_sweeps; sweep++) {
            const bool sampling = sweep > burn;
            for (int i = 0; i < N; i++) {
                const uint32_t d = pl.member[i];
                const bool male = (d >> 15) & 1u, founder = (d >> 14) & 1u;
                double w0, w1, w2;
                if constexpr (WGLOBAL) {
                    w0 = nw0;
                    w1 = nw1;
                    w2 = nw2;
                    const double *nx = wg + (size_t)(((i + 1 < N) ? i + 1 : 0) * 3 * TB);
                    nw0 = __ldcg(nx);
                    nw1 = __ldcg(nx + TB);
                    nw2 = __ldcg(nx + 2 * TB);
                } else {
                    const uint32_t wa = w_addr + (uint32_t)i * (3u * TB * 8u);
                    w0 = lds64(wa);
                    w1 = lds64(wa + TB * 8);
                    w2 = lds64(wa + 2 * TB * 8);
                }
                const int xkind = male ? 2 * 27 : 27; // chrX table of this member / of a child, by sex
                {   // own factor: transmission from the parents' current genotypes (a row of ones for founders)
                    const int row = cur.get(d & 127u) * 3 + cur.get((d >> 7) & 127u);
                    const int e = founder ? kOnes : (chrx ? xkind : 0) + row;
                    const uint32_t ta = tab_addr + (uint32_t)e * (kCopies * 8u);
                    w0 *= lds64(ta);
                    w1 *= lds64(ta + 9 * kCopies * 8);
                    w2 *= lds64(ta + 18 * kCopies * 8);
                }
                const int lb = (d >> 16) & 0xffu;
                const int n_links = (!chrx || male) ? (int)(d >> 24) : 0; // chrX: only males get the children factor
                const uint32_t step = male ? kCopies * 8u : 3u * kCopies * 8u;     // this member sits in the father slot when male
                auto child_entry = [&](int k) -> int {
                    const uint32_t l = pl.link[k];
                    const int kind = chrx ? (((l >> 14) & 1u) ? 2 * 27 : 27) : 0;
                    const int other = cur.get((l >> 7) & 127u);
                    return kind + cur.get(l & 127u) * 9 + (male ? other * 3 : other);
                };
                {   // first two children in line (a row of ones when there are fewer)
                    const int e0 = n_links > 0 ? child_entry(lb) : kOnes;
                    const int e1 = n_links > 1 ? child_entry(lb + 1) : kOnes;
                    const uint32_t a0 = tab_addr + (uint32_t)e0 * (kCopies * 8u), a1 = tab_addr + (uint32_t)e1 * (kCopies * 8u);
                    const double x0 = lds64(a0), x1 = lds64(a0 + step), x2 = lds64(a0 + 2 * step);
                    const double y0 = lds64(a1), y1 = lds64(a1 + step), y2 = lds64(a1 + 2 * step);
                    w0 = (w0 * x0) * y0;
                    w1 = (w1 * x1) * y1;
                    w2 = (w2 * x2) * y2;
                }
                for (int k = lb + 2; k < lb + n_links; k++) { // third and later children
                    const uint32_t a0 = tab_addr + (uint32_t)child_entry(k) * (kCopies * 8u);
                    w0 *= lds64(a0);
                    w1 *= lds64(a0 + step);
                    w2 *= lds64(a0 + 2 * step);
                }
                const double sum = (w0 + w1) + w2;
                if ((i & 3) == 0) philox4x32_10((uint32_t)sweep, (uint32_t)(i >> 2), gv_lo, gv_hi, k0, k1, r0, r1, r2, r3);
                const uint32_t u = r0; // words are consumed in order: rotate the block
                r0 = r1;
                r1 = r2;
                r2 = r3;
                const double rd = ((double)(u >> 1) + 0.5) * (1.0 / 2147483648.0); // 31 bits, like the reference's rand() / RAND_MAX
                // rd < w0/sum and rd > 1 - w2/sum decided without the division; a non-positive sum means all-zero
                // weights in the reference, which then draws genotype 1 (family.cpp:2142-2173)
                const double thr = rd * sum;
                int g = (thr < w0) ? 0 : ((thr > sum - w2) ? 2 : 1);
                if (!(sum > 0.0)) g = 1;
                cur.set(i, g);
                if (sampling) { // family.cpp:2175-2178, Rao-Blackwellised; all-zero weights add nothing
                    const double inv = sum > 0.0 ? fast_reciprocal(sum) : 0.0;
                    double *a = acc + (size_t)((i * 3) * TB);
                    atomicAdd(a, w0 * inv);
                    atomicAdd(a + TB, w1 * inv);
                    atomicAdd(a + 2 * TB, w2 * inv);
                }
            }
        }

        // ---- postProb = genoFry / numRep, not renormalised; a row summing to <= 0 fails (family.cpp:2082-2092) ----
        const double nrep = (double)rep;
        for (int i = 0; i < N; i++) {
            const double *a = acc + (size_t)((i * 3) * TB);
            const double p0 = __ldcg(a) / nrep, p1 = __ldcg(a + TB) / nrep, p2 = __ldcg(a + 2 * TB) / nrep;
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
}

template <int TB, bool WGLOBAL, int GW = 2>
cudaError_t launch_tb(const McmcParams &P, const BatchPtrs &B, int burn, int rep, uint64_t seed, int64_t v_offset, int sm_count,
                      cudaStream_t stream, int blocks_cap, bool fixup, unsigned long long *fixup_count) {
    const size_t smem = WGLOBAL ? (size_t)kEntries * kCopies * sizeof(double) : mcmc_smem_bytes(P, TB);
    cudaError_t rc = cudaFuncSetAttribute(mcmc_kernel<TB, WGLOBAL, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    int per_sm = 0;
    rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mcmc_kernel<TB, WGLOBAL, GW>, TB, smem);
    if (rc != cudaSuccess) return rc;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    if (WGLOBAL) { // resident blocks per SM (measured on ped40: 1 -> 1.60e5, 2 -> 1.63e5, 3 -> 1.79e5 variants/s)
        per_sm = std::min(per_sm, std::max(1, blocks_cap));
    }
    const int64_t n_tiles = (B.V + TB - 1) / TB;
    if (n_tiles > 0x7fffffff) return cudaErrorInvalidValue;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sm_count * per_sm);
    // thread-private chain state: [grid][N][3][TB] accumulators (+ as many own factors when WGLOBAL), stream-ordered
    double *scratch = nullptr;
    rc = cudaMallocAsync(&scratch, (size_t)grid * P.plan.n * 3 * TB * sizeof(double) * (WGLOBAL ? 2 : 1), stream);
    if (rc != cudaSuccess) return rc;
    mcmc_kernel<TB, WGLOBAL, GW><<<grid, TB, smem, stream>>>(P, B, burn, rep, seed, v_offset, scratch, (int)n_tiles, fixup ? 1 : 0,
                                                           fixup_count);
    rc = cudaGetLastError();
    const cudaError_t rc2 = cudaFreeAsync(scratch, stream);
    return rc != cudaSuccess ? rc : rc2;
}

} // namespace

size_t mcmc_smem_bytes(const McmcParams &P, int tb) { return (size_t)(kEntries * kCopies + P.plan.n * 3 * tb) * sizeof(double); }

// Block size that keeps the most chains resident per SM.
int mcmc_pick_block(const McmcParams &P, size_t smem_limit, size_t smem_per_sm) {
    int best = 0, best_threads = 0;
    for (int tb = 256; tb >= 32; tb -= 32) {
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

cudaError_t launch_mcmc(const McmcParams &P, const BatchPtrs &B, int tb, int burn, int rep, uint64_t seed, int64_t v_offset,
                        int sm_count, cudaStream_t stream, const McmcTuning &tune, bool fixup, unsigned long long *fixup_count) {
    if (B.V <= 0) return cudaSuccess;
    // Large pedigrees: the own factors of fewer than 512 chains fit in an SM's shared memory -> keep them in L2 instead.
    const bool wglobal = tune.wglobal < 0 ? tb < 256 : tune.wglobal != 0;
    if (P.plan.n > 64) return launch_tb<256, true, 4>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count); // wide genotype vector
    if (wglobal) return launch_tb<256, true>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    switch (tb) {
    case 256: return launch_tb<256, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    case 224: return launch_tb<224, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    case 192: return launch_tb<192, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    case 160: return launch_tb<160, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    case 128: return launch_tb<128, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    case 96: return launch_tb<96, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    case 64: return launch_tb<64, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    case 32: return launch_tb<32, false>(P, B, burn, rep, seed, v_offset, sm_count, stream, tune.blocks_cap, fixup, fixup_count);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace famseq
