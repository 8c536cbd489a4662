// es_nuclear_kernel.cu -- Elston-Stewart peeling for NUCLEAR FAMILIES (two founders and their C <= 5
// childless children: trios, quads, ...), the pedigree shape of almost every real FamSeq run and of the
// headline benchmark (BASELINE.json: 10 M-variant trio).  One variant per thread, everything in registers.
//
// It computes exactly what the message program of es_kernel.cu computes for such a pedigree -- the same
// products in the same association order as the reference (family.cpp:1501-1649, :1783-1845, :1292-1314;
// compiled with -fmad=false) -- so the doubles are bit-identical to the reference CPU build; the general
// interpreter remains the path for every other loop-free pedigree.  What the specialisation buys:
//   * no interpreter, no scratch in shared memory: ~3x fewer instructions per variant;
//   * the per-child vectors  K_c[a][b] = sum_l T_c[l][a][b] * lk_c[l]  are formed once and shared by the two
//     posterior messages and by the sibs' anterior messages (the recursion re-derives them each time);
//   * x/s for the three genotypes of a row shares one correctly rounded reciprocal (one Markstein
//     correction step per quotient; operands outside the safe exponent range take the plain IEEE divide);
//     tests/test_parity_gpu.py::test_nuclear_fast_path_is_bit_identical checks the result bit for bit;
//   * the block's input tile ([TB][S][3] FP64, contiguous in HBM) arrives through one TMA bulk copy
//     (cp.async.bulk + mbarrier) and the post / single / gt / status tiles leave through TMA bulk stores, so
//     global traffic is fully coalesced 16-byte-granular and costs no LSU wavefronts.
// HBM-bound: 73*S+2 algorithmic bytes per variant (221 B for a trio).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.hpp"

namespace famseq {

namespace {

// ---- TMA bulk copy helpers (sm_90+ PTX, SASS: UBLKCP) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool X> __device__ __forceinline__ double trans(const RunConstants &C, int sel, int g, int a, int b) {
    return X ? C.tab[sel][g * 9 + a * 3 + b] : C.tab[0][g * 9 + a * 3 + b];
}

struct Row3 {
    double v[3];
};

__device__ __forceinline__ Row3 load_lk(const double *in_row, int col) {
    Row3 r;
    if (col >= 0) {
        r.v[0] = in_row[col * 3];
        r.v[1] = in_row[col * 3 + 1];
        r.v[2] = in_row[col * 3 + 2];
    } else {
        r.v[0] = r.v[1] = r.v[2] = 1.0; // unsequenced member (file.cpp:565)
    }
    return r;
}

// marginal of one member: v = (m * l) * a, row sum == 0 fails the variant (family.cpp:1296-1314)
__device__ __forceinline__ bool finish(const Row3 &m, const Row3 &l, const Row3 &a, int col, double *post_row) {
    const double v0 = (m.v[0] * l.v[0]) * a.v[0], v1 = (m.v[1] * l.v[1]) * a.v[1], v2 = (m.v[2] * l.v[2]) * a.v[2];
    const double sum = (v0 + v1) + v2;
    if (col >= 0) div3(v0, v1, v2, sum, post_row[col * 3], post_row[col * 3 + 1], post_row[col * 3 + 2]);
    return sum == 0.0;
}

// The peeling of a nuclear family; returns true when the reference would return false.
template <int NC, bool X>
__device__ __forceinline__ bool peel(const NuclearParams &P, const VariantPriors &pr, const double *in_row, double *post_row) {
    const RunConstants &C = P.C;
    const Row3 lf = load_lk(in_row, P.col_father), lm = load_lk(in_row, P.col_mother);
    Row3 wf, wm, prior_f, prior_m, one;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        prior_f.v[g] = pr.m[g]; // the father is male: chrX male prior on X (family.cpp:1281-1290 / :1337-1358)
        prior_m.v[g] = pr.a[g];
        wf.v[g] = prior_f.v[g] * lf.v[g]; // ant * lk
        wm.v[g] = prior_m.v[g] * lm.v[g];
        one.v[g] = 1.0;
    }
    // K[c][a][b] = sum_l (T_c[l][a][b] * lk_c[l])
    Row3 lc[NC];
    double K[NC][3][3];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        lc[c] = load_lk(in_row, P.col_child[c]);
        const int sel = P.male_child[c] ? K_TAB_XM : K_TAB_XF;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                double sc = trans<X>(C, sel, 0, a, b) * lc[c].v[0];
                sc = sc + trans<X>(C, sel, 1, a, b) * lc[c].v[1];
                sc = sc + trans<X>(C, sel, 2, a, b) * lc[c].v[2];
                K[c][a][b] = sc;
            }
    }
    // product over all children, in ped order (the posterior messages of both parents use it)
    double kids[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            double p = K[0][a][b];
#pragma unroll
            for (int c = 1; c < NC; c++) p = p * K[c][a][b];
            kids[a][b] = p;
        }
    // posterior messages.  Autosome: the table index is (own genotype, spouse genotype) for both parents
    // (family.cpp:1836); chrX: (mother, father) (family.cpp:1900-1921).
    Row3 pos_f, pos_m;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        double a = wf.v[0] * kids[g][0]; // mother g, father b
        a = a + wf.v[1] * kids[g][1];
        a = a + wf.v[2] * kids[g][2];
        pos_m.v[g] = a;
        double f = wm.v[0] * (X ? kids[0][g] : kids[g][0]);
        f = f + wm.v[1] * (X ? kids[1][g] : kids[g][1]);
        f = f + wm.v[2] * (X ? kids[2][g] : kids[g][2]);
        pos_f.v[g] = f;
    }
    bool failed = false;
    // FIN in ped order does not matter for the values; every member's row sum is checked.
    failed |= finish(pos_f, lf, prior_f, P.col_father, post_row);
    failed |= finish(pos_m, lm, prior_m, P.col_mother, post_row);
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const int sel = P.male_child[c] ? K_TAB_XM : K_TAB_XF;
        Row3 ant;
#pragma unroll
        for (int g = 0; g < 3; g++) {
            double over_m = 0.0;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                double over_f = 0.0;
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    double term = wf.v[b] * trans<X>(C, sel, g, a, b);
                    if (NC > 1) { // full sibs in the mother's child order (family.cpp:1576-1587, :1616-1631)
                        double sibs = 1.0;
                        bool first = true;
#pragma unroll
                        for (int k = 0; k < NC; k++) {
                            if (k == c) continue;
                            sibs = first ? K[k][a][b] : sibs * K[k][a][b];
                            first = false;
                        }
                        term = term * sibs;
                    }
                    over_f = over_f + term;
                }
                over_m = over_m + wm.v[a] * over_f;
            }
            ant.v[g] = over_m;
        }
        failed |= finish(one, lc[c], ant, P.col_child[c], post_row);
    }
    return failed;
}

// One variant: individual-only posterior, LRC gate, peeling, genotype calls -- from the thread's row of the input tile
// into its rows of the output tiles (all in shared memory).
template <int NC>
__device__ __forceinline__ void variant_thread(const NuclearParams &P, const VariantPriors &pr, unsigned flag, int tid, const double *s_in,
                                               double *s_post, double *s_single, uint8_t *s_gt, uint8_t *s_status) {
    const RunConstants &C = P.C;
    const int S = C.s, S3 = 3 * S;
    if (P.io_probe) { // I/O ceiling probe: same tiles in and out, no arithmetic
        for (int k = 0; k < S3; k++) s_post[tid * S3 + k] = s_single[tid * S3 + k] = s_in[tid * S3 + k];
        for (int c = 0; c < S; c++) s_gt[tid * S + c] = 0;
        s_status[tid] = 0;
    } else {
        const double *in_row = s_in + tid * S3;
        double *post_row = s_post + tid * S3;
        double *single_row = s_single + tid * S3;
        // individual-only posterior (family.cpp:1405-1499) and LRC gate (family.cpp:1140-1162)
        bool failed = C.unseq_fail[flag & 3u] != 0;
        bool pedigree_needed = false;
        for (int c = 0; c < S; c++) {
            const double l0 = in_row[c * 3], l1 = in_row[c * 3 + 1], l2 = in_row[c * 3 + 2];
            const bool male = C.col_male[c] != 0;
            const double r0 = l0 * (male ? pr.m[0] : pr.a[0]);
            const double r1 = l1 * (male ? pr.m[1] : pr.a[1]);
            const double r2 = l2 * (male ? pr.m[2] : pr.a[2]);
            const double rs = (r0 + r1) + r2;
            if (rs <= 0.0) failed = true;
            div3(r0, r1, r2, rs, single_row[c * 3], single_row[c * 3 + 1], single_row[c * 3 + 2]);
            double big = 0.0;
            if (big < l0) big = l0;
            if (big < l1) big = l1;
            if (big < l2) big = l2;
            const double ls = (l0 + l1) + l2;
            if (lrc_wants_pedigree(C.lrc, l0, l1, l2, big, ls)) pedigree_needed = true;
        }
        if (!failed) {
            if (!pedigree_needed) {
                for (int k = 0; k < S3; k++) post_row[k] = single_row[k];
            } else if ((flag >> 1) & 1u) {
                failed = peel<NC, true>(P, pr, in_row, post_row);
            } else {
                failed = peel<NC, false>(P, pr, in_row, post_row);
            }
        }
        if (failed)
            for (int k = 0; k < S3; k++) post_row[k] = single_row[k] = 0.0;
        for (int c = 0; c < S; c++)
            s_gt[tid * S + c] = failed ? (uint8_t)255 : call_genotype(post_row[c * 3], post_row[c * 3 + 1], post_row[c * 3 + 2]);
        s_status[tid] = failed ? 1 : 0;
    }
}

template <int NC, int TB>
__global__ void __launch_bounds__(TB) es_nuclear_kernel(const __grid_constant__ NuclearParams P, const BatchPtrs B) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const RunConstants &C = P.C;
    const int S = C.s, S3 = 3 * S;
    double *s_in = reinterpret_cast<double *>(smem_raw); // [TB][S][3]
    double *s_post = s_in + TB * S3;
    double *s_single = s_post + TB * S3;
    uint8_t *s_gt = reinterpret_cast<uint8_t *>(s_single + TB * S3); // [TB][S]
    uint8_t *s_status = s_gt + ((TB * S + 15) & ~15);                 // [TB]
    __shared__ uint64_t bar;

    const int tid = threadIdx.x;
    const int64_t v0 = (int64_t)blockIdx.x * TB;
    const int nv = (int)min((int64_t)TB, B.V - v0);
    const bool full = nv == TB; // full tiles go through TMA; the ragged last tile uses plain loads/stores
    const unsigned tile_bytes = (unsigned)(TB * S3 * sizeof(double));

    if (full) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, tile_bytes);
            bulk_load(s_in, B.lk + v0 * S3, tile_bytes, &bar);
        }
    } else {
        for (int k = tid; k < nv * S3; k += TB) s_in[k] = B.lk[v0 * S3 + k];
    }
    unsigned flag = 0;
    if (tid < nv && B.flags) flag = B.flags[v0 + tid];
    const VariantPriors pr = select_priors(C, flag);
    if (full)
        mbar_wait(&bar, 0);
    else
        __syncthreads();

    if (tid < nv) variant_thread<NC>(P, pr, flag, tid, s_in, s_post, s_single, s_gt, s_status);

    if (full) {
        fence_async_smem(); // make this thread's shared-memory writes visible to the TMA engine
        __syncthreads();
        if (tid == 0) {
            bulk_store(B.post + v0 * S3, s_post, tile_bytes);
            bulk_store(B.single + v0 * S3, s_single, tile_bytes);
            bulk_store(B.gt + v0 * S, s_gt, (unsigned)(TB * S));
            bulk_store(B.status + v0, s_status, (unsigned)TB);
            bulk_commit_and_wait_read(); // shared memory must stay alive until the engine has read it
        }
    } else {
        __syncthreads();
        for (int k = tid; k < nv * S3; k += TB) {
            B.post[v0 * S3 + k] = s_post[k];
            B.single[v0 * S3 + k] = s_single[k];
        }
        for (int k = tid; k < nv * S; k += TB) B.gt[v0 * S + k] = s_gt[k];
        if (tid < nv) B.status[v0 + tid] = s_status[tid];
    }
}

template <int NC, int TB> cudaError_t launch_nc(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream) {
    const size_t S = (size_t)P.C.s;
    const size_t smem = 3 * TB * S * 3 * sizeof(double) + ((TB * S + 15) & ~(size_t)15) + TB;
    cudaError_t rc = cudaFuncSetAttribute(es_nuclear_kernel<NC, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    const unsigned grid = (unsigned)((B.V + TB - 1) / TB);
    es_nuclear_kernel<NC, TB><<<grid, TB, smem, stream>>>(P, B);
    return cudaGetLastError();
}

} // namespace

cudaError_t launch_es_nuclear(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream) {
    if (B.V <= 0) return cudaSuccess;
    // Variants per block (= per TMA tile).  Measured on the trio (profiles/es_tb_check.sh), fraction of the HBM peak:
    // 32 -> 0.946, 64 -> 0.933, 128 -> 0.904, 256 -> 0.72: many one-warp blocks per SM interleave their load / compute /
    // store phases best.  A persistent, double-buffered variant (one block per slot looping over tiles, next tile
    // requested before the current one is computed) was slower (0.854): its barriers serialise what the block
    // scheduler overlaps for free.
    static const int tb = [] {
        const char *env = std::getenv("FAMSEQ_ES_TB");
        return env ? std::atoi(env) : 32;
    }();
    switch (P.n_children * 1000 + tb) {
    case 1032: return launch_nc<1, 32>(P, B, stream);
    case 2032: return launch_nc<2, 32>(P, B, stream);
    case 3032: return launch_nc<3, 32>(P, B, stream);
    case 4032: return launch_nc<4, 32>(P, B, stream);
    case 5032: return launch_nc<5, 32>(P, B, stream);
    case 1064: return launch_nc<1, 64>(P, B, stream);
    case 2064: return launch_nc<2, 64>(P, B, stream);
    case 3064: return launch_nc<3, 64>(P, B, stream);
    case 4064: return launch_nc<4, 64>(P, B, stream);
    case 5064: return launch_nc<5, 64>(P, B, stream);
    case 1128: return launch_nc<1, 128>(P, B, stream);
    case 2128: return launch_nc<2, 128>(P, B, stream);
    case 3128: return launch_nc<3, 128>(P, B, stream);
    case 4128: return launch_nc<4, 128>(P, B, stream);
    case 5128: return launch_nc<5, 128>(P, B, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace famseq
