// es_kernel.cu -- Elston-Stewart peeling, one variant per thread, for sm_100a.
//
// Replaces family::calPostProbPeeling + calAntProb[X] + calPosProb[X] (src/family.cpp:1126-1403,
// :1501-1930).  The reference walks a memoised recursion per variant with heap-allocated matrices;
// here the recursion is compiled once on the host into a message program (host/es_program.cpp) that
// every thread interprets in lock step for its own variant, so there is no divergence and no
// per-variant allocation.
//
// Data movement (the kernel is HBM-bound for small pedigrees):
//   * a block owns TB consecutive variants = one contiguous [TB][S][3] FP64 tile of the input; it is
//     copied to shared memory with 16-byte streaming loads (fully coalesced) and read back per thread;
//   * message 3-vectors live in shared memory as [slot][g][thread] (conflict-free);
//   * post / single / gt tiles are staged in shared memory and written back as contiguous 16-byte
//     streaming stores.
// Arithmetic: this file is compiled with -fmad=false and every product is formed in the reference's
// association order, so the FP64 results are bit-identical to the reference CPU build.
#include "common.cuh"
#include "kernels.hpp"

namespace famseq {

namespace {

template <bool X> __device__ __forceinline__ double trans(const RunConstants &C, int sel, int g, int a, int b) {
    return X ? C.tab[sel][g * 9 + a * 3 + b] : C.tab[0][g * 9 + a * 3 + b];
}

template <int TB> struct EsThread {
    const double *in_row; // this variant's [S][3] likelihoods in shared memory
    double *slot;         // base of the [slot][g][TB] scratch, already offset by the thread index
    VariantPriors pr;

    __device__ __forceinline__ void load(uint32_t r, double v[3]) const {
        const uint32_t kind = es_ref_kind(r), idx = es_ref_index(r);
        if (kind == ES_REF_SLOT) {
#pragma unroll
            for (int g = 0; g < 3; g++) v[g] = slot[(idx * 3 + g) * TB];
        } else if (kind == ES_REF_LK) {
#pragma unroll
            for (int g = 0; g < 3; g++) v[g] = in_row[idx * 3 + g];
        } else if (kind == ES_REF_PRIOR) {
#pragma unroll
            for (int g = 0; g < 3; g++) v[g] = idx ? pr.m[g] : pr.a[g];
        } else {
            v[0] = v[1] = v[2] = 1.0;
        }
    }
    __device__ __forceinline__ void store(uint32_t dst, const double v[3]) const {
#pragma unroll
        for (int g = 0; g < 3; g++) slot[(dst * 3 + g) * TB] = v[g];
    }
};

// Interprets the message program for one variant.  Returns true when a member's row sum was exactly
// zero (the reference returns false there, family.cpp:1305-1310).
template <bool X, int TB>
__device__ bool es_interpret(const EsParams &P, const EsThread<TB> &t, double *post_row) {
    const RunConstants &C = P.C;
    bool failed = false;
    int pc = 0;
    for (;;) {
        const uint32_t w0 = P.prog.words[pc];
        const uint32_t op = w0 & 0xffu;
        if (op == ES_OP_END) break;
        if (op == ES_OP_MUL) {
            const uint32_t w1 = P.prog.words[pc + 1];
            double a[3], b[3];
            t.load(w1 & 0xffffu, a);
            t.load(w1 >> 16, b);
#pragma unroll
            for (int g = 0; g < 3; g++) a[g] = a[g] * b[g];
            t.store((w0 >> 8) & 0xffffu, a);
            pc += 2;
        } else if (op == ES_OP_ANT) {
            const uint32_t w1 = P.prog.words[pc + 1];
            const int nsib = w0 >> 25;
            const int sel_c = ((w0 >> 24) & 1u) ? K_TAB_XM : K_TAB_XF;
            double wm[3], wf[3], sibs[3][3];
            t.load(w1 & 0xffffu, wm);
            t.load(w1 >> 16, wf);
            for (int k = 0; k < nsib; k++) {
                const uint32_t wk = P.prog.words[pc + 2 + k];
                const int sel_k = ((wk >> 16) & 1u) ? K_TAB_XM : K_TAB_XF;
                double d[3];
                t.load(wk & 0xffffu, d);
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        double sc = d[0] * trans<X>(C, sel_k, 0, a, b);
                        sc = sc + d[1] * trans<X>(C, sel_k, 1, a, b);
                        sc = sc + d[2] * trans<X>(C, sel_k, 2, a, b);
                        sibs[a][b] = (k == 0) ? sc : sibs[a][b] * sc;
                    }
            }
            double out[3];
#pragma unroll
            for (int g = 0; g < 3; g++) {
                double over_m = 0.0;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    double over_f = 0.0;
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        double term = wf[b] * trans<X>(C, sel_c, g, a, b);
                        if (nsib) term = term * sibs[a][b];
                        over_f = over_f + term;
                    }
                    over_m = over_m + wm[a] * over_f;
                }
                out[g] = over_m;
            }
            t.store((w0 >> 8) & 0xffffu, out);
            pc += 2 + nsib;
        } else if (op == ES_OP_POS) {
            const uint32_t w1 = P.prog.words[pc + 1];
            const int nkid = w0 >> 25;
            // autosome: table index is (g_i, g_j) whatever the sexes (family.cpp:1836);
            // chrX: (mother, father), so a male i goes second (family.cpp:1900-1921)
            const bool i_second = X && ((w0 >> 24) & 1u);
            double wj[3], kids[3][3];
            t.load(w1 & 0xffffu, wj);
            for (int k = 0; k < nkid; k++) {
                const uint32_t wk = P.prog.words[pc + 2 + 2 * k];
                const int sel_k = (P.prog.words[pc + 3 + 2 * k] & 1u) ? K_TAB_XM : K_TAB_XF;
                double lkc[3], mc[3];
                t.load(wk & 0xffffu, lkc);
                t.load(wk >> 16, mc);
#pragma unroll
                for (int g = 0; g < 3; g++)
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        double sc = 0.0;
#pragma unroll
                        for (int l = 0; l < 3; l++) {
                            const double tr = i_second ? trans<X>(C, sel_k, l, b, g) : trans<X>(C, sel_k, l, g, b);
                            const double term = (tr * lkc[l]) * mc[l];
                            sc = (l == 0) ? term : sc + term;
                        }
                        kids[g][b] = (k == 0) ? sc : kids[g][b] * sc;
                    }
            }
            double out[3];
#pragma unroll
            for (int g = 0; g < 3; g++) {
                double over_j = wj[0] * kids[g][0];
                over_j = over_j + wj[1] * kids[g][1];
                over_j = over_j + wj[2] * kids[g][2];
                out[g] = over_j;
            }
            t.store((w0 >> 8) & 0xffffu, out);
            pc += 2 + 2 * nkid;
        } else { // ES_OP_FIN
            const uint32_t w1 = P.prog.words[pc + 1], w2 = P.prog.words[pc + 2];
            double m[3], l[3], a[3];
            t.load(w1 & 0xffffu, m);
            t.load(w1 >> 16, l);
            t.load(w2 & 0xffffu, a);
#pragma unroll
            for (int g = 0; g < 3; g++) m[g] = (m[g] * l[g]) * a[g];
            const double sum = (m[0] + m[1]) + m[2];
            if (sum == 0.0) failed = true;
            if ((w0 >> 8) & 1u) {
                const int col = w0 >> 9;
                div3(m[0], m[1], m[2], sum, post_row[col * 3], post_row[col * 3 + 1], post_row[col * 3 + 2]);
            }
            pc += 3;
        }
    }
    return failed;
}

template <int TB> __global__ void __launch_bounds__(TB) es_kernel(const __grid_constant__ EsParams P, const BatchPtrs B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RunConstants &C = P.C;
    const int S = C.s, S3 = 3 * S;
    double *s_in = reinterpret_cast<double *>(smem_raw); // [TB][S][3], same layout as global
    double *s_post = s_in + TB * S3;
    double *s_single = s_post + TB * S3;
    double *s_slot = s_single + TB * S3; // [n_slots][3][TB]
    uint8_t *s_gt = reinterpret_cast<uint8_t *>(s_slot + (size_t)P.prog.n_slots * 3 * TB); // [TB][S]
    uint8_t *s_status = s_gt + TB * S;                                                     // [TB]

    const int tid = threadIdx.x;
    const int64_t v0 = (int64_t)blockIdx.x * TB;
    const int nv = (int)min((int64_t)TB, B.V - v0);
    const int ndbl = nv * S3;

    { // tile load: contiguous, 16-byte, streaming
        const double *gin = B.lk + v0 * S3;
        const double2 *g2 = reinterpret_cast<const double2 *>(gin);
        double2 *s2 = reinterpret_cast<double2 *>(s_in);
        for (int k = tid; k < (ndbl >> 1); k += TB) s2[k] = __ldcs(g2 + k);
        if (tid == 0 && (ndbl & 1)) s_in[ndbl - 1] = __ldcs(gin + ndbl - 1);
    }
    __syncthreads();

    if (tid < nv) {
        const unsigned flag = B.flags ? B.flags[v0 + tid] : 0u;
        EsThread<TB> t;
        t.in_row = s_in + tid * S3;
        t.slot = s_slot + tid;
        t.pr = select_priors(C, flag);
        double *post_row = s_post + tid * S3;
        double *single_row = s_single + tid * S3;

        // individual-only posterior (family.cpp:1405-1499) and LRC gate (family.cpp:1140-1162)
        bool failed = C.unseq_fail[flag & 3u] != 0;
        bool pedigree_needed = false;
        for (int c = 0; c < S; c++) {
            const double l0 = t.in_row[c * 3], l1 = t.in_row[c * 3 + 1], l2 = t.in_row[c * 3 + 2];
            const bool male = C.col_male[c] != 0;
            const double r0 = l0 * (male ? t.pr.m[0] : t.pr.a[0]);
            const double r1 = l1 * (male ? t.pr.m[1] : t.pr.a[1]);
            const double r2 = l2 * (male ? t.pr.m[2] : t.pr.a[2]);
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
            if (!pedigree_needed) { // family.cpp:1164-1253: FPP := individual-only posterior
                for (int k = 0; k < S3; k++) post_row[k] = single_row[k];
            } else if ((flag >> 1) & 1u) {
                failed = es_interpret<true, TB>(P, t, post_row);
            } else {
                failed = es_interpret<false, TB>(P, t, post_row);
            }
        }
        if (failed) {
            for (int k = 0; k < S3; k++) {
                post_row[k] = 0.0;
                single_row[k] = 0.0;
            }
        }
        for (int c = 0; c < S; c++)
            s_gt[tid * S + c] = failed ? (uint8_t)255 : call_genotype(post_row[c * 3], post_row[c * 3 + 1], post_row[c * 3 + 2]);
        s_status[tid] = failed ? 1 : 0;
    }
    __syncthreads();

    { // tile store
        double2 *gp = reinterpret_cast<double2 *>(B.post + v0 * S3);
        double2 *gs = reinterpret_cast<double2 *>(B.single + v0 * S3);
        const double2 *sp = reinterpret_cast<const double2 *>(s_post);
        const double2 *ss = reinterpret_cast<const double2 *>(s_single);
        for (int k = tid; k < (ndbl >> 1); k += TB) {
            __stcs(gp + k, sp[k]);
            __stcs(gs + k, ss[k]);
        }
        if (tid == 0 && (ndbl & 1)) {
            B.post[v0 * S3 + ndbl - 1] = s_post[ndbl - 1];
            B.single[v0 * S3 + ndbl - 1] = s_single[ndbl - 1];
        }
        const int ngt = nv * S;
        uint8_t *ggt = B.gt + v0 * S;
        if ((ngt & 3) == 0 && ((reinterpret_cast<uintptr_t>(ggt) & 3u) == 0)) {
            const uint32_t *s4 = reinterpret_cast<const uint32_t *>(s_gt);
            uint32_t *g4 = reinterpret_cast<uint32_t *>(ggt);
            for (int k = tid; k < (ngt >> 2); k += TB) g4[k] = s4[k];
        } else {
            for (int k = tid; k < ngt; k += TB) ggt[k] = s_gt[k];
        }
        if (tid < nv) B.status[v0 + tid] = s_status[tid];
    }
}

} // namespace

size_t es_smem_bytes(const EsParams &P, int tb) {
    const size_t S = (size_t)P.C.s;
    size_t bytes = 3 * (size_t)tb * S * 3 * sizeof(double);        // in, post, single tiles
    bytes += (size_t)P.prog.n_slots * 3 * tb * sizeof(double);     // message scratch
    bytes += (size_t)tb * S + tb;                                  // gt, status
    return (bytes + 15) & ~(size_t)15;
}

// Picks the largest block the scratch fits in (more threads per SM hide FP64 and LDS latency).
int es_pick_block(const EsParams &P, size_t smem_limit) {
    const int candidates[] = {128, 64, 32};
    for (int tb : candidates)
        if (es_smem_bytes(P, tb) <= smem_limit) return tb;
    return 0;
}

cudaError_t launch_es(const EsParams &P, const BatchPtrs &B, int tb, cudaStream_t stream) {
    if (B.V <= 0) return cudaSuccess;
    const size_t smem = es_smem_bytes(P, tb);
    const unsigned grid = (unsigned)((B.V + tb - 1) / tb);
    cudaError_t rc;
    switch (tb) {
    case 128:
        rc = cudaFuncSetAttribute(es_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        es_kernel<128><<<grid, 128, smem, stream>>>(P, B);
        break;
    case 64:
        rc = cudaFuncSetAttribute(es_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        es_kernel<64><<<grid, 64, smem, stream>>>(P, B);
        break;
    case 32:
        rc = cudaFuncSetAttribute(es_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        es_kernel<32><<<grid, 32, smem, stream>>>(P, B);
        break;
    default:
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace famseq
