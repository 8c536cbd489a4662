// es_jit.cu -- Elston-Stewart peeling of ONE loop-free pedigree as straight-line code, generated from the message
// program (host/es_program.cpp) and compiled at run time for sm_100a (NVRTC, through gibbs_jit_compile()).
//
// Replaces family::calPostProbPeeling + calAntProb[X] + calPosProb[X] (src/family.cpp:1126-1403, :1501-1930) for
// pedigrees that are not nuclear families, where the interpreter of es_kernel.cu is the general path.  The interpreter
// keeps every operand of the program in a per-variant vector file in shared memory (720 B per variant for the
// 14-member pedigree), which limits it to 9 warps per SM and makes every operand three shared-memory loads.  Here the
// program is known when the code is written:
//   * every message is a set of named doubles (SSA), so operands live in registers and the compiler schedules the
//     whole pedigree as two basic blocks (autosomal and chrX rules);
//   * transmission-table entries are literal constants (constant-bank operands of the FP64 instructions);
//   * one warp per block and one variant per thread: the block's likelihood tile arrives through one TMA bulk copy,
//     each thread pulls its row into registers, writes its `single` row into a second tile and its `post` row over
//     its own input row, and both tiles leave through TMA bulk stores (as in es_nuclear_kernel.cu).
// Same operations in the same order as the interpreter (no FMA contraction, the same shared-reciprocal division), so
// the results are the same doubles; the only liberty is that sums start from their first term instead of from 0.0,
// which can only change the sign of a zero.  tests/test_parity_gpu.py compares both kernels bit for bit.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

#include "../../../include/famseq_b200.h"
#include "es_jit.hpp"
#include "gibbs_jit.hpp"

namespace famseq {

namespace {

// Variants (threads) per block.  One warp per block moves its tile with the least synchronisation; several warps per block start
// together and walk the (long, straight-line) program roughly in step, which keeps their instruction fetches in the same
// cache lines.  FAMSEQ_ES_JIT_TB (32 .. 256, multiple of 32) picks it; see es_jit_tb().
int es_jit_tb() {
    static const int tb = [] {
        const char *env = std::getenv("FAMSEQ_ES_JIT_TB");
        const int v = env ? std::atoi(env) : 32;
        return (v >= 32 && v <= 256 && v % 32 == 0) ? v : 32;
    }();
    return tb;
}
#define kTB es_jit_tb()

std::string lit(double x) {
    unsigned long long u;
    std::memcpy(&u, &x, 8);
    char buf[64];
    std::snprintf(buf, sizeof buf, "__longlong_as_double(0x%016llxll)", u);
    return buf;
}

const char *kPrelude = R"CUDA(
typedef unsigned int u32;
typedef unsigned long long u64;
typedef long long i64;
typedef unsigned char u8;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
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
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, u64 *bar) {
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

// x[0..2] / s, correctly rounded: shared reciprocal + one Markstein correction, exponent-guarded (cuda/common.cuh: div3)
__device__ __forceinline__ bool safe_dividend(double x) {
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu, lo = (unsigned)__double2loint(x);
    return (hi - ((1023u - 900u) << 20)) <= (1800u << 20) || (hi | lo) == 0u;
}
__device__ __forceinline__ void div3(double x0, double x1, double x2, double s, double &q0, double &q1, double &q2) {
    const bool fast = ((unsigned)__double2hiint(s) - ((1023u - 900u) << 20)) <= (1800u << 20) & safe_dividend(x0) & safe_dividend(x1) &
                      safe_dividend(x2);
    if (fast) {
        const double r = __drcp_rn(s);
        const double a = __dmul_rn(x0, r), b = __dmul_rn(x1, r), c = __dmul_rn(x2, r);
        q0 = __fma_rn(__fma_rn(-s, a, x0), r, a);
        q1 = __fma_rn(__fma_rn(-s, b, x1), r, b);
        q2 = __fma_rn(__fma_rn(-s, c, x2), r, c);
    } else {
        q0 = __ddiv_rn(x0, s);
        q1 = __ddiv_rn(x1, s);
        q2 = __ddiv_rn(x2, s);
    }
}
__device__ __forceinline__ bool lrc_wants_pedigree(double lrc, double l0, double l1, double l2, double big, double ls) {
    const bool no_negative = (__double2hiint(l0) | __double2hiint(l1) | __double2hiint(l2)) >= 0;
    if (lrc == 1.0 && no_negative) return big < ls;
    return __ddiv_rn(big, ls) < lrc;
}
__device__ __forceinline__ u8 call_genotype(double p0, double p1, double p2) {
    double big = -1.0;
    int arg = -1;
    if (big < p0) { big = p0; arg = 0; }
    if (big < p1) { big = p1; arg = 1; }
    if (big < p2) { big = p2; arg = 2; }
    return (u8)arg;
}
)CUDA";

// Emits the message program as SSA code.  `x` selects the chrX rules.
class Emitter {
  public:
    Emitter(const EsParams &P, bool x, std::ostringstream &o) : P_(P), C_(P.C), x_(x), o_(o), S_(P.C.s), cur_(P.prog.n_slots, -1) {}

    void run() {
        const uint32_t *w = P_.prog.words;
        int pc = 0;
        for (;;) {
            const uint32_t w0 = w[pc], op = w0 & 0xffu;
            if (op == ES_OP_END) break;
            if (op == ES_OP_MUL) {
                const uint32_t w1 = w[pc + 1];
                const std::string a = name(w1 & 0xffffu), b = name(w1 >> 16), d = define((w0 >> 8) & 0xffffu);
                for (int g = 0; g < 3; g++) o_ << "    const double " << d << g << " = __dmul_rn(" << comp(a, g) << ", " << comp(b, g) << ");\n";
                pc += 2;
            } else if (op == ES_OP_ANT) {
                const uint32_t w1 = w[pc + 1];
                const int nsib = (int)(w0 >> 25), sel_c = ((w0 >> 24) & 1u) ? K_TAB_XM : K_TAB_XF;
                const std::string wm = name(w1 & 0xffffu), wf = name(w1 >> 16);
                const std::string t = tmp();
                for (int k = 0; k < nsib; k++) {
                    const uint32_t wk = w[pc + 2 + k];
                    const int sel_k = ((wk >> 16) & 1u) ? K_TAB_XM : K_TAB_XF;
                    const std::string d = name(wk & 0xffffu);
                    for (int a = 0; a < 3; a++)
                        for (int b = 0; b < 3; b++) {
                            // sc = d0*T0 + d1*T1 + d2*T2, left to right; sibs = sibs * sc
                            std::string sc = "__dadd_rn(__dadd_rn(__dmul_rn(" + comp(d, 0) + ", " + T(sel_k, 0, a, b) + "), __dmul_rn(" + comp(d, 1) + ", " +
                                             T(sel_k, 1, a, b) + ")), __dmul_rn(" + comp(d, 2) + ", " + T(sel_k, 2, a, b) + "))";
                            const std::string v = t + "s" + std::to_string(k) + "_" + std::to_string(a) + std::to_string(b);
                            if (k == 0)
                                o_ << "    const double " << v << " = " << sc << ";\n";
                            else
                                o_ << "    const double " << v << " = __dmul_rn(" << t << "s" << k - 1 << "_" << a << b << ", " << sc << ");\n";
                        }
                }
                const std::string d = define((w0 >> 8) & 0xffffu);
                for (int g = 0; g < 3; g++) {
                    std::string over_m;
                    for (int a = 0; a < 3; a++) {
                        std::string over_f;
                        for (int b = 0; b < 3; b++) {
                            std::string term = "__dmul_rn(" + comp(wf, b) + ", " + T(sel_c, g, a, b) + ")";
                            if (nsib) term = "__dmul_rn(" + term + ", " + t + "s" + std::to_string(nsib - 1) + "_" + std::to_string(a) + std::to_string(b) + ")";
                            over_f = b == 0 ? term : "__dadd_rn(" + over_f + ", " + term + ")";
                        }
                        const std::string term = "__dmul_rn(" + comp(wm, a) + ", " + over_f + ")";
                        over_m = a == 0 ? term : "__dadd_rn(" + over_m + ", " + term + ")";
                    }
                    o_ << "    const double " << d << g << " = " << over_m << ";\n";
                }
                pc += 2 + nsib;
            } else if (op == ES_OP_POS) {
                const uint32_t w1 = w[pc + 1];
                const int nkid = (int)(w0 >> 25);
                const bool i_second = x_ && ((w0 >> 24) & 1u); // chrX: (mother, father) order, a male i goes second
                const std::string wj = name(w1 & 0xffffu);
                const std::string t = tmp();
                for (int k = 0; k < nkid; k++) {
                    const uint32_t wk = w[pc + 2 + 2 * k];
                    const int sel_k = (w[pc + 3 + 2 * k] & 1u) ? K_TAB_XM : K_TAB_XF;
                    const std::string lkc = name(wk & 0xffffu), mc = name(wk >> 16);
                    for (int g = 0; g < 3; g++)
                        for (int b = 0; b < 3; b++) {
                            std::string sc;
                            for (int l = 0; l < 3; l++) {
                                const std::string tr = i_second ? T(sel_k, l, b, g) : T(sel_k, l, g, b);
                                const std::string term = "__dmul_rn(__dmul_rn(" + tr + ", " + comp(lkc, l) + "), " + comp(mc, l) + ")";
                                sc = l == 0 ? term : "__dadd_rn(" + sc + ", " + term + ")";
                            }
                            const std::string v = t + "k" + std::to_string(k) + "_" + std::to_string(g) + std::to_string(b);
                            if (k == 0)
                                o_ << "    const double " << v << " = " << sc << ";\n";
                            else
                                o_ << "    const double " << v << " = __dmul_rn(" << t << "k" << k - 1 << "_" << g << b << ", " << sc << ");\n";
                        }
                }
                const std::string d = define((w0 >> 8) & 0xffffu);
                for (int g = 0; g < 3; g++) {
                    std::string over_j;
                    for (int b = 0; b < 3; b++) {
                        const std::string kid = nkid ? t + "k" + std::to_string(nkid - 1) + "_" + std::to_string(g) + std::to_string(b) : std::string("0.0");
                        const std::string term = "__dmul_rn(" + comp(wj, b) + ", " + kid + ")";
                        over_j = b == 0 ? term : "__dadd_rn(" + over_j + ", " + term + ")";
                    }
                    o_ << "    const double " << d << g << " = " << over_j << ";\n";
                }
                pc += 2 + 2 * nkid;
            } else { // ES_OP_FIN
                const uint32_t w1 = w[pc + 1], w2 = w[pc + 2];
                const std::string m = name(w1 & 0xffffu), l = name(w1 >> 16), a = name(w2 & 0xffffu);
                const std::string t = tmp();
                for (int g = 0; g < 3; g++)
                    o_ << "    const double " << t << "m" << g << " = __dmul_rn(__dmul_rn(" << comp(m, g) << ", " << comp(l, g) << "), " << comp(a, g) << ");\n";
                o_ << "    const double " << t << "sum = __dadd_rn(__dadd_rn(" << t << "m0, " << t << "m1), " << t << "m2);\n";
                o_ << "    if (" << t << "sum == 0.0) failed = true;\n";
                if ((w0 >> 8) & 1u) {
                    const int col = (int)(w0 >> 9);
                    o_ << "    { double q0, q1, q2; div3(" << t << "m0, " << t << "m1, " << t << "m2, " << t << "sum, q0, q1, q2);\n"
                       << "      row[" << col * 3 << "] = q0; row[" << col * 3 + 1 << "] = q1; row[" << col * 3 + 2 << "] = q2; gt_row[" << col
                       << "] = call_genotype(q0, q1, q2); }\n";
                }
                pc += 3;
            }
        }
    }

  private:
    const EsParams &P_;
    const RunConstants &C_;
    bool x_;
    std::ostringstream &o_;
    int S_;
    std::vector<int> cur_; // SSA version of every scratch slot
    int n_def_ = 0, n_tmp_ = 0;

    std::string tmp() { return (x_ ? "xt" : "at") + std::to_string(n_tmp_++) + "_"; }
    // name prefix of operand u; components are prefix + "0/1/2"
    std::string name(uint32_t u) const {
        const int s = (int)u;
        if (s < S_) return "L" + std::to_string(s) + "_";
        if (s < S_ + P_.prog.n_slots) return (x_ ? "xv" : "av") + std::to_string(cur_[s - S_]) + "_";
        if (s == S_ + P_.prog.n_slots) return "pa";
        if (s == S_ + P_.prog.n_slots + 1) return "pm";
        return "ONE";
    }
    static std::string comp(const std::string &prefix, int g) { return prefix == "ONE" ? std::string("1.0") : prefix + std::to_string(g); }
    std::string define(uint32_t u) {
        cur_[(int)u - S_] = n_def_++;
        return name(u);
    }
    std::string T(int sel, int g, int a, int b) const { return lit(x_ ? C_.tab[sel][g * 9 + a * 3 + b] : C_.tab[0][g * 9 + a * 3 + b]); }
};

} // namespace

std::string es_jit_source(const EsParams &P) {
    const RunConstants &C = P.C;
    const int S = C.s, S3 = 3 * S;
    std::ostringstream o;
    o << "// generated by famseq_b200 (es_jit.cu): Elston-Stewart peeling of one pedigree, " << C.n << " members, " << S << " input columns, "
      << P.prog.n_ops << " messages\n";
    o << "#define TB " << kTB << "\n#define NCOL " << S << "\n#define S3 " << S3 << "\n";
    o << kPrelude;
    o << "__constant__ u8 COL_MALE[NCOL + 1] = {";
    for (int c = 0; c < S; c++) o << (int)C.col_male[c] << ", ";
    o << "0};\n";
    const unsigned unseq = (C.unseq_fail[0] ? 1u : 0u) | (C.unseq_fail[1] ? 2u : 0u) | (C.unseq_fail[2] ? 4u : 0u) | (C.unseq_fail[3] ? 8u : 0u);
    for (int x = 0; x < 2; x++) {
        o << "\n// the message program with the " << (x ? "chrX" : "autosomal") << " rules; returns true when a member's row sum was exactly zero\n"
          << "__device__ __forceinline__ bool peel_" << (x ? "x" : "a") << "(double *row, u8 *gt_row, double pa0, double pa1, double pa2, double pm0, double pm1, double pm2";
        for (int c = 0; c < S; c++)
            for (int g = 0; g < 3; g++) o << ", double L" << c << "_" << g;
        o << ") {\n    bool failed = false;\n";
        Emitter(P, x != 0, o).run();
        o << "    return failed;\n}\n";
    }
    int min_blocks = 0; // tuning knob: resident blocks per SM the register allocation is forced to allow
    if (const char *env = std::getenv("FAMSEQ_ES_JIT_BLOCKS")) min_blocks = std::max(0, std::min(32, std::atoi(env)));
    o << "\nextern \"C\" __global__ void __launch_bounds__(TB" << (min_blocks ? ", " + std::to_string(min_blocks) : std::string()) << ")\n"
      << "famseq_es(const double *__restrict__ lk, const u8 *__restrict__ flags, double *__restrict__ post, double *__restrict__ single,\n"
      << "          u8 *__restrict__ gt, u8 *__restrict__ status, i64 V) {\n"
      << "    extern __shared__ __align__(128) unsigned char smem_raw[];\n"
      << "    double *s_tile = (double *)smem_raw;   // [TB][S3] likelihood rows in, posterior rows out\n"
      << "    double *s_single = s_tile + TB * S3;   // [TB][S3]\n"
      << "    u8 *s_gt = (u8 *)(s_single + TB * S3); // [TB][NCOL]\n"
      << "    u8 *s_status = s_gt + ((TB * NCOL + 15) & ~15);\n"
      << "    __shared__ u64 bar;\n"
      << "    const int tid = threadIdx.x;\n"
      << "    const i64 v0 = (i64)blockIdx.x * TB;\n"
      << "    const int nv = (int)((V - v0) < (i64)TB ? (V - v0) : (i64)TB);\n"
      << "    const bool full = nv == TB; // full tiles go through TMA; the ragged last tile uses plain loads / stores\n"
      << "    const unsigned tile_bytes = (unsigned)(TB * S3 * sizeof(double));\n"
      << "    if (full) {\n"
      << "        if (tid == 0) mbar_init(&bar, 1);\n"
      << "        __syncthreads();\n"
      << "        if (tid == 0) { mbar_expect_tx(&bar, tile_bytes); bulk_load(s_tile, lk + v0 * S3, tile_bytes, &bar); }\n"
      << "    } else {\n"
      << "        for (int k = tid; k < nv * S3; k += TB) s_tile[k] = lk[v0 * S3 + k];\n"
      << "    }\n"
      << "    u32 flag = 0;\n"
      << "    if (tid < nv && flags) flag = flags[v0 + tid];\n"
      << "    const bool known = flag & 1u, chrx = (flag >> 1) & 1u;\n";
    for (int g = 0; g < 3; g++)
        o << "    const double pa" << g << " = known ? " << lit(C.prior[1][g]) << " : " << lit(C.prior[0][g]) << ";\n"
          << "    const double pm" << g << " = chrx ? (known ? " << lit(C.prior[3][g]) << " : " << lit(C.prior[2][g]) << ") : pa" << g << ";\n";
    o << "    if (full) mbar_wait(&bar, 0); else __syncthreads();\n"
      << "    if (tid < nv) {\n"
      << "        double *row = s_tile + tid * S3, *single_row = s_single + tid * S3;\n"
      << "        u8 *gt_row = s_gt + tid * NCOL;\n";
    for (int c = 0; c < S; c++)
        o << "        const double L" << c << "_0 = row[" << c * 3 << "], L" << c << "_1 = row[" << c * 3 + 1 << "], L" << c << "_2 = row[" << c * 3 + 2 << "];\n";
    o << "        // individual-only posterior (family.cpp:1405-1499) and LRC gate (family.cpp:1140-1162); the posterior row starts\n"
      << "        // as a copy of it and stays that way when the gate keeps the pedigree out (family.cpp:1164-1253)\n"
      << "        bool failed = (" << unseq << "u >> (flag & 3u)) & 1u;\n"
      << "        bool pedigree_needed = false;\n"
      << "        const double lrc = " << lit(C.lrc) << ";\n";
    for (int c = 0; c < S; c++) {
        const char *pr = C.col_male[c] ? "pm" : "pa";
        o << "        {\n"
          << "            const double r0 = __dmul_rn(L" << c << "_0, " << pr << "0), r1 = __dmul_rn(L" << c << "_1, " << pr << "1), r2 = __dmul_rn(L" << c << "_2, " << pr << "2);\n"
          << "            const double rs = __dadd_rn(__dadd_rn(r0, r1), r2);\n"
          << "            if (rs <= 0.0) failed = true;\n"
          << "            double q0, q1, q2; div3(r0, r1, r2, rs, q0, q1, q2);\n"
          << "            single_row[" << c * 3 << "] = q0; single_row[" << c * 3 + 1 << "] = q1; single_row[" << c * 3 + 2 << "] = q2;\n"
          << "            row[" << c * 3 << "] = q0; row[" << c * 3 + 1 << "] = q1; row[" << c * 3 + 2 << "] = q2;\n"
          << "            gt_row[" << c << "] = call_genotype(q0, q1, q2);\n"
          << "            double big = 0.0;\n"
          << "            if (big < L" << c << "_0) big = L" << c << "_0;\n"
          << "            if (big < L" << c << "_1) big = L" << c << "_1;\n"
          << "            if (big < L" << c << "_2) big = L" << c << "_2;\n"
          << "            if (lrc_wants_pedigree(lrc, L" << c << "_0, L" << c << "_1, L" << c << "_2, big, __dadd_rn(__dadd_rn(L" << c << "_0, L" << c << "_1), L" << c
          << "_2))) pedigree_needed = true;\n"
          << "        }\n";
    }
    auto call = [&](const char *fn) {
        o << fn << "(row, gt_row, pa0, pa1, pa2, pm0, pm1, pm2";
        for (int c = 0; c < S; c++)
            for (int g = 0; g < 3; g++) o << ", L" << c << "_" << g;
        o << ")";
    };
    o << "        if (!failed && pedigree_needed) {\n            if (chrx) failed = ";
    call("peel_x");
    o << ";\n            else failed = ";
    call("peel_a");
    o << ";\n        }\n"
      << "        if (failed) {\n"
      << "            for (int k = 0; k < S3; k++) row[k] = single_row[k] = 0.0;\n"
      << "            for (int c = 0; c < NCOL; c++) gt_row[c] = 255;\n"
      << "        }\n"
      << "        s_status[tid] = failed ? 1 : 0;\n"
      << "    }\n"
      << "    if (full) {\n"
      << "        fence_async_smem(); // make this thread's shared-memory writes visible to the TMA engine\n"
      << "        __syncthreads();\n"
      << "        if (tid == 0) {\n"
      << "            bulk_store(post + v0 * S3, s_tile, tile_bytes);\n"
      << "            bulk_store(single + v0 * S3, s_single, tile_bytes);\n"
      << "            bulk_store(gt + v0 * NCOL, s_gt, (unsigned)(TB * NCOL));\n"
      << "            bulk_store(status + v0, s_status, (unsigned)TB);\n"
      << "            bulk_commit_and_wait_read(); // shared memory must stay alive until the engine has read it\n"
      << "        }\n"
      << "    } else {\n"
      << "        __syncthreads();\n"
      << "        for (int k = tid; k < nv * S3; k += TB) { post[v0 * S3 + k] = s_tile[k]; single[v0 * S3 + k] = s_single[k]; }\n"
      << "        for (int k = tid; k < nv * NCOL; k += TB) gt[v0 * NCOL + k] = s_gt[k];\n"
      << "        if (tid < nv) status[v0 + tid] = s_status[tid];\n"
      << "    }\n"
      << "}\n";
    return o.str();
}

int es_jit_build(const EsParams &P, std::string &cubin, std::string &log, std::string &err) {
    const int rc = gibbs_jit_compile(es_jit_source(P), cubin, log, err);
    if (rc != FS_OK) return rc;
    if (const char *v = std::getenv("FAMSEQ_JIT_VERBOSE"))
        if (v[0] == '1') std::fprintf(stderr, "[famseq] ES JIT: %d messages, %d input columns\n%s\n", P.prog.n_ops, P.C.s, log.c_str());
    return FS_OK;
}

struct EsJitKernel {
    cudaLibrary_t library = nullptr;
    cudaKernel_t kernel = nullptr;
    size_t smem = 0;
};

static size_t es_jit_smem(const EsParams &P) {
    const size_t S = (size_t)P.C.s;
    return 2 * kTB * S * 3 * sizeof(double) + ((kTB * S + 15) & ~(size_t)15) + kTB;
}

// Register-resident straight-line code only pays for pedigrees of moderate size: beyond ~24 sequenced members the
// likelihoods alone exceed the register file (the 100-member test pedigree compiles for minutes and spills 12 KB per
// thread); those stay with the interpreter.
bool es_jit_fits(const EsParams &P, size_t smem_limit) {
    return P.C.s <= 24 && P.prog.n_words <= 400 && es_jit_smem(P) + 64 <= smem_limit;
}

int es_jit_load(const EsParams &P, const std::string &cubin, EsJitKernel **out, std::string &err) {
    *out = nullptr;
    EsJitKernel *k = new EsJitKernel();
    k->smem = es_jit_smem(P);
    auto cuda_err = [&](cudaError_t e, const char *what) {
        err = std::string("ES JIT: ") + what + ": " + cudaGetErrorString(e);
        es_jit_unload(k);
        return FS_E_CUDA;
    };
    cudaError_t e = cudaLibraryLoadData(&k->library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) return cuda_err(e, "cudaLibraryLoadData");
    e = cudaLibraryGetKernel(&k->kernel, k->library, "famseq_es");
    if (e != cudaSuccess) return cuda_err(e, "cudaLibraryGetKernel");
    e = cudaFuncSetAttribute((const void *)k->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->smem);
    if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(shared memory)");
    *out = k;
    return FS_OK;
}

void es_jit_unload(EsJitKernel *k) {
    if (!k) return;
    if (k->library) cudaLibraryUnload(k->library);
    delete k;
}

cudaError_t es_jit_launch(EsJitKernel *k, const BatchPtrs &B, cudaStream_t stream) {
    if (B.V <= 0) return cudaSuccess;
    const int64_t n_tiles = (B.V + kTB - 1) / kTB;
    if (n_tiles > 0x7fffffff) return cudaErrorInvalidValue;
    const double *lk = B.lk;
    const uint8_t *flags = B.flags;
    double *post = B.post, *single = B.single;
    uint8_t *gt = B.gt, *status = B.status;
    long long V = B.V;
    void *args[] = {&lk, &flags, &post, &single, &gt, &status, &V};
    return cudaLaunchKernel((const void *)k->kernel, dim3((unsigned)n_tiles), dim3(kTB), args, k->smem, stream);
}

} // namespace famseq
