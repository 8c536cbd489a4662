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
//   * one warp per block, one variant per thread, and every block walks a strided list of 32-variant tiles with the copies of
//     neighbouring tiles running under the arithmetic (see the kernel text below): tile k+1 arrives through a TMA bulk copy
//     while tile k is peeled, the `single` rows leave as soon as they exist and the posterior rows when the tile is done, and
//     nothing waits for a store except the one that needs its buffer back, thousands of cycles later.  (Round 1 / early round
//     2: one tile per block -- the warp sat idle for 18 % of its life waiting for its tile to arrive and for its stores to
//     drain, profiles/r2i_r2i_es14.txt, and with 255 registers there are only two warps per scheduler to cover for it.)
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

constexpr int kTB = 32; // variants per tile = threads per block: one warp (several warps per block were tried, profiles/r2l_es14.log)

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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// x[0..2] / s, correctly rounded: shared reciprocal + one Markstein correction, exponent-guarded (cuda/common.cuh: div3)
__device__ __forceinline__ bool safe_dividend(double x) {
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu, lo = (unsigned)__double2loint(x);
    return (hi - ((1023u - 900u) << 20)) <= (1800u << 20) || (hi | lo) == 0u;
}
// `care`: this lane's quotients will be used (a lane that has failed, or only keeps the others company through the pedigree,
// must not send the whole warp down the complete path).  The warp decides together -- all 32 lanes are here together by
// construction: a uniform branch needs no divergence bookkeeping, and the complete division is right for every lane
// whenever one lane needs it.
__device__ __forceinline__ void div3(double x0, double x1, double x2, double s, double &q0, double &q1, double &q2, bool care) {
    const bool in_range = ((unsigned)__double2hiint(s) - ((1023u - 900u) << 20)) <= (1800u << 20) & safe_dividend(x0) & safe_dividend(x1) &
                          safe_dividend(x2);
    const bool fast = __all_sync(0xffffffffu, in_range | !care);
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
    if (lrc == 1.0 && __all_sync(0xffffffffu, no_negative)) return big < ls;
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

// The point of a tile's program where its buffers change hands (once per tile, all 32 lanes together; the generator puts it
// in front of the first posterior row the program writes, a quarter of the way in).  By now the stores issued before it --
// the previous tile's posterior rows out of B, this tile's `single` rows out of A -- have long been read by the copy engine,
// so the wait costs nothing; B then starts as a copy of A (a member's posterior is its individual-only posterior until the
// pedigree says otherwise, family.cpp:1164-1253) and A is free for the next tile's likelihoods.
struct Pipe {
    double *a, *b;          // A: likelihood rows in, then `single` rows out; B: posterior rows out
    u64 *bar;
    const double *next_lk;  // the next full tile of this block, or nullptr
    int lane;
    bool done;
};
__device__ __forceinline__ void pipeline_point(Pipe &p) {
    if (p.done) return;
    p.done = true;
    if (p.lane == 0) bulk_wait_read();
    __syncwarp();
    for (int k = p.lane; k < TB * S3; k += TB) p.b[k] = p.a[k];
    fence_async_smem(); // order this lane's reads of A before the copy engine's writes to it
    __syncwarp();
    if (p.lane == 0 && p.next_lk) {
        mbar_expect_tx(p.bar, (unsigned)(TB * S3 * sizeof(double)));
        bulk_load(p.a, p.next_lk, (unsigned)(TB * S3 * sizeof(double)), p.bar);
    }
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
                    if (!point_emitted_) o_ << "    pipeline_point(pipe);\n";
                    point_emitted_ = true;
                    o_ << "    { double q0, q1, q2; div3(" << t << "m0, " << t << "m1, " << t << "m2, " << t << "sum, q0, q1, q2, mine);\n"
                       << "      if (mine) { row[" << col * 3 << "] = q0; row[" << col * 3 + 1 << "] = q1; row[" << col * 3 + 2 << "] = q2; gt_row[" << col
                       << "] = call_genotype(q0, q1, q2); } }\n";
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
    bool point_emitted_ = false;

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
        o << "\n// the message program with the " << (x ? "chrX" : "autosomal") << " rules, run by all 32 lanes together; lanes with `mine` keep their rows.\n"
          << "// Returns true when a member's row sum was exactly zero\n"
          << "__device__ __forceinline__ bool peel_" << (x ? "x" : "a")
          << "(Pipe &pipe, bool mine, double *row, u8 *gt_row, double pa0, double pa1, double pa2, double pm0, double pm1, double pm2";
        for (int c = 0; c < S; c++)
            for (int g = 0; g < 3; g++) o << ", double L" << c << "_" << g;
        o << ") {\n    bool failed = false;\n";
        Emitter(P, x != 0, o).run();
        o << "    return failed;\n}\n";
    }
    int min_blocks = 0; // tuning knob: resident blocks per SM the register allocation is forced to allow
    if (const char *env = std::getenv("FAMSEQ_ES_JIT_BLOCKS")) min_blocks = std::max(0, std::min(32, std::atoi(env)));
    o << "\n#define GT_BYTES ((TB * NCOL + 15) & ~15)\n"
      << "extern \"C\" __global__ void __launch_bounds__(TB" << (min_blocks ? ", " + std::to_string(min_blocks) : std::string()) << ")\n"
      << "famseq_es(const double *__restrict__ lk, const u8 *__restrict__ flags, double *__restrict__ post, double *__restrict__ single,\n"
      << "          u8 *__restrict__ gt, u8 *__restrict__ status, i64 V) {\n"
      << "    extern __shared__ __align__(128) unsigned char smem_raw[];\n"
      << "    double *s_a = (double *)smem_raw;   // [TB][S3] A: likelihood rows in, `single` rows out\n"
      << "    double *s_b = s_a + TB * S3;        // [TB][S3] B: posterior rows out\n"
      << "    u8 *s_gt2 = (u8 *)(s_b + TB * S3);  // [2][TB][NCOL]  (these two alternate between tiles: their stores are only\n"
      << "    u8 *s_status2 = s_gt2 + 2 * GT_BYTES; // [2][TB]       waited for one tile later)\n"
      << "    __shared__ u64 bar;\n"
      << "    const int lane = threadIdx.x;\n"
      << "    const unsigned tile_bytes = (unsigned)(TB * S3 * sizeof(double));\n"
      << "    const i64 n_tiles = (V + TB - 1) / TB, n_full = V / TB; // tiles below n_full move by TMA; a ragged last tile by plain loads / stores\n"
      << "    i64 tile = blockIdx.x;\n"
      << "    if (tile >= n_tiles) return;\n"
      << "    if (lane == 0) {\n"
      << "        mbar_init(&bar, 1);\n"
      << "        if (tile < n_full) { mbar_expect_tx(&bar, tile_bytes); bulk_load(s_a, lk + tile * TB * S3, tile_bytes, &bar); }\n"
      << "    }\n"
      << "    __syncwarp();\n"
      << "    u32 flag_next = (flags && tile * TB + lane < V) ? flags[tile * TB + lane] : 0u;\n"
      << "    unsigned phase = 0, it = 0;\n"
      << "    for (; tile < n_tiles; tile += gridDim.x, it ^= 1u) {\n"
      << "    const i64 v0 = tile * TB, next = tile + gridDim.x;\n"
      << "    const int nv = (int)((V - v0) < (i64)TB ? (V - v0) : (i64)TB);\n"
      << "    const bool full = nv == TB;\n"
      << "    const u32 flag = flag_next;\n"
      << "    flag_next = (flags && next * TB + lane < V) ? flags[next * TB + lane] : 0u; // arrives while this tile is peeled\n"
      << "    const bool known = flag & 1u, chrx = (flag >> 1) & 1u;\n";
    for (int g = 0; g < 3; g++)
        o << "    const double pa" << g << " = known ? " << lit(C.prior[1][g]) << " : " << lit(C.prior[0][g]) << ";\n"
          << "    const double pm" << g << " = chrx ? (known ? " << lit(C.prior[3][g]) << " : " << lit(C.prior[2][g]) << ") : pa" << g << ";\n";
    o << "    if (full) {\n"
      << "        mbar_wait(&bar, phase);\n"
      << "        phase ^= 1u;\n"
      << "    } else {\n"
      << "        // (A is free: the previous tile's pipeline point saw its `single` store off and fetched nothing)\n"
      << "        for (int k = lane; k < TB * S3; k += TB) s_a[k] = k < nv * S3 ? lk[v0 * S3 + k] : 0.0;\n"
      << "        __syncwarp();\n"
      << "    }\n"
      << "    Pipe pipe;\n"
      << "    pipe.a = s_a; pipe.b = s_b; pipe.bar = &bar; pipe.lane = lane; pipe.done = false;\n"
      << "    pipe.next_lk = next < n_full ? lk + next * TB * S3 : nullptr;\n"
      << "    double *row = s_b + lane * S3, *single_row = s_a + lane * S3;\n"
      << "    u8 *s_gt = s_gt2 + it * GT_BYTES, *s_status = s_status2 + it * TB;\n"
      << "    u8 *gt_row = s_gt + lane * NCOL;\n"
      << "    const bool live = lane < nv;\n";
    for (int c = 0; c < S; c++)
        o << "    const double L" << c << "_0 = single_row[" << c * 3 << "], L" << c << "_1 = single_row[" << c * 3 + 1 << "], L" << c << "_2 = single_row[" << c * 3 + 2
          << "];\n";
    o << "    // individual-only posterior (family.cpp:1405-1499) and LRC gate (family.cpp:1140-1162)\n"
      << "    bool failed = (" << unseq << "u >> (flag & 3u)) & 1u;\n"
      << "    bool pedigree_needed = false;\n"
      << "    const double lrc = " << lit(C.lrc) << ";\n";
    for (int c = 0; c < S; c++) {
        const char *pr = C.col_male[c] ? "pm" : "pa";
        o << "    {\n"
          << "        const double r0 = __dmul_rn(L" << c << "_0, " << pr << "0), r1 = __dmul_rn(L" << c << "_1, " << pr << "1), r2 = __dmul_rn(L" << c << "_2, " << pr << "2);\n"
          << "        const double rs = __dadd_rn(__dadd_rn(r0, r1), r2);\n"
          << "        if (rs <= 0.0) failed = true;\n"
          << "        double q0, q1, q2; div3(r0, r1, r2, rs, q0, q1, q2, !(rs <= 0.0));\n"
          << "        single_row[" << c * 3 << "] = q0; single_row[" << c * 3 + 1 << "] = q1; single_row[" << c * 3 + 2 << "] = q2;\n"
          << "        gt_row[" << c << "] = call_genotype(q0, q1, q2);\n"
          << "        double big = 0.0;\n"
          << "        if (big < L" << c << "_0) big = L" << c << "_0;\n"
          << "        if (big < L" << c << "_1) big = L" << c << "_1;\n"
          << "        if (big < L" << c << "_2) big = L" << c << "_2;\n"
          << "        if (lrc_wants_pedigree(lrc, L" << c << "_0, L" << c << "_1, L" << c << "_2, big, __dadd_rn(__dadd_rn(L" << c << "_0, L" << c << "_1), L" << c
          << "_2))) pedigree_needed = true;\n"
          << "    }\n";
    }
    auto call = [&](const char *fn, const char *mine) {
        o << fn << "(pipe, " << mine << ", row, gt_row, pa0, pa1, pa2, pm0, pm1, pm2";
        for (int c = 0; c < S; c++)
            for (int g = 0; g < 3; g++) o << ", L" << c << "_" << g;
        o << ")";
    };
    o << "    const bool failed_early = failed;\n"
      << "    if (failed) for (int k = 0; k < S3; k++) single_row[k] = 0.0;\n"
      << "    fence_async_smem(); // this lane's shared-memory writes, for the copy engine\n"
      << "    __syncwarp();\n"
      << "    if (full && lane == 0) { bulk_store(single + v0 * S3, s_a, tile_bytes); bulk_commit(); }\n"
      << "    // the pedigree: the warp runs a rule set when any of its variants wants it, and only those variants keep the rows\n"
      << "    const bool want = live && !failed && pedigree_needed;\n"
      << "    const bool mine_a = want && !chrx, mine_x = want && chrx;\n"
      << "    if (__any_sync(0xffffffffu, mine_a)) { const bool f = ";
    call("peel_a", "mine_a");
    o << "; if (mine_a && f) failed = true; }\n"
      << "    if (__any_sync(0xffffffffu, mine_x)) { const bool f = ";
    call("peel_x", "mine_x");
    o << "; if (mine_x && f) failed = true; }\n"
      << "    pipeline_point(pipe); // (a tile the gate kept the pedigree out of altogether)\n"
      << "    if (failed) {\n"
      << "        for (int k = 0; k < S3; k++) row[k] = 0.0;\n"
      << "        for (int c = 0; c < NCOL; c++) gt_row[c] = 255;\n"
      << "    }\n"
      << "    s_status[lane] = failed ? 1 : 0;\n"
      << "    if (full) {\n"
      << "        fence_async_smem();\n"
      << "        __syncwarp();\n"
      << "        if (lane == 0) {\n"
      << "            bulk_store(post + v0 * S3, s_b, tile_bytes);\n"
      << "            bulk_store(gt + v0 * NCOL, s_gt, (unsigned)(TB * NCOL));\n"
      << "            bulk_store(status + v0, s_status, (unsigned)TB);\n"
      << "            bulk_commit();\n"
      << "        }\n"
      << "        // a row sum that vanished inside the pedigree (family.cpp:1376-1384) clears `single` too, and that row is on its way\n"
      << "        // out already: let the copy land, then overwrite it\n"
      << "        if (__any_sync(0xffffffffu, failed && !failed_early)) {\n"
      << "            if (lane == 0) bulk_wait_all();\n"
      << "            __syncwarp();\n"
      << "            if (failed && !failed_early) for (int k = 0; k < S3; k++) single[(v0 + lane) * S3 + k] = 0.0;\n"
      << "        }\n"
      << "    } else {\n"
      << "        __syncwarp();\n"
      << "        if (failed && !failed_early) for (int k = 0; k < S3; k++) single_row[k] = 0.0;\n"
      << "        __syncwarp();\n"
      << "        for (int k = lane; k < nv * S3; k += TB) { post[v0 * S3 + k] = s_b[k]; single[v0 * S3 + k] = s_a[k]; }\n"
      << "        for (int k = lane; k < nv * NCOL; k += TB) gt[v0 * NCOL + k] = s_gt[k];\n"
      << "        if (live) status[v0 + lane] = s_status[lane];\n"
      << "    }\n"
      << "    }\n"
      << "    if (lane == 0) bulk_wait_read(); // shared memory must stay alive until the copy engine has read it\n"
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
    int64_t max_blocks = 0; // resident blocks of the whole device: the grid of the tile loop
};

static size_t es_jit_smem(const EsParams &P) {
    const size_t S = (size_t)P.C.s;
    return 2 * kTB * S * 3 * sizeof(double) + 2 * (((kTB * S + 15) & ~(size_t)15) + kTB);
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
    int device = 0, n_sm = 0, per_sm = 0;
    e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k->kernel, kTB, k->smem);
    if (e != cudaSuccess) return cuda_err(e, "occupancy query");
    if (per_sm < 1) {
        err = "ES JIT: the generated kernel does not fit on a multiprocessor";
        es_jit_unload(k);
        return FS_E_CUDA;
    }
    if (const char *env = std::getenv("FAMSEQ_ES_JIT_GRID")) per_sm = std::max(1, std::min(per_sm * 4, std::atoi(env))); // tuning: blocks per SM in the grid
    k->max_blocks = (int64_t)n_sm * per_sm;
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
    const int64_t grid = std::min(n_tiles, k->max_blocks);
    const double *lk = B.lk;
    const uint8_t *flags = B.flags;
    double *post = B.post, *single = B.single;
    uint8_t *gt = B.gt, *status = B.status;
    long long V = B.V;
    void *args[] = {&lk, &flags, &post, &single, &gt, &status, &V};
    return cudaLaunchKernel((const void *)k->kernel, dim3((unsigned)grid), dim3(kTB), args, k->smem, stream);
}

} // namespace famseq
