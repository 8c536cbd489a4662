// gibbs_jit.cu -- pedigree-specialised Gibbs sampler, generated and compiled at run time for sm_100a.
//
// Replaces family::calPostProbMCMC + estGenoProb (src/family.cpp:1932-2096, :2098-2299) for large batches.  The
// table-driven kernel of mcmc_kernel.cu spends ~220 instructions per Gibbs step, most of them decoding the pedigree
// (descriptor words, dynamic 2-bit genotype fields, padded child loops), and it keeps its 48 N bytes of chain state in
// global memory, which makes it HBM-bound (profiles/r1i).  Here the pedigree is known when the code is written:
//   * a sweep is ONE basic block of ~50 instructions per member: parents, children and spouses are named registers,
//     sex and founder cases are resolved at generation time, there is no descriptor, no loop over members or
//     children and no branch, so ptxas interleaves the Philox rounds, the table look-ups and the FP64 chains of
//     neighbouring members;
//   * a genotype is kept as the byte offset of its table row (g * 128): a transmission look-up is one integer
//     multiply-add and three LDS.64 with immediate offsets; consecutive full sibs share the look-up of their row;
//   * per chain, 3 own factors (1e6 * prior * lk) and 3 Rao-Blackwell accumulators per member are placed in registers,
//     in thread-private shared-memory columns, or in a block-private scratch that stays in L2 (own factors through a
//     software prefetch queue, accumulators through red.global.add.f64); gibbs_jit_default_config() has the measured
//     trade-offs.  The sweep loop causes no DRAM traffic.
// The kernel carries the sweep twice, with the autosomal and with the chrX rules (a thread takes the pair of loops of
// its variant).  The straight-line code covers what a sweep almost always is: weight sums that are positive normal
// numbers.  Chains in which a sum leaves that range are marked (status 2) and redone by the table-driven kernel,
// which the engine launches right behind this one (engine.cu) -- so nothing is approximated.
// The arithmetic is written with explicit round-to-nearest intrinsics in the order of mcmc_kernel.cu, and the Philox
// counters are the same: both kernels return the same bytes (tests/test_parity_gpu.py checks it), which is what makes
// it safe to compile in the background and switch kernels in the middle of a run.
#include <dlfcn.h>
#include <nvrtc.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

#include "../../../include/famseq_b200.h"
#include "gibbs_jit.hpp"

#define FS_STR2(x) #x
#define FS_STR(x) FS_STR2(x)

namespace famseq {

namespace {

constexpr int kCopies = 16;            // replicas of every table entry (bank spreading, as in mcmc_kernel.cu)
constexpr int kRow = kCopies * 8;      // bytes between consecutive table entries
constexpr int kTabBytes = 81 * kRow;   // autosome, X daughter, X son
constexpr size_t kSmemPerBlockMax = 227 * 1024;

int env_int(const char *name, int fallback) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : fallback;
}

std::string bits(double x) {
    unsigned long long u;
    std::memcpy(&u, &x, 8);
    char buf[40];
    std::snprintf(buf, sizeof buf, "0x%016llxull", u);
    return buf;
}

// ---- fixed text: types, helpers -------------------------------------------------------------------------------
const char *kPrelude = R"CUDA(
typedef unsigned int u32;
typedef unsigned long long u64;
typedef long long i64;
typedef unsigned char u8;

__device__ __forceinline__ double lds64(u32 a) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

__device__ __forceinline__ void philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 &o0, u32 &o1, u32 &o2, u32 &o3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// 1/s to ~1 ulp for a positive normal s away from the exponent limits: MUFU seed + two Newton steps (as mcmc_kernel.cu)
__device__ __forceinline__ double newton_reciprocal(double s) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(s));
    double e = __fma_rn(-s, x, 1.0);
    x = __fma_rn(x, e, x);
    e = __fma_rn(-s, x, 1.0);
    x = __fma_rn(x, e, x);
    return x;
}

// family::get_postRlt (family.cpp:636-665): strict '<' from -1, first maximum wins, NaN rows give -1
__device__ __forceinline__ u8 call_genotype(double p0, double p1, double p2) {
    double big = -1.0;
    int arg = -1;
    if (big < p0) { big = p0; arg = 0; }
    if (big < p1) { big = p1; arg = 1; }
    if (big < p2) { big = p2; arg = 2; }
    return (u8)arg;
}
)CUDA";

struct Member {
    int mother = 0, father = 0, col = -1;
    bool founder = false, male = false;
    struct Link {
        int child, other;
        bool child_male;
    };
    std::vector<Link> links;
};

std::vector<Member> decode(const McmcPlan &pl) {
    std::vector<Member> m(pl.n);
    for (int i = 0; i < pl.n; i++) {
        const uint32_t d = pl.member[i];
        m[i].mother = d & 127u;
        m[i].father = (d >> 7) & 127u;
        m[i].founder = (d >> 14) & 1u;
        m[i].male = (d >> 15) & 1u;
        m[i].col = pl.col[i];
        const int first = (d >> 16) & 0xffu, cnt = d >> 24;
        for (int k = first; k < first + cnt; k++) {
            const uint32_t l = pl.link[k];
            m[i].links.push_back({(int)(l & 127u), (int)((l >> 7) & 127u), ((l >> 14) & 1u) != 0});
        }
    }
    return m;
}

enum Place { REG, SMEM, GLOB };

// Where the 3 own factors and the 3 accumulators of every member live.  Members are assigned in ped order: registers
// first, then shared-memory rows, the rest in the block-private global scratch (own factors: read through a software
// prefetch queue; accumulators: fire-and-forget red.global.add.f64).
struct Layout {
    int n = 0;
    std::vector<Place> lk_place, acc_place;
    std::vector<int> lk_row, acc_row; // row (3-vector) in the shared-memory area or in the scratch
    int smem_rows = 0, glob_rows = 0;
    std::vector<int> lk_glob; // members whose own factors come from the scratch, in ped order
    int depth = 1;            // prefetch distance, in such members
};

// Members that go to the scratch are spread evenly over the sweep (so that a short prefetch queue covers the L2
// latency and the reductions do not bunch up); of the others, the first n_reg sit in registers, the rest in shared memory.
std::vector<Place> spread(int n, int n_reg, int n_smem) {
    std::vector<Place> place(n, SMEM);
    const int n_glob = std::max(0, n - n_reg - n_smem);
    for (int i = 0; i < n; i++)
        if ((long)(i + 1) * n_glob / n != (long)i * n_glob / n) place[i] = GLOB;
    int left = n_reg;
    for (int i = 0; i < n && left > 0; i++)
        if (place[i] != GLOB) {
            place[i] = REG;
            left--;
        }
    return place;
}

Layout make_layout(int n, const GibbsJitConfig &cfg) {
    Layout L;
    L.n = n;
    L.acc_place = spread(n, cfg.n_acc_reg, cfg.n_acc_smem);
    L.lk_place = spread(n, cfg.n_lk_reg, cfg.n_lk_smem);
    L.lk_row.assign(n, -1);
    L.acc_row.assign(n, -1);
    for (int i = 0; i < n; i++) {
        if (L.acc_place[i] == SMEM) L.acc_row[i] = L.smem_rows++;
        if (L.acc_place[i] == GLOB) L.acc_row[i] = L.glob_rows++;
    }
    for (int i = 0; i < n; i++) {
        if (L.lk_place[i] == SMEM) L.lk_row[i] = L.smem_rows++;
        if (L.lk_place[i] == GLOB) {
            L.lk_row[i] = L.glob_rows++;
            L.lk_glob.push_back(i);
        }
    }
    L.depth = std::max(1, cfg.prefetch);
    return L;
}

size_t smem_bytes(const Layout &L, int tb) { return (size_t)kTabBytes + (size_t)L.smem_rows * 3 * tb * 8; }

std::string smem_ref(int row, int g) { return "sa[" + std::to_string(row * 3 + g) + " * TB]"; }
std::string glob_ref(int row, int g) { return "wg + " + std::to_string(row * 3 + g) + " * TB"; }

// One Gibbs step of member i (family.cpp:2113-2178 / :2195-2295), as straight-line code.
// `row` names three variables that hold the member's transmission row T[.][mother][father] (loaded by emit_sweep,
// shared by consecutive full sibs).
// chrX sweeps (family.cpp:2183-2297): a member's own transmission comes from the table of its sex, a child's from the
// table of the child's sex, and only males get the children factor.
const char *table_of(bool chrx, bool male) { return chrx ? (male ? "tXM" : "tXF") : "tA"; }

void emit_member(std::ostringstream &o, const std::vector<Member> &M, const Layout &L, int i, bool accumulate, const std::string &row,
                 bool chrx) {
    const Member &m = M[i];
    o << "            { // member " << i << (m.founder ? " (founder" : " (child of ") ;
    if (!m.founder) o << m.mother << " x " << m.father;
    o << (m.male ? ", male)\n" : ", not male)\n");
    if (L.lk_place[i] == GLOB) {
        const int G = (int)L.lk_glob.size();
        const int k = (int)(std::find(L.lk_glob.begin(), L.lk_glob.end(), i) - L.lk_glob.begin());
        const int slot = k % L.depth, ahead = L.lk_glob[(k + L.depth) % G];
        o << "                double w0 = q" << slot << "_0, w1 = q" << slot << "_1, w2 = q" << slot << "_2;\n";
        o << "                q" << slot << "_0 = __ldcg(" << glob_ref(L.lk_row[ahead], 0) << "); q" << slot << "_1 = __ldcg("
          << glob_ref(L.lk_row[ahead], 1) << "); q" << slot << "_2 = __ldcg(" << glob_ref(L.lk_row[ahead], 2) << ");\n";
    } else if (L.lk_place[i] == SMEM) {
        o << "                double w0 = " << smem_ref(L.lk_row[i], 0) << ", w1 = " << smem_ref(L.lk_row[i], 1) << ", w2 = "
          << smem_ref(L.lk_row[i], 2) << ";\n";
    } else {
        o << "                double w0 = W" << i << "_0, w1 = W" << i << "_1, w2 = W" << i << "_2;\n";
    }
    if (!m.founder) // transmission from the parents' current genotypes
        o << "                w0 = __dmul_rn(w0, " << row << "_0); w1 = __dmul_rn(w1, " << row << "_1); w2 = __dmul_rn(w2, " << row << "_2);\n";
    for (const Member::Link &l : m.links) {
        if (chrx && !m.male) break; // family.cpp:2230-2257
        const char *tA = table_of(chrx, l.child_male);
        if (m.male) // this member is the father: entry child*9 + mother*3 + g
            o << "                { const u32 ta = " << tA << " + o" << l.child << " * 9u + o" << l.other << " * 3u;\n"
              << "                  w0 = __dmul_rn(w0, lds64(ta)); w1 = __dmul_rn(w1, lds64(ta + " << kRow
              << ")); w2 = __dmul_rn(w2, lds64(ta + " << 2 * kRow << ")); }\n";
        else // this member is the mother: entry child*9 + g*3 + father
            o << "                { const u32 ta = " << tA << " + o" << l.child << " * 9u + o" << l.other << ";\n"
              << "                  w0 = __dmul_rn(w0, lds64(ta)); w1 = __dmul_rn(w1, lds64(ta + " << 3 * kRow
              << ")); w2 = __dmul_rn(w2, lds64(ta + " << 6 * kRow << ")); }\n";
    }
    o << "                const double sum = __dadd_rn(__dadd_rn(w0, w1), w2);\n";
    if ((i & 3) == 0) o << "                philox((u32)sweep, " << (i >> 2) << "u, gv_lo, gv_hi, k0, k1, r0, r1, r2, r3);\n";
    // (u + 0.5) * 2^-32 without an int-to-double conversion: 1 + u * 2^-32 assembled from bits, then one exact subtraction
    o << "                const double rd = __dsub_rn(__hiloint2double((int)(0x3ff00000u | (r" << (i & 3) << " >> 12)), (int)(r" << (i & 3)
      << " << 20)), 0x1.ffffffffp-1);\n";
    o << "                const double thr = __dmul_rn(rd, sum);\n";
    // The straight-line code assumes a positive normal sum with exponent in [-963, 963) (no sign test, Newton reciprocal);
    // `worst` records whether that ever failed, in which case the chain is redone by the table-driven kernel.
    o << "                worst = max(worst, (u32)__double2hiint(sum) - 0x03c00000u);\n";
    o << "                o" << i << " = (thr < w0) ? 0u : ((thr > __dsub_rn(sum, w2)) ? " << 2 * kRow << "u : " << kRow << "u);\n";
    if (accumulate) {
        o << "                const double inv = newton_reciprocal(sum);\n";
        for (int g = 0; g < 3; g++) {
            const std::string term = "__dmul_rn(w" + std::to_string(g) + ", inv)";
            if (L.acc_place[i] == SMEM)
                o << "                " << smem_ref(L.acc_row[i], g) << " = __dadd_rn(" << smem_ref(L.acc_row[i], g) << ", " << term << ");\n";
            else if (L.acc_place[i] == GLOB)
                o << "                atomicAdd(" << glob_ref(L.acc_row[i], g) << ", " << term << ");\n";
            else
                o << "                A" << i << "_" << g << " = __dadd_rn(A" << i << "_" << g << ", " << term << ");\n";
        }
    }
    o << "            }\n";
}

// One sweep over the members in ped order.  A non-founder's own factor needs the row T[g][mother][father], g = 0..2
// (entry g*9 + mother*3 + father): full sibs that follow each other before either parent is updated again share one
// look-up -- three shared-memory loads saved per sib, and shared-memory bandwidth is what bounds this kernel.
void emit_sweep(std::ostringstream &o, const std::vector<Member> &M, const Layout &L, bool accumulate, bool chrx, const char *tag) {
    const int n = (int)M.size();
    std::vector<int> version(n, 0);
    struct Row {
        int mother, father, vm, vf;
        bool male;
        std::string name;
    };
    std::vector<Row> rows;
    for (int i = 0; i < n; i++) {
        std::string row;
        if (!M[i].founder) {
            const int mo = M[i].mother, fa = M[i].father;
            for (const Row &r : rows)
                if (r.mother == mo && r.father == fa && r.vm == version[mo] && r.vf == version[fa] && (!chrx || r.male == M[i].male))
                    row = r.name;
            if (row.empty()) {
                row = std::string("T") + tag + std::to_string(i);
                o << "            const u32 a" << row << " = " << table_of(chrx, M[i].male) << " + o" << mo << " * 3u + o" << fa << ";\n";
                o << "            const double " << row << "_0 = lds64(a" << row << "), " << row << "_1 = lds64(a" << row << " + " << 9 * kRow << "), "
                  << row << "_2 = lds64(a" << row << " + " << 18 * kRow << ");\n";
                rows.push_back({mo, fa, version[mo], version[fa], M[i].male, row});
            }
        }
        emit_member(o, M, L, i, accumulate, row, chrx);
        version[i]++;
    }
}

// End of a sweep: the loads in flight belong to the first members of the next sweep; put them where it expects them.
void emit_queue_rotation(std::ostringstream &o, const Layout &L) {
    const int G = (int)L.lk_glob.size(), D = L.depth;
    if (G == 0 || G % D == 0) return;
    o << "            { // prefetch queue: slot j of the next sweep is slot (j + " << G % D << ") % " << D << " of this one\n";
    for (int j = 0; j < D; j++)
        for (int g = 0; g < 3; g++) o << "                const double t" << j << "_" << g << " = q" << (G + j) % D << "_" << g << ";\n";
    for (int j = 0; j < D; j++)
        for (int g = 0; g < 3; g++) o << "                q" << j << "_" << g << " = t" << j << "_" << g << ";\n";
    o << "            }\n";
}

} // namespace

GibbsJitConfig gibbs_jit_default_config(const McmcParams &P) {
    const int n = P.plan.n;
    GibbsJitConfig c;
    // Two warps per SM sub-partition: 256 chains per SM, up to 255 registers each.  Measured on the 40-member pedigree
    // (profiles/jit_sweep*.sh): registers are the only free storage -- a shared-memory row costs LDS bandwidth (the unit
    // that bounds the kernel), an accumulator in L2 costs three reductions (the L2 sustains ~6.5e11 FP64 reductions/s
    // per GPU), own factors in L2 cost three loads.  So: accumulators in registers as far as they go (the step itself
    // needs ~124 + n), a sixth of the shared-memory rows for more accumulators, the rest of them for own factors, and
    // whatever is left in L2.
    c.tb = 256;
    c.blocks = 1;
    c.prefetch = 2;
    const int reg_rows = std::max(0, (254 - (124 + n)) / 6);
    const int smem_rows = (int)((kSmemPerBlockMax - kTabBytes) / ((size_t)24 * c.tb));
    c.n_acc_reg = std::min(n, reg_rows);
    c.n_lk_reg = std::min(n, reg_rows - c.n_acc_reg);
    c.n_acc_smem = std::min(n - c.n_acc_reg, smem_rows / 6);
    c.n_lk_smem = std::min(n - c.n_lk_reg, smem_rows - c.n_acc_smem);
    if (c.n_acc_reg == n && c.n_lk_reg == n) // small pedigree, everything in registers: more than one block per SM
        c.blocks = std::max(1, std::min(4, 65536 / (c.tb * (70 + n + 12 * n))));
    c.tb = env_int("FAMSEQ_JIT_TB", c.tb);
    c.blocks = std::max(1, env_int("FAMSEQ_JIT_BLOCKS", c.blocks));
    c.prefetch = std::max(1, env_int("FAMSEQ_JIT_PF", c.prefetch));
    c.n_acc_reg = std::min(n, std::max(0, env_int("FAMSEQ_JIT_RACC", c.n_acc_reg)));
    c.n_acc_smem = std::min(n - c.n_acc_reg, std::max(0, env_int("FAMSEQ_JIT_SACC", c.n_acc_smem)));
    c.n_lk_reg = std::min(n, std::max(0, env_int("FAMSEQ_JIT_RLK", c.n_lk_reg)));
    c.n_lk_smem = std::min(n - c.n_lk_reg, std::max(0, env_int("FAMSEQ_JIT_SLK", c.n_lk_smem)));
    return c;
}

std::string gibbs_jit_source(const McmcParams &P, const GibbsJitConfig &cfg) {
    const RunConstants &C = P.C;
    const std::vector<Member> M = decode(P.plan);
    const int n = (int)M.size(), S = C.s;
    const Layout L = make_layout(n, cfg);
    std::ostringstream o;
    o << "// generated by famseq_b200 (gibbs_jit.cu) for one pedigree: " << n << " members, " << S << " input columns\n";
    o << "// layout: " << cfg.tb << " chains per block; accumulators " << cfg.n_acc_reg << " reg / " << cfg.n_acc_smem << " smem / "
      << n - cfg.n_acc_reg - cfg.n_acc_smem << " L2; own factors " << cfg.n_lk_reg << " reg / " << cfg.n_lk_smem << " smem / "
      << L.lk_glob.size() << " L2 (prefetch " << L.depth << ")\n";
    o << "#define TB " << cfg.tb << "\n#define NCOL " << S << "\n";
    o << kPrelude;
    o << "__constant__ u64 TAB_BITS[81] = {";
    for (int t = 0; t < 3; t++)
        for (int k = 0; k < 27; k++) o << (t + k ? ", " : "") << bits(C.tab[t][k]);
    o << "};\n__constant__ u64 PRIOR_BITS[12] = {";
    for (int t = 0; t < 4; t++)
        for (int g = 0; g < 3; g++) o << (t + g ? ", " : "") << bits(C.prior[t][g]);
    o << "};\n__constant__ u8 COL_MALE[NCOL + 1] = {";
    for (int c = 0; c < S; c++) o << (int)C.col_male[c] << ", ";
    o << "0};\n";
    const unsigned unseq = (C.unseq_fail[0] ? 1u : 0u) | (C.unseq_fail[1] ? 2u : 0u) | (C.unseq_fail[2] ? 4u : 0u) | (C.unseq_fail[3] ? 8u : 0u);

    o << "\nextern \"C\" __global__ void __launch_bounds__(TB, " << cfg.blocks << ")\n"
      << "famseq_gibbs(const double *__restrict__ lk, const u8 *__restrict__ flags, double *__restrict__ post,\n"
      << "             double *__restrict__ single, u8 *__restrict__ gt, u8 *__restrict__ status, i64 V, int burn, int rep,\n"
      << "             u64 seed, i64 v_offset, double *scratch, int n_tiles) {\n"
      << "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
      << "    double *s_tab = (double *)smem_raw;              // [81][" << kCopies << "]\n"
      << "    double *s_vec = s_tab + 81 * " << kCopies << ";           // [" << L.smem_rows << " * 3][TB] thread-private columns\n"
      << "    const int tid = threadIdx.x, lane = tid & 31;\n"
      << "    for (int e = tid; e < 81 * " << kCopies << "; e += TB) s_tab[e] = __longlong_as_double((i64)TAB_BITS[e / " << kCopies << "]);\n"
      << "    __syncthreads();\n"
      << "    u32 tab_addr = (u32)__cvta_generic_to_shared(s_tab) + (u32)(lane & " << (kCopies - 1) << ") * 8u;\n"
      << "    asm volatile(\"\" : \"+r\"(tab_addr) :: \"memory\"); // table reads stay below the barrier\n"
      << "    double *sa = s_vec + tid;\n"
      << "    double *wg = scratch + (size_t)blockIdx.x * " << std::max(1, L.glob_rows) * 3 << " * TB + tid; // [row][g][TB], block-private\n"
      << "    const u32 k0 = (u32)seed, k1 = (u32)(seed >> 32);\n"
      << "    const double lrc = __longlong_as_double((i64)" << bits(C.lrc) << ");\n"
      << "    (void)sa; (void)wg;\n\n"
      << "    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {\n"
      << "        const i64 v = (i64)tile * TB + tid;\n"
      << "        if (v >= V) continue;\n"
      << "        const u32 flag = flags ? flags[v] : 0u;\n"
      << "        const bool known = flag & 1u, chrx = (flag >> 1) & 1u;\n"
      << "        const double pa0 = __longlong_as_double((i64)PRIOR_BITS[known ? 3 : 0]), pa1 = __longlong_as_double((i64)PRIOR_BITS[known ? 4 : 1]),\n"
      << "                     pa2 = __longlong_as_double((i64)PRIOR_BITS[known ? 5 : 2]);\n"
      << "        const double pm0 = chrx ? __longlong_as_double((i64)PRIOR_BITS[known ? 9 : 6]) : pa0, pm1 = chrx ? __longlong_as_double((i64)PRIOR_BITS[known ? 10 : 7]) : pa1,\n"
      << "                     pm2 = chrx ? __longlong_as_double((i64)PRIOR_BITS[known ? 11 : 8]) : pa2;\n"
      << "        const double *lkv = lk + v * (NCOL * 3);\n"
      << "        double *gp = post + v * (NCOL * 3), *gs = single + v * (NCOL * 3);\n"
      << "        u8 *gg = gt + v * NCOL;\n"
      << "        // individual-only posterior + LRC gate (family.cpp:1940-1971)\n"
      << "        bool failed = (" << unseq << "u >> (flag & 3u)) & 1u;\n"
      << "        bool pedigree_needed = false;\n"
      << "        for (int c = 0; c < NCOL; c++) {\n"
      << "            const double l0 = lkv[c * 3], l1 = lkv[c * 3 + 1], l2 = lkv[c * 3 + 2];\n"
      << "            const bool male = COL_MALE[c] != 0;\n"
      << "            const double q0 = __dmul_rn(l0, male ? pm0 : pa0), q1 = __dmul_rn(l1, male ? pm1 : pa1), q2 = __dmul_rn(l2, male ? pm2 : pa2);\n"
      << "            const double rs = __dadd_rn(__dadd_rn(q0, q1), q2);\n"
      << "            if (rs <= 0.0) failed = true;\n"
      << "            gs[c * 3] = __ddiv_rn(q0, rs); gs[c * 3 + 1] = __ddiv_rn(q1, rs); gs[c * 3 + 2] = __ddiv_rn(q2, rs);\n"
      << "            double big = 0.0;\n"
      << "            if (big < l0) big = l0;\n"
      << "            if (big < l1) big = l1;\n"
      << "            if (big < l2) big = l2;\n"
      << "            if (__ddiv_rn(big, __dadd_rn(__dadd_rn(l0, l1), l2)) < lrc) pedigree_needed = true;\n"
      << "        }\n"
      << "        if (failed) {\n"
      << "            for (int k = 0; k < NCOL * 3; k++) gp[k] = gs[k] = 0.0;\n"
      << "            for (int c = 0; c < NCOL; c++) gg[c] = 255;\n"
      << "            status[v] = 1;\n"
      << "            continue;\n"
      << "        }\n"
      << "        if (!pedigree_needed) { // family.cpp:1973-2058: FPP := individual-only posterior\n"
      << "            for (int c = 0; c < NCOL; c++) {\n"
      << "                const double p0 = gs[c * 3], p1 = gs[c * 3 + 1], p2 = gs[c * 3 + 2];\n"
      << "                gp[c * 3] = p0; gp[c * 3 + 1] = p1; gp[c * 3 + 2] = p2;\n"
      << "                gg[c] = call_genotype(p0, p1, p2);\n"
      << "            }\n"
      << "            status[v] = 0;\n"
      << "            continue;\n"
      << "        }\n"
      << "        const u32 tA = tab_addr, tXF = tab_addr + " << 27 * kRow << "u, tXM = tab_addr + " << 54 * kRow << "u;\n"
      << "        (void)tA; (void)tXF; (void)tXM;\n"
      << "        u32 worst = 0u;\n\n"
      << "        // chain state: own factors (1e6 * prior) * lk for founders, 1e6 * lk otherwise (family.cpp:2115-2126)\n";
    for (int i = 0; i < n; i++) {
        const Member &m = M[i];
        for (int g = 0; g < 3; g++) {
            std::ostringstream lkx, base;
            if (m.col >= 0)
                lkx << "lkv[" << m.col * 3 + g << "]";
            else
                lkx << "1.0";
            if (m.founder)
                base << "__dmul_rn(1000000.0, " << (m.male ? "pm" : "pa") << g << ")";
            else
                base << "1000000.0";
            const std::string value = "__dmul_rn(" + base.str() + ", " + lkx.str() + ")";
            if (L.lk_place[i] == GLOB)
                o << "        *(" << glob_ref(L.lk_row[i], g) << ") = " << value << ";\n";
            else if (L.lk_place[i] == SMEM)
                o << "        " << smem_ref(L.lk_row[i], g) << " = " << value << ";\n";
            else
                o << "        const double W" << i << "_" << g << " = " << value << ";\n";
            if (L.acc_place[i] == GLOB)
                o << "        *(" << glob_ref(L.acc_row[i], g) << ") = 0.0;\n";
            else if (L.acc_place[i] == SMEM)
                o << "        " << smem_ref(L.acc_row[i], g) << " = 0.0;\n";
            else
                o << "        double A" << i << "_" << g << " = 0.0;\n";
        }
    }
    o << "        const u64 gv = (u64)(v_offset + v);\n"
      << "        const u32 gv_lo = (u32)gv, gv_hi = (u32)(gv >> 32);\n"
      << "        u32 r0 = 0, r1 = 0, r2 = 0, r3 = 0;\n"
      << "        // initial genotypes (family.cpp:2063-2067), stored as table-row byte offsets g * " << kRow << "\n";
    for (int i = 0; i < n; i++) {
        if ((i & 3) == 0) o << "        philox(0u, " << (i >> 2) << "u, gv_lo, gv_hi, k0, k1, r0, r1, r2, r3);\n";
        o << "        u32 o" << i << " = (r" << (i & 3) << " % 3u) * " << kRow << "u;\n";
    }
    if (!L.lk_glob.empty()) {
        const int G = (int)L.lk_glob.size();
        o << "        // prefetch queue of own factors: " << L.depth << " member(s) ahead\n";
        for (int j = 0; j < L.depth; j++) {
            const int row = L.lk_row[L.lk_glob[j % G]];
            o << "        double q" << j << "_0 = __ldcg(" << glob_ref(row, 0) << "), q" << j << "_1 = __ldcg(" << glob_ref(row, 1) << "), q" << j
              << "_2 = __ldcg(" << glob_ref(row, 2) << ");\n";
        }
    }
    o << "\n        int sweep = 1;\n"
      << "        const int last = burn + rep;\n"
      << "        if (!chrx) {\n"
      << "        for (; sweep <= burn; sweep++) { // burn-in: no accumulation\n";
    emit_sweep(o, M, L, false, false, "b");
    emit_queue_rotation(o, L);
    o << "        }\n"
      << "        for (; sweep <= last; sweep++) { // sampling sweeps, Rao-Blackwellised (family.cpp:2175-2178)\n";
    emit_sweep(o, M, L, true, false, "s");
    emit_queue_rotation(o, L);
    o << "        }\n"
      << "        } else { // the same two loops with the chrX rules\n"
      << "        for (; sweep <= burn; sweep++) {\n";
    emit_sweep(o, M, L, false, true, "xb");
    emit_queue_rotation(o, L);
    o << "        }\n"
      << "        for (; sweep <= last; sweep++) {\n";
    emit_sweep(o, M, L, true, true, "xs");
    emit_queue_rotation(o, L);
    o << "        }\n"
      << "        }\n"
      << "        if (worst >= 0x78600000u) { status[v] = 2; continue; } // a sum left the fast range: redo with the table-driven kernel\n\n"
      << "        // postProb = genoFry / numRep, not renormalised; a row summing to <= 0 fails (family.cpp:2082-2092)\n"
      << "        const double nrep = (double)rep;\n";
    for (int i = 0; i < n; i++) {
        o << "        {\n";
        for (int g = 0; g < 3; g++) {
            std::string a;
            if (L.acc_place[i] == GLOB)
                a = "__ldcg(" + glob_ref(L.acc_row[i], g) + ")";
            else if (L.acc_place[i] == SMEM)
                a = smem_ref(L.acc_row[i], g);
            else
                a = "A" + std::to_string(i) + "_" + std::to_string(g);
            o << "            const double p" << g << " = __ddiv_rn(" << a << ", nrep);\n";
        }
        o << "            if (__dadd_rn(__dadd_rn(p0, p1), p2) <= 0.0) failed = true;\n";
        if (M[i].col >= 0) {
            const int c = M[i].col;
            o << "            gp[" << c * 3 << "] = p0; gp[" << c * 3 + 1 << "] = p1; gp[" << c * 3 + 2 << "] = p2;\n"
              << "            gg[" << c << "] = call_genotype(p0, p1, p2);\n";
        }
        o << "        }\n";
    }
    o << "        if (failed) {\n"
      << "            for (int k = 0; k < NCOL * 3; k++) gp[k] = gs[k] = 0.0;\n"
      << "            for (int c = 0; c < NCOL; c++) gg[c] = 255;\n"
      << "        }\n"
      << "        status[v] = failed ? 1 : 0;\n"
      << "    }\n"
      << "}\n";
    return o.str();
}

// ---- NVRTC, bound at run time so that the engine library itself has no link-time dependency on it ---------------
namespace {

struct Nvrtc {
    void *handle = nullptr;
    nvrtcResult (*create)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*compile)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*destroy)(nvrtcProgram *) = nullptr;
    nvrtcResult (*cubin_size)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*cubin)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*log_size)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*log)(nvrtcProgram, char *) = nullptr;
    const char *(*error_string)(nvrtcResult) = nullptr;
    std::string why;
};

const Nvrtc &nvrtc() {
    static Nvrtc N = [] {
        Nvrtc n;
        const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char *name : names)
            if ((n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
        if (!n.handle) {
            n.why = std::string("libnvrtc not found (") + dlerror() + ")";
            return n;
        }
        auto sym = [&](const char *s) { return dlsym(n.handle, s); };
        n.create = reinterpret_cast<decltype(n.create)>(sym("nvrtcCreateProgram"));
        n.compile = reinterpret_cast<decltype(n.compile)>(sym("nvrtcCompileProgram"));
        n.destroy = reinterpret_cast<decltype(n.destroy)>(sym("nvrtcDestroyProgram"));
        n.cubin_size = reinterpret_cast<decltype(n.cubin_size)>(sym("nvrtcGetCUBINSize"));
        n.cubin = reinterpret_cast<decltype(n.cubin)>(sym("nvrtcGetCUBIN"));
        n.log_size = reinterpret_cast<decltype(n.log_size)>(sym("nvrtcGetProgramLogSize"));
        n.log = reinterpret_cast<decltype(n.log)>(sym("nvrtcGetProgramLog"));
        n.error_string = reinterpret_cast<decltype(n.error_string)>(sym("nvrtcGetErrorString"));
        if (!n.create || !n.compile || !n.destroy || !n.cubin_size || !n.cubin || !n.log_size || !n.log || !n.error_string) {
            n.why = "libnvrtc lacks a required entry point";
            dlclose(n.handle);
            n.handle = nullptr;
        }
        return n;
    }();
    return N;
}

} // namespace

// Optional cubin cache (FAMSEQ_JIT_CACHE_DIR=<directory>): a run on a pedigree that was compiled before -- same
// generated source, same compiler -- loads the cubin from <directory>/<hash of the source>.cubin instead of compiling.
static std::string cache_file(const std::string &source) {
    const char *dir = std::getenv("FAMSEQ_JIT_CACHE_DIR");
    if (!dir || !*dir) return std::string();
    unsigned long long h = 1469598103934665603ull; // FNV-1a over the source and the compiler version
    auto mix = [&h](const char *p, size_t n) {
        for (size_t i = 0; i < n; i++) {
            h ^= (unsigned char)p[i];
            h *= 1099511628211ull;
        }
    };
    mix(source.data(), source.size());
    const char *ver = "nvrtc-" FS_STR(CUDART_VERSION) "-sm_100a";
    mix(ver, std::strlen(ver));
    char name[64];
    std::snprintf(name, sizeof name, "/%016llx.cubin", h);
    return std::string(dir) + name;
}

int gibbs_jit_compile(const std::string &source, std::string &cubin, std::string &log, std::string &err) {
    const std::string cached = cache_file(source);
    if (!cached.empty()) {
        if (FILE *f = std::fopen(cached.c_str(), "rb")) {
            std::fseek(f, 0, SEEK_END);
            const long n = std::ftell(f);
            std::fseek(f, 0, SEEK_SET);
            cubin.resize(n > 0 ? (size_t)n : 0);
            const size_t got = cubin.empty() ? 0 : std::fread(&cubin[0], 1, cubin.size(), f);
            std::fclose(f);
            if (got == cubin.size() && got > 0) {
                log = "cubin loaded from " + cached;
                return FS_OK;
            }
            cubin.clear();
        }
    }
    const Nvrtc &N = nvrtc();
    if (!N.handle) {
        err = "run-time compilation unavailable: " + N.why;
        return FS_E_CUDA;
    }
    nvrtcProgram prog = nullptr;
    nvrtcResult rc = N.create(&prog, source.c_str(), "famseq_gibbs.cu", 0, nullptr, nullptr);
    if (rc != NVRTC_SUCCESS) {
        err = std::string("nvrtcCreateProgram: ") + N.error_string(rc);
        return FS_E_CUDA;
    }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", "--ptxas-options=-v"};
    rc = N.compile(prog, (int)(sizeof opts / sizeof opts[0]), opts);
    size_t ls = 0;
    if (N.log_size(prog, &ls) == NVRTC_SUCCESS && ls > 1) {
        log.resize(ls);
        N.log(prog, &log[0]);
        while (!log.empty() && log.back() == '\0') log.pop_back();
    }
    if (rc != NVRTC_SUCCESS) {
        err = std::string("nvrtcCompileProgram: ") + N.error_string(rc) + "\n" + log;
        N.destroy(&prog);
        return FS_E_CUDA;
    }
    size_t cs = 0;
    rc = N.cubin_size(prog, &cs);
    if (rc == NVRTC_SUCCESS && cs > 0) {
        cubin.resize(cs);
        rc = N.cubin(prog, &cubin[0]);
    }
    N.destroy(&prog);
    if (rc != NVRTC_SUCCESS || cs == 0) {
        err = std::string("nvrtcGetCUBIN: ") + N.error_string(rc);
        return FS_E_CUDA;
    }
    if (!cached.empty()) { // best effort, atomic: write beside and rename
        const std::string tmp = cached + ".tmp" + std::to_string((long long)getpid());
        if (FILE *f = std::fopen(tmp.c_str(), "wb")) {
            const bool ok = std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
            std::fclose(f);
            if (!ok || std::rename(tmp.c_str(), cached.c_str()) != 0) std::remove(tmp.c_str());
        }
    }
    return FS_OK;
}

// ---- loaded kernel --------------------------------------------------------------------------------------------
struct GibbsJitKernel {
    cudaLibrary_t library = nullptr;
    cudaKernel_t kernel = nullptr;
    GibbsJitConfig cfg;
    size_t smem = 0;
    int glob_rows = 0;
    int blocks_per_sm = 1;
};

int gibbs_jit_load(const McmcParams &P, const GibbsJitConfig &cfg, const std::string &cubin, GibbsJitKernel **out, std::string &err) {
    *out = nullptr;
    const Layout L = make_layout(P.plan.n, cfg);
    const size_t smem = smem_bytes(L, cfg.tb);
    GibbsJitKernel *k = new GibbsJitKernel();
    k->cfg = cfg;
    k->smem = smem;
    k->glob_rows = L.glob_rows;
    auto cuda_err = [&](cudaError_t e, const char *what) {
        err = std::string("Gibbs JIT: ") + what + ": " + cudaGetErrorString(e);
        gibbs_jit_unload(k);
        return FS_E_CUDA;
    };
    cudaError_t e = cudaLibraryLoadData(&k->library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) return cuda_err(e, "cudaLibraryLoadData");
    e = cudaLibraryGetKernel(&k->kernel, k->library, "famseq_gibbs");
    if (e != cudaSuccess) return cuda_err(e, "cudaLibraryGetKernel");
    e = cudaFuncSetAttribute((const void *)k->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute(shared memory)");
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k->kernel, cfg.tb, smem);
    if (e != cudaSuccess) return cuda_err(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) return cuda_err(cudaErrorInvalidConfiguration, "kernel does not fit on an SM");
    k->blocks_per_sm = std::min(per_sm, cfg.blocks);
    *out = k;
    return FS_OK;
}

int gibbs_jit_build(const McmcParams &P, const GibbsJitConfig &cfg, std::string &cubin, std::string &log, std::string &err) {
    const Layout L = make_layout(P.plan.n, cfg);
    const size_t smem = smem_bytes(L, cfg.tb);
    if (cfg.tb < 32 || cfg.tb > 1024 || cfg.tb % 32 || smem > kSmemPerBlockMax) {
        err = "Gibbs JIT: invalid layout (tb " + std::to_string(cfg.tb) + ", shared memory " + std::to_string(smem) + " B)";
        return FS_E_TOO_LARGE;
    }
    const int rc = gibbs_jit_compile(gibbs_jit_source(P, cfg), cubin, log, err);
    if (rc != FS_OK) return rc;
    if (const char *v = std::getenv("FAMSEQ_JIT_VERBOSE"))
        if (v[0] == '1')
            std::fprintf(stderr, "[famseq] Gibbs JIT tb=%d blocks=%d acc %d reg/%d smem, own factors %d reg/%d smem, prefetch %d, smem=%zu\n%s\n",
                         cfg.tb, cfg.blocks, cfg.n_acc_reg, cfg.n_acc_smem, cfg.n_lk_reg, cfg.n_lk_smem, cfg.prefetch, smem, log.c_str());
    return FS_OK;
}

void gibbs_jit_unload(GibbsJitKernel *k) {
    if (!k) return;
    if (k->library) cudaLibraryUnload(k->library);
    delete k;
}

cudaError_t gibbs_jit_launch(GibbsJitKernel *k, const BatchPtrs &B, int burn, int rep, uint64_t seed, int64_t v_offset,
                             int sm_count, cudaStream_t stream) {
    if (B.V <= 0) return cudaSuccess;
    const int tb = k->cfg.tb;
    const int64_t n_tiles64 = (B.V + tb - 1) / tb;
    if (n_tiles64 > 0x7fffffff) return cudaErrorInvalidValue;
    int n_tiles = (int)n_tiles64;
    const int grid = (int)std::min<int64_t>(n_tiles64, (int64_t)sm_count * k->blocks_per_sm);
    double *scratch = nullptr; // own factors [grid][member][g][tb], stream-ordered
    cudaError_t rc = cudaMallocAsync(&scratch, (size_t)grid * std::max(1, k->glob_rows) * 3 * tb * sizeof(double), stream);
    if (rc != cudaSuccess) return rc;
    const double *lk = B.lk;
    const uint8_t *flags = B.flags;
    double *post = B.post, *single = B.single;
    uint8_t *gt = B.gt, *status = B.status;
    long long V = B.V, voff = v_offset;
    unsigned long long sd = seed;
    void *args[] = {&lk, &flags, &post, &single, &gt, &status, &V, &burn, &rep, &sd, &voff, &scratch, &n_tiles};
    rc = cudaLaunchKernel((const void *)k->kernel, dim3(grid), dim3(tb), args, k->smem, stream);
    const cudaError_t rc2 = cudaFreeAsync(scratch, stream);
    return rc != cudaSuccess ? rc : rc2;
}

} // namespace famseq
