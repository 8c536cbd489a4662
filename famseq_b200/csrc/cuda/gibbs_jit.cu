// gibbs_jit.cu -- pedigree-specialised Gibbs sampler, generated and compiled at run time for sm_100a.
//
// Replaces family::calPostProbMCMC + estGenoProb (src/family.cpp:1932-2096, :2098-2299) for large batches.  The
// table-driven kernel of mcmc_kernel.cu spends ~220 instructions per Gibbs step, most of them decoding the pedigree
// (descriptor words, dynamic 2-bit genotype fields, padded child loops), and it keeps its chain state in global memory.
// Here the pedigree is known when the code is written: parents, children and spouses are named registers, sex and
// founder cases are resolved at generation time, a genotype is kept as the byte offset of its table row (g * 128), a
// transmission look-up is one integer multiply-add and three LDS.64 with immediate offsets.
//
// Round 2: CACHED FULL CONDITIONALS, INTEGER DRAWS, SPECULATIVE GROUPS.
// A member's three conditional weights depend on the current genotypes of its parents, children and spouses only.  On
// sequencing data most chains sit in one state nearly all the time (a Gibbs step changes the member's genotype in 8e-5 of
// the steps of the benchmark pedigrees), so recomputing the weights in every step -- what the reference does, and what
// round 1's kernel did at ~50 instructions, 23 of them FP64, and 14-17 shared-memory wavefronts per step -- is almost
// always redundant.  Every chain keeps, per member, the outcome of the last evaluation and a dirty bit:
//   * the draw rule rd < P0 -> 0, rd > 1 - P2 -> 2 (family.cpp:2161-2173) with rd = (u + 0.5) 2^-31, u the top 31 bits of
//     the member's Philox word, is EXACTLY u < T0 -> 0, u >= T2 -> 2 for the integers T0 = ceil(P0 2^31 - 1/2),
//     T2 = floor((1 - P2) 2^31 - 1/2) + 1 in [0, 2^31]; the two thresholds (two 32-bit registers per member) are all a
//     draw needs, and a draw is four integer instructions -- no FP64, no shared memory;
//   * members are taken in groups of eight.  If no member of the group is dirty, the eight draws are made side by side
//     from the cached thresholds (independent instructions: this is where the kernel gets its instruction-level
//     parallelism back); if then none of the changed members has a neighbour inside the group, the group commits:
//     genotypes replaced, neighbours of changed members marked dirty (constant masks).  Otherwise -- a dirty member, or a
//     change next to a group mate; about one group in ten per warp on the benchmark data -- the group is redone member by
//     member in the reference's order: evaluate if dirty (an out-of-line function per member: close the run of the old
//     weights, recompute, store), draw, mark.  The random words are the same on both paths, so the chain is the one the
//     sweep-by-sweep sampler produces with this stream.
//   * Rao-Blackwellisation (family.cpp:2175-2178) adds the SAME normalised weights once per sampling sweep for as long as
//     they stay valid; the kernel adds n * P when the run of n sampling sweeps ends instead of n times P (one rounding per
//     run instead of n: differs from sweep-by-sweep accumulation by ~1e-14 relative, far inside the 1e-9 parity
//     tolerance; tests compare the two kernels to 1e-12).  Accumulators, P0..P2 and the sweep index of the last
//     evaluation live in a block-private scratch that stays in L2 and is touched only when a run ends.
// The kernel carries the sweep twice, with the autosomal and with the chrX rules (a thread takes the loop of its
// variant).  The straight-line code covers what a sweep almost always is: weight sums that are positive normal
// numbers.  Chains in which a sum leaves that range are marked (status 2) and redone by the table-driven kernel,
// which the engine launches right behind this one (engine.cu) -- so nothing is approximated.
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

// ================================================================================================================
// Generator 1: cached full conditionals (the default where most members are sequenced; see the header of this file)
// ================================================================================================================
namespace cached {

// Where the cached draw thresholds (T0, T2: two 32-bit words) of every member live: the first n_p_reg members in registers,
// the others in thread-private shared-memory columns.  Per member, seven rows of the block-private global scratch hold the
// three accumulators, P0, P1, P2 and the sweep index of the last evaluation.
struct Layout {
    int n = 0, n_reg = 0, smem_pairs = 0;
    std::vector<int> pair; // shared-memory pair index of the member, -1 = registers
};

Layout make_layout(int n, const GibbsJitConfig &cfg) {
    Layout L;
    L.n = n;
    L.n_reg = std::max(0, std::min(n, cfg.n_p_reg));
    L.pair.assign(n, -1);
    for (int i = L.n_reg; i < n; i++) L.pair[i] = L.smem_pairs++;
    return L;
}

constexpr int kScratchRows = 7; // A0 A1 A2 P0 P1 P2 LAST
constexpr int kGroup = 8;       // members drawn side by side; divides 32, so a group lies within one word of the dirty mask
size_t smem_bytes(const Layout &L, int tb) { return (size_t)kTabBytes + (size_t)L.smem_pairs * 2 * tb * 4; }

std::string t0_ref(const Layout &L, int i) { return L.pair[i] < 0 ? "T" + std::to_string(i) : "st[" + std::to_string(L.pair[i] * 2) + " * TB]"; }
std::string t2_ref(const Layout &L, int i) { return L.pair[i] < 0 ? "U" + std::to_string(i) : "st[" + std::to_string(L.pair[i] * 2 + 1) + " * TB]"; }
std::string scratch_ref(int i, int k) { return "wg + " + std::to_string(i * kScratchRows + k) + " * TB"; }
enum { ROW_A0 = 0, ROW_P0 = 3, ROW_P1 = 4, ROW_P2 = 5, ROW_LAST = 6 };

// Members whose weights depend on the genotype of member j: its parents (j is one of their children), its children (their
// own transmission) and its spouses (co-parents of its children).  The chrX sweep skips the children factor of females,
// so its true sets are subsets of these; a superset only costs an evaluation.
std::vector<std::vector<int>> neighbours(const std::vector<Member> &M) {
    const int n = (int)M.size();
    std::vector<std::vector<int>> nb(n);
    auto add = [&](int j, int who) {
        if (who != j && std::find(nb[j].begin(), nb[j].end(), who) == nb[j].end()) nb[j].push_back(who);
    };
    for (int i = 0; i < n; i++) {
        if (!M[i].founder) {
            add(i, M[i].mother), add(i, M[i].father); // i changes -> its parents' children factors change
            add(M[i].mother, i), add(M[i].father, i); // a parent changes -> i's own transmission changes
        }
        for (const Member::Link &l : M[i].links) add(i, l.other), add(l.other, i);
    }
    return nb;
}

const char *table_of(bool chrx, bool male) { return chrx ? (male ? "tXM" : "tXF") : "tA"; }

// Members whose current genotype the evaluation of member i reads (generation-time constant).
std::vector<int> inputs_of(const std::vector<Member> &M, int i, bool chrx) {
    std::vector<int> in;
    auto add = [&](int k) {
        if (std::find(in.begin(), in.end(), k) == in.end()) in.push_back(k);
    };
    const Member &m = M[i];
    if (!m.founder) add(m.mother), add(m.father);
    if (!(chrx && !m.male))
        for (const Member::Link &l : m.links) add(l.child), add(l.other);
    std::sort(in.begin(), in.end());
    return in;
}

std::string eval_name(int i, bool chrx) { return std::string(chrx ? "eval_x_" : "eval_a_") + std::to_string(i); }

// The evaluation of member i's full conditional (family.cpp:2113-2178 / :2195-2295) as a function of its own, kept out of
// line (the sweep loop then holds only draws and call sites): closes the run of the old weights, forms the three weights
// from the member's likelihood and its parents', children's and spouses' current genotypes, normalises them, stores them
// and returns the draw thresholds.
// chrX sweeps (family.cpp:2183-2297): a member's own transmission comes from the table of its sex, a child's from the
// table of the child's sex, and only males get the children factor.
void emit_evaluation_function(std::ostringstream &o, const std::vector<Member> &M, int i, bool chrx) {
    const Member &m = M[i];
    const char *ind = "    ";
    o << "__device__ __noinline__ Rc " << eval_name(i, chrx) << "(const double *__restrict__ lkv, double *wg, double tD, double first, u32 tA, u32 tXF, u32 tXM, "
      << "u32 worst";
    if (m.founder) o << ", double " << (m.male ? "pm0, double pm1, double pm2" : "pa0, double pa1, double pa2");
    for (int k : inputs_of(M, i, chrx)) o << ", u32 o" << k;
    o << ") {\n    (void)tA; (void)tXF; (void)tXM;\n";
    // close the run of the old weights: they were valid in sweeps LAST .. sweep-1, of which `run` are sampling sweeps
    o << ind << "{ const double run = fmax(0.0, __dsub_rn(tD, fmax(__ldcg(" << scratch_ref(i, ROW_LAST) << "), first)));\n";
    for (int g = 0; g < 3; g++)
        o << ind << "  *(" << scratch_ref(i, ROW_A0 + g) << ") = __dadd_rn(__ldcg(" << scratch_ref(i, ROW_A0 + g) << "), __dmul_rn(run, __ldcg("
          << scratch_ref(i, ROW_P0 + g) << ")));\n";
    o << ind << "}\n";
    // own factor (1e6 * prior) * lk for founders, 1e6 * lk otherwise (family.cpp:2115-2126)
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
        o << ind << "double w" << g << " = __dmul_rn(" << base.str() << ", " << lkx.str() << ");\n";
    }
    if (!m.founder) // transmission from the parents' current genotypes: entry g*9 + mother*3 + father
        o << ind << "{ const u32 ta = " << table_of(chrx, m.male) << " + o" << m.mother << " * 3u + o" << m.father << ";\n"
          << ind << "  w0 = __dmul_rn(w0, lds64(ta)); w1 = __dmul_rn(w1, lds64(ta + " << 9 * kRow << ")); w2 = __dmul_rn(w2, lds64(ta + "
          << 18 * kRow << ")); }\n";
    for (const Member::Link &l : m.links) {
        if (chrx && !m.male) break; // family.cpp:2230-2257
        const char *tA = table_of(chrx, l.child_male);
        if (m.male) // this member is the father: entry child*9 + mother*3 + g
            o << ind << "{ const u32 ta = " << tA << " + o" << l.child << " * 9u + o" << l.other << " * 3u;\n"
              << ind << "  w0 = __dmul_rn(w0, lds64(ta)); w1 = __dmul_rn(w1, lds64(ta + " << kRow << ")); w2 = __dmul_rn(w2, lds64(ta + " << 2 * kRow
              << ")); }\n";
        else // this member is the mother: entry child*9 + g*3 + father
            o << ind << "{ const u32 ta = " << tA << " + o" << l.child << " * 9u + o" << l.other << ";\n"
              << ind << "  w0 = __dmul_rn(w0, lds64(ta)); w1 = __dmul_rn(w1, lds64(ta + " << 3 * kRow << ")); w2 = __dmul_rn(w2, lds64(ta + " << 6 * kRow
              << ")); }\n";
    }
    o << ind << "const double sum = __dadd_rn(__dadd_rn(w0, w1), w2);\n";
    // The straight-line code assumes a positive normal sum with exponent in [-963, 963) (no sign test, Newton reciprocal);
    // `worst` records whether that ever failed, in which case the chain is redone by the table-driven kernel.
    o << ind << "worst = max(worst, (u32)__double2hiint(sum) - 0x03c00000u);\n";
    o << ind << "const double inv = newton_reciprocal(sum);\n";
    o << ind << "const double p0 = __dmul_rn(w0, inv), p1 = __dmul_rn(w1, inv), p2 = __dmul_rn(w2, inv);\n";
    o << ind << "*(" << scratch_ref(i, ROW_P0) << ") = p0; *(" << scratch_ref(i, ROW_P1) << ") = p1; *(" << scratch_ref(i, ROW_P2) << ") = p2; *("
      << scratch_ref(i, ROW_LAST) << ") = tD;\n";
    // rd < p0  <=>  u < ceil(p0 2^31 - 1/2);   rd > 1 - p2  <=>  u >= floor((1 - p2) 2^31 - 1/2) + 1     (rd = (u + 1/2) 2^-31, exact)
    o << ind << "Rc r;\n"
      << ind << "r.t0 = __double2uint_ru(__fma_rn(p0, 2147483648.0, -0.5));\n"
      << ind << "r.t2 = (u32)(__double2int_rd(__fma_rn(__dsub_rn(1.0, p2), 2147483648.0, -0.5)) + 1);\n"
      << ind << "r.worst = worst;\n"
      << ind << "return r;\n}\n";
}

std::vector<unsigned> neighbour_mask(const std::vector<std::vector<int>> &nb, int i, int words) {
    std::vector<unsigned> mask(words, 0u);
    for (int j : nb[i]) mask[j / 32] |= 1u << (j % 32);
    return mask;
}

// One member the reference's way: evaluate if a neighbour changed, draw, mark the neighbours if the genotype changed.
void emit_member_serial(std::ostringstream &o, const std::vector<Member> &M, const Layout &L, const std::vector<std::vector<int>> &nb, int i, bool chrx,
                        const std::string &u) {
    const Member &m = M[i];
    const int n = (int)M.size(), words = (n + 31) / 32;
    const std::string dw = "D" + std::to_string(i / 32), bit = std::to_string(1u << (i % 32)) + "u";
    o << "                if (" << dw << " & " << bit << ") { // member " << i << ": a neighbour changed since its weights were evaluated\n"
      << "                    " << dw << " &= ~" << bit << ";\n"
      << "                    const Rc r = " << eval_name(i, chrx) << "(lkv, wg, tD, first, tA, tXF, tXM, worst";
    if (m.founder) o << (m.male ? ", pm0, pm1, pm2" : ", pa0, pa1, pa2");
    for (int k : inputs_of(M, i, chrx)) o << ", o" << k;
    o << ");\n                    " << t0_ref(L, i) << " = r.t0; " << t2_ref(L, i) << " = r.t2; worst = r.worst;\n                }\n";
    o << "                { const u32 g = (" << u << " < " << t0_ref(L, i) << ") ? 0u : ((" << u << " >= " << t2_ref(L, i) << ") ? " << 2 * kRow << "u : " << kRow
      << "u);\n                  const u32 changed = (g != o" << i << ") ? 0xffffffffu : 0u;\n";
    const std::vector<unsigned> mask = neighbour_mask(nb, i, words);
    for (int w = 0; w < words; w++)
        if (mask[w]) o << "                  D" << w << " |= changed & " << mask[w] << "u;\n";
    o << "                  o" << i << " = g; }\n";
}

// One sweep: groups of kGroup members, each drawn side by side when it can be, member by member when it must be.  The choice
// is made per WARP (a vote): control flow stays uniform, so the lanes of a warp never drift apart (with a per-lane branch the
// rare member-by-member lanes were left behind for good and the warp degenerated into single lanes: 20 times the
// instructions, profiles/r2e_r2e_mcmc.txt).  For a lane whose own group was fine, the member-by-member path evaluates nothing
// and draws the same genotypes from the same thresholds and random words.
void emit_sweep(std::ostringstream &o, const std::vector<Member> &M, const Layout &L, const std::vector<std::vector<int>> &nb, bool chrx) {
    const int n = (int)M.size(), words = (n + 31) / 32;
    for (int a = 0; a < n; a += kGroup) {
        const int b = std::min(n, a + kGroup), w = a / 32;
        unsigned gmask = 0;
        for (int i = a; i < b; i++) gmask |= 1u << (i % 32);
        o << "            { // members " << a << " .. " << b - 1 << "\n";
        // the random words of the group (word i%4 of Philox block i/4), kept in their own variables so that both paths read the same
        for (int i = a; i < b; i++)
            if ((i & 3) == 0 || i == a) {
                const int blk = i >> 2;
                o << "                u32 x" << blk << "_0, x" << blk << "_1, x" << blk << "_2, x" << blk << "_3;\n"
                  << "                philox((u32)sweep, " << blk << "u, gv_lo, gv_hi, k0, k1, x" << blk << "_0, x" << blk << "_1, x" << blk << "_2, x" << blk
                  << "_3);\n";
            }
        auto u_of = [](int i) { return "(x" + std::to_string(i >> 2) + "_" + std::to_string(i & 3) + " >> 1)"; };
        // the draws of the group, side by side, from the cached thresholds (only used if the group turns out to be clean)
        for (int i = a; i < b; i++)
            o << "                const u32 g" << i << " = (" << u_of(i) << " < " << t0_ref(L, i) << ") ? 0u : ((" << u_of(i) << " >= " << t2_ref(L, i) << ") ? "
              << 2 * kRow << "u : " << kRow << "u);\n";
        std::vector<bool> word_used(words, false);
        for (int ww = 0; ww < words; ww++) {
            for (int i = a; i < b; i++) word_used[ww] = word_used[ww] || neighbour_mask(nb, i, words)[ww];
            if (!word_used[ww]) continue;
            o << "                const u32 n" << ww << " = 0u";
            for (int i = a; i < b; i++) {
                const unsigned mk = neighbour_mask(nb, i, words)[ww];
                if (mk) o << " | ((g" << i << " != o" << i << ") ? " << mk << "u : 0u)";
            }
            o << ";\n";
        }
        // fine for this lane: nobody in the group is dirty, and no member that changes has a neighbour inside the group
        o << "                const bool fine = ((D" << w << (word_used[w] ? " | n" + std::to_string(w) : std::string()) << ") & " << gmask << "u) == 0u;\n"
          << "                if (__all_sync(live, fine)) { // commit\n";
        for (int i = a; i < b; i++) o << "                    o" << i << " = g" << i << ";\n";
        for (int ww = 0; ww < words; ww++)
            if (word_used[ww]) o << "                    D" << ww << " |= n" << ww << ";\n";
        o << "                } else { // member by member, in the reference's order\n"
          << "                    n_serial += sweep > burn ? 1u : 0u;\n";
        for (int i = a; i < b; i++) emit_member_serial(o, M, L, nb, i, chrx, u_of(i));
        o << "                }\n            }\n";
    }
}

GibbsJitConfig default_config(const McmcParams &P) {
    const int n = P.plan.n;
    GibbsJitConfig c;
    c.cached = 1;
    // A chain needs ~60 registers for the step itself, one per member for the genotypes and two per member whose draw
    // thresholds sit in registers; the other members' pairs go to thread-private shared-memory columns (8 bytes per member
    // and chain).  The block is as large as the register file allows (the step is a chain of short integer dependencies:
    // warps hide it), in multiples of 128 threads.
    c.blocks = 1;
    c.n_p_reg = std::max(0, std::min(n, (250 - (64 + n)) / 2));
    const int regs = 64 + n + 2 * c.n_p_reg;
    c.tb = regs <= 124 ? 512 : (regs <= 164 ? 384 : 256);
    while (c.tb > 32 && kTabBytes + (size_t)(n - c.n_p_reg) * 8 * c.tb > kSmemPerBlockMax) c.tb -= 32;
    c.tb = env_int("FAMSEQ_JIT_TB", c.tb);
    c.blocks = std::max(1, env_int("FAMSEQ_JIT_BLOCKS", c.blocks));
    c.n_p_reg = std::min(n, std::max(0, env_int("FAMSEQ_JIT_PREG", c.n_p_reg)));
    return c;
}

std::string source(const McmcParams &P, const GibbsJitConfig &cfg) {
    const RunConstants &C = P.C;
    const std::vector<Member> M = decode(P.plan);
    const int n = (int)M.size(), S = C.s, words = (n + 31) / 32;
    const Layout L = make_layout(n, cfg);
    const std::vector<std::vector<int>> nb = neighbours(M);
    std::ostringstream o;
    o << "// generated by famseq_b200 (gibbs_jit.cu) for one pedigree: " << n << " members, " << S << " input columns\n";
    o << "// layout: " << cfg.tb << " chains per block; draw thresholds of " << L.n_reg << " members in registers, of " << L.smem_pairs
      << " in shared memory; weights, accumulators and run bookkeeping in a block-private L2 scratch\n";
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

    o << "\nstruct Rc { u32 t0, t2, worst; };\n";
    for (int x = 0; x < 2; x++)
        for (int i = 0; i < n; i++) emit_evaluation_function(o, M, i, x == 1);
    o << "\nextern \"C\" __global__ void __launch_bounds__(TB, " << cfg.blocks << ")\n"
      << "famseq_gibbs(const double *__restrict__ lk, const u8 *__restrict__ flags, double *__restrict__ post,\n"
      << "             double *__restrict__ single, u8 *__restrict__ gt, u8 *__restrict__ status, i64 V, int burn, int rep,\n"
      << "             u64 seed, i64 v_offset, double *scratch, int n_tiles, u64 *vote_stats) {\n"
      << "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
      << "    double *s_tab = (double *)smem_raw;              // [81][" << kCopies << "]\n"
      << "    u32 *s_thr = (u32 *)(s_tab + 81 * " << kCopies << ");    // [" << L.smem_pairs << " * 2][TB] thread-private columns: T0, T2\n"
      << "    const int tid = threadIdx.x, lane = tid & 31;\n"
      << "    for (int e = tid; e < 81 * " << kCopies << "; e += TB) s_tab[e] = __longlong_as_double((i64)TAB_BITS[e / " << kCopies << "]);\n"
      << "    __syncthreads();\n"
      << "    u32 tab_addr = (u32)__cvta_generic_to_shared(s_tab) + (u32)(lane & " << (kCopies - 1) << ") * 8u;\n"
      << "    asm volatile(\"\" : \"+r\"(tab_addr) :: \"memory\"); // table reads stay below the barrier\n"
      << "    u32 *st = s_thr + tid;\n"
      << "    double *wg = scratch + (size_t)blockIdx.x * " << n * kScratchRows << " * TB + tid; // [member][A0 A1 A2 P0 P1 P2 LAST][TB], block-private\n"
      << "    const u32 k0 = (u32)seed, k1 = (u32)(seed >> 32);\n"
      << "    const double lrc = __longlong_as_double((i64)" << bits(C.lrc) << ");\n"
      << "    (void)st; (void)wg;\n\n"
      << "    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {\n"
      << "        const i64 v = (i64)tile * TB + tid;\n"
      << "        __syncwarp(); // every lane of every warp runs the same tiles: start each one converged\n"
      << "        const u32 in_tile = __ballot_sync(0xffffffffu, v < V);\n"
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
      << "        } else if (!pedigree_needed) { // family.cpp:1973-2058: FPP := individual-only posterior\n"
      << "            for (int c = 0; c < NCOL; c++) {\n"
      << "                const double p0 = gs[c * 3], p1 = gs[c * 3 + 1], p2 = gs[c * 3 + 2];\n"
      << "                gp[c * 3] = p0; gp[c * 3 + 1] = p1; gp[c * 3 + 2] = p2;\n"
      << "                gg[c] = call_genotype(p0, p1, p2);\n"
      << "            }\n"
      << "            status[v] = 0;\n"
      << "        }\n"
      << "        // the lanes that go on to sample (the mask of the per-group votes in the sweeps)\n"
      << "        const bool sampling = !failed && pedigree_needed;\n"
      << "        const u32 sampling_lanes = __ballot_sync(in_tile, sampling);\n"
      << "        if (!sampling) continue;\n"
      << "        // autosomal and chrX variants of a warp run different loops: each loop votes among its own lanes\n"
      << "        const u32 autosomal_lanes = __ballot_sync(sampling_lanes, !chrx);\n"
      << "        const u32 live = chrx ? (sampling_lanes & ~autosomal_lanes) : autosomal_lanes;\n"
      << "        const u32 tA = tab_addr, tXF = tab_addr + " << 27 * kRow << "u, tXM = tab_addr + " << 54 * kRow << "u;\n"
      << "        (void)tA; (void)tXF; (void)tXM;\n"
      << "        u32 worst = 0u;\n\n"
      << "        // chain state: nothing evaluated yet (every member dirty), empty runs\n";
    for (int i = 0; i < n; i++) {
        if (L.pair[i] < 0)
            o << "        u32 T" << i << " = 0u, U" << i << " = 0u;\n";
        else
            o << "        " << t0_ref(L, i) << " = 0u; " << t2_ref(L, i) << " = 0u;\n";
        for (int k = 0; k < kScratchRows; k++) o << "        *(" << scratch_ref(i, k) << ") = 0.0;\n";
    }
    for (int w = 0; w < words; w++) {
        const int members = std::min(32, n - 32 * w);
        o << "        u32 D" << w << " = " << (members >= 32 ? 0xffffffffu : ((1u << members) - 1u)) << "u;\n";
    }
    o << "        const u64 gv = (u64)(v_offset + v);\n"
      << "        const u32 gv_lo = (u32)gv, gv_hi = (u32)(gv >> 32);\n"
      << "        u32 r0 = 0, r1 = 0, r2 = 0, r3 = 0;\n"
      << "        // initial genotypes (family.cpp:2063-2067), stored as table-row byte offsets g * " << kRow << "\n";
    for (int i = 0; i < n; i++) {
        if ((i & 3) == 0) o << "        philox(0u, " << (i >> 2) << "u, gv_lo, gv_hi, k0, k1, r0, r1, r2, r3);\n";
        o << "        u32 o" << i << " = (r" << (i & 3) << " % 3u) * " << kRow << "u;\n";
    }
    o << "\n        const int last = burn + rep;\n"
      << "        const double first = (double)(burn + 1); // first sampling sweep: runs are counted from here (family.cpp:2069-2080)\n"
      << "        double tD = 1.0;                          // the sweep index as a double\n"
      << "        u32 n_serial = 0u;                        // groups of the sampling sweeps this warp redid member by member\n"
      << "        if (!chrx) {\n"
      << "        for (int sweep = 1; sweep <= last; sweep++, tD = __dadd_rn(tD, 1.0)) {\n";
    emit_sweep(o, M, L, nb, false);
    o << "        }\n"
      << "        } else { // the same loop with the chrX rules\n"
      << "        for (int sweep = 1; sweep <= last; sweep++, tD = __dadd_rn(tD, 1.0)) {\n";
    emit_sweep(o, M, L, nb, true);
    o << "        }\n"
      << "        }\n"
      << "        if (vote_stats && (live & (live - 1u)) == (live ^ (1u << lane))) { // the lowest live lane reports for the warp\n"
      << "            atomicAdd(vote_stats, (u64)n_serial);\n"
      << "            atomicAdd(vote_stats + 1, (u64)rep * " << (n + kGroup - 1) / kGroup << "ull);\n"
      << "        }\n"
      << "        if (worst >= 0x78600000u) { status[v] = 2; continue; } // a sum left the fast range: redo with the table-driven kernel\n\n"
      << "        // postProb = genoFry / numRep, not renormalised; a row summing to <= 0 fails (family.cpp:2082-2092).\n"
      << "        // tD is last + 1 here: the weights still standing close their runs.\n"
      << "        const double nrep = (double)rep;\n";
    for (int i = 0; i < n; i++) {
        o << "        {\n"
          << "            const double run = fmax(0.0, __dsub_rn(tD, fmax(__ldcg(" << scratch_ref(i, ROW_LAST) << "), first)));\n"
          << "            const double p0 = __ddiv_rn(__dadd_rn(__ldcg(" << scratch_ref(i, ROW_A0) << "), __dmul_rn(run, __ldcg(" << scratch_ref(i, ROW_P0)
          << "))), nrep);\n"
          << "            const double p1 = __ddiv_rn(__dadd_rn(__ldcg(" << scratch_ref(i, ROW_A0 + 1) << "), __dmul_rn(run, __ldcg(" << scratch_ref(i, ROW_P1)
          << "))), nrep);\n"
          << "            const double p2 = __ddiv_rn(__dadd_rn(__ldcg(" << scratch_ref(i, ROW_A0 + 2) << "), __dmul_rn(run, __ldcg(" << scratch_ref(i, ROW_P2)
          << "))), nrep);\n";
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

} // namespace cached

// ================================================================================================================
// Generator 2: dense sweeps -- every full conditional evaluated in every step (round 1's kernel).  Its speed does not depend
// on the data, so it is the choice where the chains keep moving: pedigrees with many unsequenced members.  A sweep is ONE
// basic block of ~50 instructions per member; per chain, 3 own factors (1e6 * prior * lk) and 3 Rao-Blackwell accumulators
// per member are placed in registers, in thread-private shared-memory columns, or in a block-private scratch that stays in
// L2 (own factors through a software prefetch queue, accumulators through red.global.add.f64).  Same chains as the other
// two kernels; accumulates sweep by sweep like the table-driven kernel (same bytes as that one).
// ================================================================================================================
namespace dense {

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
    // ((r >> 1) + 0.5) * 2^-31 = (r | 1) * 2^-32 without an int-to-double conversion: 1 + (r | 1) * 2^-32 assembled from bits, minus 1 (exact)
    o << "                const double rd = __dsub_rn(__hiloint2double((int)(0x3ff00000u | ((r" << (i & 3) << " | 1u) >> 12)), (int)((r" << (i & 3)
      << " | 1u) << 20)), 1.0);\n";
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

GibbsJitConfig default_config(const McmcParams &P) {
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
    c.cached = 0;
    c.tb = env_int("FAMSEQ_JIT_TB", c.tb);
    c.blocks = std::max(1, env_int("FAMSEQ_JIT_BLOCKS", c.blocks));
    c.prefetch = std::max(1, env_int("FAMSEQ_JIT_PF", c.prefetch));
    c.n_acc_reg = std::min(n, std::max(0, env_int("FAMSEQ_JIT_RACC", c.n_acc_reg)));
    c.n_acc_smem = std::min(n - c.n_acc_reg, std::max(0, env_int("FAMSEQ_JIT_SACC", c.n_acc_smem)));
    c.n_lk_reg = std::min(n, std::max(0, env_int("FAMSEQ_JIT_RLK", c.n_lk_reg)));
    c.n_lk_smem = std::min(n - c.n_lk_reg, std::max(0, env_int("FAMSEQ_JIT_SLK", c.n_lk_smem)));
    return c;
}

std::string source(const McmcParams &P, const GibbsJitConfig &cfg) {
    const RunConstants &C = P.C;
    const std::vector<Member> M = decode(P.plan);
    const int n = (int)M.size(), S = C.s;
    const Layout L = make_layout(n, cfg);
    std::ostringstream o;
    o << "// generated by famseq_b200 (gibbs_jit.cu, dense sweeps) for one pedigree: " << n << " members, " << S << " input columns\n";
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
      << "             u64 seed, i64 v_offset, double *scratch, int n_tiles, u64 *vote_stats) {\n"
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
      << "    (void)sa; (void)wg; (void)vote_stats;\n\n"
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

} // namespace dense

} // namespace

// Which generator: FAMSEQ_JIT_CACHED=0/1 decides; otherwise cached conditionals when at least three quarters of the
// members are sequenced (unsequenced members have flat likelihoods: their genotypes, and with them their neighbours'
// conditionals, change all the time, and a cache that is always stale only costs).
GibbsJitConfig gibbs_jit_default_config(const McmcParams &P) {
    int sequenced = 0;
    for (int i = 0; i < P.plan.n; i++) sequenced += P.plan.col[i] >= 0;
    const int want = env_int("FAMSEQ_JIT_CACHED", 4 * sequenced >= 3 * P.plan.n ? 1 : 0);
    return want ? cached::default_config(P) : dense::default_config(P);
}

GibbsJitConfig gibbs_jit_config(const McmcParams &P, int cached) { return cached ? cached::default_config(P) : dense::default_config(P); }

bool gibbs_jit_generator_forced() {
    const char *v = std::getenv("FAMSEQ_JIT_CACHED");
    return v && *v;
}

std::string gibbs_jit_source(const McmcParams &P, const GibbsJitConfig &cfg) { return cfg.cached ? cached::source(P, cfg) : dense::source(P, cfg); }

namespace {
// shared memory and block-private scratch (doubles per block) of a layout
size_t layout_smem(const McmcParams &P, const GibbsJitConfig &cfg) {
    return cfg.cached ? cached::smem_bytes(cached::make_layout(P.plan.n, cfg), cfg.tb) : dense::smem_bytes(dense::make_layout(P.plan.n, cfg), cfg.tb);
}
size_t layout_scratch_doubles(const McmcParams &P, const GibbsJitConfig &cfg) {
    if (cfg.cached) return (size_t)P.plan.n * cached::kScratchRows * cfg.tb;
    return (size_t)std::max(1, dense::make_layout(P.plan.n, cfg).glob_rows) * 3 * cfg.tb;
}
} // namespace

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
    size_t scratch_doubles = 0; // block-private scratch, doubles per block
    int blocks_per_sm = 1;
};

int gibbs_jit_load(const McmcParams &P, const GibbsJitConfig &cfg, const std::string &cubin, GibbsJitKernel **out, std::string &err) {
    *out = nullptr;
    const size_t smem = layout_smem(P, cfg);
    GibbsJitKernel *k = new GibbsJitKernel();
    k->cfg = cfg;
    k->smem = smem;
    k->scratch_doubles = layout_scratch_doubles(P, cfg);
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
    const size_t smem = layout_smem(P, cfg);
    if (cfg.tb < 32 || cfg.tb > 1024 || cfg.tb % 32 || smem > kSmemPerBlockMax) {
        err = "Gibbs JIT: invalid layout (tb " + std::to_string(cfg.tb) + ", shared memory " + std::to_string(smem) + " B)";
        return FS_E_TOO_LARGE;
    }
    const int rc = gibbs_jit_compile(gibbs_jit_source(P, cfg), cubin, log, err);
    if (rc != FS_OK) return rc;
    if (const char *v = std::getenv("FAMSEQ_JIT_VERBOSE"))
        if (v[0] == '1')
            std::fprintf(stderr, "[famseq] Gibbs JIT (%s) tb=%d blocks=%d smem=%zu\n%s\n", cfg.cached ? "cached conditionals" : "dense sweeps", cfg.tb,
                         cfg.blocks, smem, log.c_str());
    return FS_OK;
}

void gibbs_jit_unload(GibbsJitKernel *k) {
    if (!k) return;
    if (k->library) cudaLibraryUnload(k->library);
    delete k;
}

cudaError_t gibbs_jit_launch(GibbsJitKernel *k, const BatchPtrs &B, int burn, int rep, uint64_t seed, int64_t v_offset,
                             int sm_count, cudaStream_t stream, unsigned long long *vote_stats) {
    if (B.V <= 0) return cudaSuccess;
    const int tb = k->cfg.tb;
    const int64_t n_tiles64 = (B.V + tb - 1) / tb;
    if (n_tiles64 > 0x7fffffff) return cudaErrorInvalidValue;
    int n_tiles = (int)n_tiles64;
    const int grid = (int)std::min<int64_t>(n_tiles64, (int64_t)sm_count * k->blocks_per_sm);
    double *scratch = nullptr; // accumulators and run bookkeeping [grid][member][6][tb], stream-ordered
    cudaError_t rc = cudaMallocAsync(&scratch, (size_t)grid * std::max<size_t>(1, k->scratch_doubles) * sizeof(double), stream);
    if (rc != cudaSuccess) return rc;
    const double *lk = B.lk;
    const uint8_t *flags = B.flags;
    double *post = B.post, *single = B.single;
    uint8_t *gt = B.gt, *status = B.status;
    long long V = B.V, voff = v_offset;
    unsigned long long sd = seed;
    void *args[] = {&lk, &flags, &post, &single, &gt, &status, &V, &burn, &rep, &sd, &voff, &scratch, &n_tiles, &vote_stats};
    rc = cudaLaunchKernel((const void *)k->kernel, dim3(grid), dim3(tb), args, k->smem, stream);
    const cudaError_t rc2 = cudaFreeAsync(scratch, stream);
    return rc != cudaSuccess ? rc : rc2;
}

} // namespace famseq
