"""CPU-only check of the TILE PIPELINE of the nuclear-family kernel for compact input
(famseq_b200/csrc/cuda/es_nuclear_kernel.cuh: es_nuclear_stream_kernel, and the one-tile es_nuclear_kernel beside it).

The kernel source itself -- not a copy -- is compiled for the host: the block of PTX wrappers (mbarrier, cp.async.bulk) and the
launch functions are cut out of the text, a stand-in <cuda_runtime.h> supplies the two CUDA types the headers mention, and a
shim plays the rest: a warp is 32 host threads (`__syncwarp` = a barrier), the device intrinsics are plain double operations
(-ffp-contract=off), and the copy engine is an adversary as in test_es_jit_pipeline_cpu.py -- a bulk load lands at the earliest
legal moment and a bulk store reads shared memory at the latest one (inside the wait_group that follows it), or the other way
round.  The one-tile kernel runs on FP64 and on compact input; blocks of the tile-list kernel walk 1, 3, 4 and 11 consecutive tiles (double-buffered compact tiles, flags on the tile's transaction, output
rows reused from tile to tile), the last tile is ragged; the results must be the oracle's bytes on the decoded likelihoods,
with and without `single`, without flags, for trios and quads.  Where the tool chain has ThreadSanitizer the program also runs
under it.  Test infrastructure: the product has no CPU compute path."""
import os
import subprocess

import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "famseq_b200", "csrc")

FAKE_CUDA_RUNTIME = r"""
#pragma once
// stand-in for <cuda_runtime.h> in the host build of the kernel source (tests/test_nuclear_stream_cpu.py)
#include <cstdint>
#include <cstddef>
typedef int cudaError_t;
typedef void *cudaStream_t;
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n)
#define __restrict__
#define __shared__ static
"""

SHIM = r"""
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>
#include <cuda_runtime.h>
using std::min;
using std::max;
struct Dim { unsigned x; };
static thread_local Dim threadIdx = {0};
static Dim blockIdx = {0}, gridDim = {1};
alignas(128) static unsigned char smem_raw[232448];

// ---- a warp of 32 host threads ------------------------------------------------------------------------------------
static std::barrier<> g_warp(32);
static inline void __syncwarp(unsigned = 0xffffffffu) { g_warp.arrive_and_wait(); }
static inline void __syncthreads() { std::fprintf(stderr, "shim: one warp per block only\n"); std::_Exit(4); }

// ---- the copy engine as an adversary ------------------------------------------------------------------------------
static int g_late_loads = 0;  // 0: a bulk load lands when it is issued; 1: inside the mbarrier wait
static int g_late_stores = 1; // 1: a bulk store reads shared memory inside the wait_group that follows it; 0: at once
static int g_errors = 0;
struct Copy { void *dst; const void *src; unsigned bytes; };
static std::mutex g_engine;
static std::vector<Copy> g_loads, g_open_stores, g_committed_stores;
static std::atomic<unsigned> g_phases_done{0};
static unsigned g_expected_tx = 0;
static inline void run_copies(std::vector<Copy> &q) {
    for (const Copy &c : q) std::memcpy(c.dst, c.src, c.bytes);
    q.clear();
}
namespace famseq { namespace {
inline void mbar_init(uint64_t *, unsigned) {
    g_phases_done.store(0);
    g_expected_tx = 0;
}
inline void mbar_expect_tx(uint64_t *, unsigned bytes) {
    std::lock_guard<std::mutex> lock(g_engine);
    if (g_expected_tx != 0) { std::fprintf(stderr, "shim: two transactions in flight on one barrier phase\n"); g_errors++; }
    g_expected_tx = bytes;
}
inline void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *) {
    std::lock_guard<std::mutex> lock(g_engine);
    if (bytes % 16 || ((size_t)smem_dst & 15) || ((size_t)gmem_src & 15)) { std::fprintf(stderr, "shim: misaligned bulk load\n"); g_errors++; }
    g_loads.push_back({smem_dst, gmem_src, bytes});
    if (!g_late_loads) run_copies(g_loads);
    if (bytes > g_expected_tx) { std::fprintf(stderr, "shim: more bytes than expected on the barrier\n"); g_errors++; bytes = g_expected_tx; }
    g_expected_tx -= bytes;
    if (g_expected_tx == 0 && !g_late_loads) g_phases_done.fetch_add(1);
}
inline void mbar_wait(uint64_t *, unsigned parity) {
    {
        std::lock_guard<std::mutex> lock(g_engine);
        if (g_late_loads && !g_loads.empty() && g_expected_tx == 0) { // the data arrives now, at the last moment
            run_copies(g_loads);
            g_phases_done.fetch_add(1);
        }
    }
    long spins = 0;
    while ((g_phases_done.load() & 1u) == parity) {
        std::this_thread::yield();
        if (++spins > 20000000L) { std::fprintf(stderr, "shim: mbarrier wait never ends\n"); std::_Exit(3); }
    }
}
inline void bulk_store(void *gmem_dst, const void *smem_src, unsigned bytes) {
    std::lock_guard<std::mutex> lock(g_engine);
    if (bytes % 16 || ((size_t)smem_src & 15) || ((size_t)gmem_dst & 15)) { std::fprintf(stderr, "shim: misaligned bulk store\n"); g_errors++; }
    g_open_stores.push_back({gmem_dst, smem_src, bytes});
    if (!g_late_stores) run_copies(g_open_stores);
}
inline void bulk_commit() {
    std::lock_guard<std::mutex> lock(g_engine);
    for (const Copy &c : g_open_stores) g_committed_stores.push_back(c);
    g_open_stores.clear();
}
inline void bulk_wait_read() {
    std::lock_guard<std::mutex> lock(g_engine);
    run_copies(g_committed_stores);
}
inline void bulk_commit_and_wait_read() { bulk_commit(); bulk_wait_read(); }
inline void fence_async_smem() {}
} }

// ---- device intrinsics ----------------------------------------------------------------------------------------------
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double d) { long long x; std::memcpy(&x, &d, 8); return (int)(x >> 32); }
static inline int __double2loint(double d) { long long x; std::memcpy(&x, &d, 8); return (int)x; }
template <class T> static inline T __ldg(const T *p) { return *p; }
"""

DRIVER = r"""
#include "../host/pedigree.hpp"
namespace famseq {
template <int NC, bool SINGLE, bool IDENT>
static int run_grid(const NuclearParams &P, const BatchPtrs &B, int list_len) {
    const int64_t n_tiles = (B.V + 31) / 32;
    const int grid = list_len > 0 ? (int)((n_tiles + list_len - 1) / list_len) : (int)n_tiles;
    gridDim.x = (unsigned)grid;
    for (int b = 0; b < grid; b++) { // blocks one after the other: they share nothing
        blockIdx.x = (unsigned)b;
        g_loads.clear(), g_open_stores.clear(), g_committed_stores.clear();
        std::vector<std::thread> lanes;
        for (int l = 0; l < 32; l++)
            lanes.emplace_back([&, l] {
                threadIdx.x = (unsigned)l;
                if (list_len > 0)
                    es_nuclear_stream_kernel<NC, SINGLE, IDENT, 0>(P, B, list_len);
                else if (B.pl) // the one-tile kernel on the same input
                    es_nuclear_kernel<NC, 32, true, SINGLE, IDENT, 0>(P, B);
                else // ... and on FP64 likelihoods
                    es_nuclear_kernel<NC, 32, false, SINGLE, IDENT, 0>(P, B);
            });
        for (auto &t : lanes) t.join();
        if (!g_open_stores.empty() || !g_committed_stores.empty()) {
            std::fprintf(stderr, "shim: block %d ended with a bulk store nobody waited for\n", b);
            g_errors++;
        }
        if (!g_loads.empty()) {
            std::fprintf(stderr, "shim: block %d ended with a bulk load in flight\n", b);
            g_errors++;
        }
    }
    return g_errors;
}
template <int NC> static int run_nc(const NuclearParams &P, const BatchPtrs &B, int list_len, bool ident) {
    if (ident) return B.single ? run_grid<NC, true, true>(P, B, list_len) : run_grid<NC, false, true>(P, B, list_len);
    return B.single ? run_grid<NC, true, false>(P, B, list_len) : run_grid<NC, false, false>(P, B, list_len);
}
}
// usage: prog n_children list_len late_loads late_stores want_single use_flags V in.bin out.bin
//   in : mrate, lrc, priors[4][3] (f64); S, role_col[7] (father, mother, children; -1 = unsequenced), male_child[5] (i32);
//        lut[65536] (f64), pl [V][S][3] (u16), flags [V]
//   out: post [V][S][3], single [V][S][3] (f64), gt [V][S], status [V]
int main(int argc, char **argv) {
    using namespace famseq;
    if (argc != 10) return 2;
    const int nc = std::atoi(argv[1]), list_len = std::atoi(argv[2]);
    g_late_loads = std::atoi(argv[3]);
    g_late_stores = std::atoi(argv[4]);
    const bool want_single = std::atoi(argv[5]) != 0, use_flags = std::atoi(argv[6]) != 0;
    const int64_t V = std::atoll(argv[7]);
    auto room = [](size_t bytes) { return std::aligned_alloc(128, (bytes + 127) / 128 * 128 + 128); };
    FILE *f = std::fopen(argv[8], "rb");
    if (!f) return 2;
    double head[2 + 12];
    int32_t ints[1 + 7 + 5];
    if (std::fread(head, 8, 14, f) != 14 || std::fread(ints, 4, 13, f) != 13) return 2;
    const int S = ints[0];
    const int32_t *role_col = ints + 1, *male_child = ints + 8;
    const size_t n3 = (size_t)V * S * 3, n1 = (size_t)V * S;
    double *lut = (double *)room(65536 * 8);
    uint16_t *pl = (uint16_t *)room(n3 * 2);
    uint8_t *flags = (uint8_t *)room((size_t)V);
    if (std::fread(lut, 8, 65536, f) != 65536 || std::fread(pl, 2, n3, f) != n3 || std::fread(flags, 1, (size_t)V, f) != (size_t)V) return 2;
    std::fclose(f);
    NuclearParams P;
    std::memset(&P, 0, sizeof P);
    build_tables(head[0], P.C.tab);
    P.C.lrc = head[1];
    std::memcpy(P.C.prior, head + 2, 96);
    P.C.n = nc + 2, P.C.s = S;
    P.n_children = nc;
    P.col_father = role_col[0], P.col_mother = role_col[1];
    if (role_col[0] >= 0) P.C.col_male[role_col[0]] = 1;
    bool ident = S == nc + 2 && role_col[0] == 0 && role_col[1] == 1;
    for (int c = 0; c < nc; c++) {
        P.col_child[c] = role_col[2 + c], P.male_child[c] = male_child[c];
        if (role_col[2 + c] >= 0) P.C.col_male[role_col[2 + c]] = (uint8_t)male_child[c];
        ident = ident && role_col[2 + c] == 2 + c;
    }
    // unseq_fail (engine.cu: fs_create): an unsequenced member whose prior row sums to <= 0 fails every such variant
    for (int fl = 0; fl < 4; fl++) {
        const int known = fl & 1, chrx = (fl >> 1) & 1;
        bool bad = false;
        for (int r = 0; r < nc + 2; r++) {
            if (role_col[r] >= 0) continue;
            const bool male = r == 0 || (r >= 2 && male_child[r - 2]);
            const double *pr = (chrx && male) ? P.C.prior[known ? 3 : 2] : P.C.prior[known ? 1 : 0];
            if ((pr[0] + pr[1]) + pr[2] <= 0) bad = true;
        }
        P.C.unseq_fail[fl] = bad;
    }
    P.allow_ident = 1;
    double *post = (double *)room(n3 * 8), *single = (double *)room(n3 * 8);
    uint8_t *gt = (uint8_t *)room(n1), *status = (uint8_t *)room((size_t)V);
    for (size_t k = 0; k < n3; k++) post[k] = single[k] = -1.0;
    std::memset(gt, 77, n1);
    std::memset(status, 77, (size_t)V);
    BatchPtrs B;
    B.lk = nullptr, B.flags = use_flags ? flags : nullptr, B.post = post, B.single = want_single ? single : nullptr, B.gt = gt, B.status = status, B.V = V;
    B.pl = pl, B.lut = lut;
    if (list_len < 0) { // FP64 input: the decoded likelihoods
        double *lk = (double *)room(n3 * 8);
        for (size_t k = 0; k < n3; k++) lk[k] = lut[pl[k]];
        B.lk = lk, B.pl = nullptr, B.lut = nullptr;
    }
    ident = ident && (B.pl || nc > 1); // the product's rule (launch_io): the FP64 trio keeps the general code
    int errors = 0;
    if (nc == 1) errors = run_nc<1>(P, B, list_len, ident);
    else if (nc == 2) errors = run_nc<2>(P, B, list_len, ident);
    else if (nc == 3) errors = run_nc<3>(P, B, list_len, ident);
    else return 2;
    f = std::fopen(argv[9], "wb");
    if (!f) return 2;
    std::fwrite(post, 8, n3, f), std::fwrite(single, 8, n3, f), std::fwrite(gt, 1, n1, f), std::fwrite(status, 1, (size_t)V, f);
    std::fclose(f);
    return errors ? 1 : 0;
}
"""


def host_source() -> str:
    """es_nuclear_kernel.cuh without its PTX wrappers (the shim's copy engine takes their place) and without the launch functions."""
    text = open(os.path.join(CSRC, "cuda", "es_nuclear_kernel.cuh")).read()
    a = text.index("// ---- TMA bulk copy helpers")
    b = text.index("template <bool X> __device__ __forceinline__ double trans(")
    text = text[:a] + text[b:]
    cut = text.index("template <int NC, bool SINGLE, bool IDENT, int MINB>\ncudaError_t launch_stream(")
    text = text[:cut] + "} // namespace\n} // namespace famseq\n"
    # launch_minb sits between the two kernels' definitions in the file: it goes as well
    a = text.index("template <int NC, int TB, bool PL, bool SINGLE, bool IDENT, int MINB>\ncudaError_t launch_minb(") if "cudaError_t launch_minb(" in text else -1
    if a >= 0:
        b = text.index("// The same computation for compact input as a tile pipeline")
        text = text[:a] + text[b:]
    assert "<<<" not in text and "asm volatile" not in text
    return text.replace("extern __shared__ __align__(128) unsigned char smem_raw[];", "").replace("#pragma once", "")


def build(tmp_path, source=None, with_tsan=True):
    inc = tmp_path / "inc"
    inc.mkdir(exist_ok=True)
    (inc / "cuda_runtime.h").write_text(FAKE_CUDA_RUNTIME)
    cpp = str(tmp_path / "nuclear_host.cpp")
    open(cpp, "w").write(SHIM + (source if source is not None else host_source()) + DRIVER)
    flags = ["-O1", "-std=c++20", "-ffp-contract=off", "-pthread", "-w", f"-I{inc}", f"-I{os.path.join(CSRC, 'cuda')}"]
    extra = [os.path.join(CSRC, "host", "pedigree.cpp")]
    plain, tsan = str(tmp_path / "nuclear_host"), str(tmp_path / "nuclear_host_tsan")
    jobs = [subprocess.Popen(["g++"] + flags + ["-o", plain, cpp] + extra)]
    if with_tsan:
        jobs.append(subprocess.Popen(["g++", "-fsanitize=thread", "-g"] + flags + ["-o", tsan, cpp] + extra, stderr=subprocess.DEVNULL))
    rcs = [j.wait() for j in jobs]
    assert rcs[0] == 0, "the kernel source does not compile for the host"
    return plain, (tsan if with_tsan and rcs[1] == 0 else None)


@pytest.fixture(scope="module")
def programs(tmp_path_factory):
    """The host programs of the unmodified kernel source, built once: (plain, under ThreadSanitizer or None)."""
    return build(tmp_path_factory.mktemp("nuclear_host"))


def roles_of(ped):
    """Ped row of the father, the mother and the children (ped order) of a nuclear family."""
    ids = list(ped.ids)
    kids = [i for i in range(ped.n) if ped.mids[i] != 0]
    return [ids.index(ped.fids[kids[0]]), ids.index(ped.mids[kids[0]])] + kids


def run(tmp_path, prog, ped, pl, fl, list_len, late_loads, late_stores, want_single=True, use_flags=True, cols=None):
    cols = list(ped.sequenced_cols() if cols is None else cols)
    V, S = pl.shape[0], pl.shape[1]
    assert S == len(cols)
    roles = roles_of(ped)
    nc = len(roles) - 2
    prm = fs.Params.default()
    ints = np.full(13, -1, np.int32)
    ints[0] = S
    for r, row in enumerate(roles):
        ints[1 + r] = cols.index(row) if row in cols else -1
    ints[8:13] = 0
    ints[8:8 + nc] = [1 if ped.genders[row] == 1 else 0 for row in roles[2:]]
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([prm.mrate, prm.lrc], np.float64).tobytes())
        f.write(np.ascontiguousarray(prm.priors(), np.float64).tobytes())
        f.write(ints.tobytes())
        f.write(np.ascontiguousarray(O.pl_table(), np.float64).tobytes())
        f.write(np.ascontiguousarray(pl, np.uint16).tobytes())
        f.write(np.ascontiguousarray(fl, np.uint8).tobytes())
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    r = subprocess.run([prog, str(nc), str(list_len), str(late_loads), str(late_stores), str(int(want_single)), str(int(use_flags)), str(V), inp, out],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, f"host kernel exited {r.returncode}: {r.stderr[-2000:]}"
    raw = np.fromfile(out, np.uint8)
    n3 = V * S * 3 * 8
    return (raw[:n3].view(np.float64).reshape(V, S, 3), raw[n3:2 * n3].view(np.float64).reshape(V, S, 3),
            raw[2 * n3:2 * n3 + V * S].reshape(V, S), raw[2 * n3 + V * S:])


def check(got, want, want_single, what):
    post, single, gt, status = got
    assert np.array_equal(status, want["status"].astype(np.uint8)), what
    ok = want["status"] == 0
    assert np.array_equal(post[ok], want["post"][ok]), what
    assert np.array_equal(gt[ok], want["gt"][ok].astype(np.uint8)), what
    assert not post[~ok].any() and (gt[~ok] == 255).all(), what
    if want_single:
        assert np.array_equal(single[ok], want["single"][ok]) and not single[~ok].any(), what
    else:
        assert (single == -1.0).all(), what  # never touched


def batch(ped, V, seed):
    """PL integers with the awkward values mixed in (subnormal and zero likelihoods, missing samples, impossible trios)."""
    pl, fl = synth.synth_pl(ped, V, seed, x_fraction=0.3)
    pl = pl.astype(np.uint16)
    rng = np.random.default_rng(seed)
    odd = rng.random(pl.shape) < 0.02
    pl[odd] = rng.choice(np.array([2551, 3076, 3077, 3100, 3236, 3237, 5000, 65535], np.uint16), int(odd.sum()))
    pl[rng.random(V) < 0.02] = 0
    pl[rng.random(V) < 0.02] = 65535  # every likelihood exactly zero: the variant fails
    return pl, fl


@pytest.mark.parametrize("n_children", [1, 2])
def test_tile_lists_of_the_compact_nuclear_kernel_against_an_adversarial_copy_engine(n_children, tmp_path, programs):
    rows = [(1, 0, 0, 1), (2, 0, 0, 2)] + [(3 + k, 2, 1, 1 + k % 2) for k in range(n_children)]
    ped = synth._mk(rows)
    V = 32 * 10 + 7  # eleven tiles, the last one ragged
    pl, fl = batch(ped, V, seed=31 + n_children)
    want = O.run(ped, ped.sequenced_cols(), O.pl_table()[pl], fl, method=O.ES)
    assert 0 < (want["status"] != 0).sum() < V
    want_noflags = O.run(ped, ped.sequenced_cols(), O.pl_table()[pl], np.zeros_like(fl), method=O.ES)
    plain, tsan = programs
    for list_len in (-1, 0, 1, 3, 4, 64):  # -1 / 0: the one-tile kernel on FP64 / compact input; 64: one block walks everything
        for late_loads, late_stores in ((0, 1), (1, 0)):
            for want_single in (True, False):
                got = run(tmp_path, plain, ped, pl, fl, list_len, late_loads, late_stores, want_single)
                check(got, want, want_single, f"C={n_children} list={list_len} late_loads={late_loads} late_stores={late_stores} single={want_single}")
    got = run(tmp_path, plain, ped, pl, fl, 3, 0, 1, True, use_flags=False)
    check(got, want_noflags, True, "no flags array")
    if tsan is not None:
        got = run(tmp_path, tsan, ped, pl, fl, 4, 0, 1, True)
        check(got, want, True, "under ThreadSanitizer")


@pytest.mark.parametrize("rows,cols", [
    ([(7, 5, 9, 2), (9, 0, 0, 1), (4, 5, 9, 1), (5, 0, 0, 2)], [3, 0, 2]),          # children before parents, the father unsequenced
    ([(1, 0, 0, 1), (2, 0, 0, 2), (3, 2, 1, 1), (4, 2, 1, 2), (5, 2, 1, 1)], None),   # three children, identity column map
    ([(1, 0, 0, 1), (2, 0, 0, 2), (3, 2, 1, 2), (4, 2, 1, 1), (5, 2, 1, 2)], [4, 1, 0, 2]),  # three children, one unsequenced, columns permuted
])
def test_column_maps_of_the_nuclear_kernel_on_the_host(rows, cols, tmp_path, programs):
    """The arithmetic of the nuclear kernel's source (fast pass, complete pass, both column-map code paths) against the oracle
    on the host: unsequenced members, permuted columns, three children; FP64 and compact input, one tile and lists of tiles."""
    ped = synth._mk(rows)
    cols = ped.sequenced_cols() if cols is None else cols
    V = 32 * 5 + 9
    pl, fl = batch(synth._mk([(i, 0, 0, 1) for i in range(1, len(cols) + 1)]), V, seed=17 + len(cols))
    want = O.run(ped, cols, O.pl_table()[pl], fl, method=O.ES)
    assert 0 < (want["status"] != 0).sum() < V
    plain, _ = programs
    for list_len in (-1, 0, 3):
        for want_single in (True, False):
            got = run(tmp_path, plain, ped, pl, fl, list_len, 0, 1, want_single, cols=cols)
            check(got, want, want_single, f"rows={rows} cols={cols} list={list_len} single={want_single}")


def test_the_adversary_notices_a_missing_wait(tmp_path):
    """The shim must be able to fail: without the wait for the previous tile's stores, its rows are overwritten before the copy
    engine (reading as late as it may) has taken them."""
    ped = synth.trio()
    V = 32 * 8
    pl, fl = batch(ped, V, seed=5)
    want = O.run(ped, ped.sequenced_cols(), O.pl_table()[pl], fl, method=O.ES)
    needle = "        if (lane == 0) bulk_wait_read(); // the previous tile's rows have left shared memory\n"
    text = host_source()
    assert text.count(needle) == 1
    plain, _ = build(tmp_path, text.replace(needle, ""), with_tsan=False)
    post, single, gt, status = run(tmp_path, plain, ped, pl, fl, 4, 0, 1, True)
    ok = want["status"] == 0
    assert not np.array_equal(post[ok], want["post"][ok])
