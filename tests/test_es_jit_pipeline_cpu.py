"""CPU-only check of the TILE PIPELINE of the generated Elston-Stewart kernel (famseq_b200/csrc/cuda/es_jit.cu).

test_es_jit_cpu.py runs the generated arithmetic (peel_a / peel_x) for one variant at a time.  This file runs the WHOLE
generated kernel -- tile loop, double use of the two shared-memory buffers, pipeline point, warp votes, late clean-up --
compiled for the host behind a shim in which a warp is 32 host threads and the copy engine is an adversary:

  * `__syncwarp` is a barrier of the 32 threads, `__all_sync` / `__any_sync` are reductions over them;
  * a TMA bulk LOAD lands at the EARLIEST legal moment (at once, when it is issued): a buffer handed to the next tile while
    the current tile still needs it shows up as wrong results;
  * a TMA bulk STORE reads shared memory at the LATEST legal moment (inside the `cp.async.bulk.wait_group[.read]` that
    follows it): rows overwritten before the kernel waited for them show up as wrong results; a block that ends with a
    store it never waited for is an error;
  * in a second mode both are turned round (loads land inside the mbarrier wait, stores read at once).

The results must be the oracle's bytes for every variant, on blocks that walk several tiles each, with a ragged last tile,
chrX / Known variants, LRC-gated and failing variants, and variants whose row sum vanishes inside the pedigree (-mRate 0)
after their `single` rows have left.  Where the tool chain has ThreadSanitizer the same program also runs under it: a
missing __syncwarp between a cooperative write and a read, or between a lane's rows and the copy engine, is a data race the
sanitizer reports.  Test infrastructure: the product has no CPU compute path."""
import os
import subprocess

import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O

SHIM = r"""
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
typedef unsigned int u32; typedef unsigned long long u64; typedef long long i64; typedef unsigned char u8;
#define __device__
#define __global__
#define __constant__ static const
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n)
#define __restrict__
#define __shared__ static
struct Dim { unsigned x; };
static thread_local Dim threadIdx = {0};
static Dim blockIdx = {0}, gridDim = {1};
alignas(128) static unsigned char smem_raw[232448];

// ---- a warp of 32 host threads --------------------------------------------------------------------------------
static std::barrier<> g_warp(32);
static bool g_vote[32];
static inline void __syncwarp(unsigned = 0xffffffffu) { g_warp.arrive_and_wait(); }
static inline bool __all_sync(unsigned, bool p) {
    g_vote[threadIdx.x] = p;
    g_warp.arrive_and_wait();
    bool all = true;
    for (int l = 0; l < 32; l++) all = all && g_vote[l];
    g_warp.arrive_and_wait(); // nobody overwrites its vote before everybody has read them
    return all;
}
static inline bool __any_sync(unsigned, bool p) { return !__all_sync(0xffffffffu, !p); }

// ---- the copy engine as an adversary ----------------------------------------------------------------------------
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
static inline void mbar_init(u64 *, unsigned) {
    g_phases_done.store(0);
    g_expected_tx = 0;
}
static inline void mbar_expect_tx(u64 *, unsigned bytes) {
    std::lock_guard<std::mutex> lock(g_engine);
    if (g_expected_tx != 0) { std::fprintf(stderr, "shim: two copies in flight on one barrier phase\n"); g_errors++; }
    g_expected_tx = bytes;
}
static inline void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, u64 *) {
    std::lock_guard<std::mutex> lock(g_engine);
    if (bytes % 16 || ((size_t)smem_dst & 15) || ((size_t)gmem_src & 15)) { std::fprintf(stderr, "shim: misaligned bulk load\n"); g_errors++; }
    g_loads.push_back({smem_dst, gmem_src, bytes});
    if (!g_late_loads) run_copies(g_loads);
    if (bytes > g_expected_tx) { std::fprintf(stderr, "shim: more bytes than expected on the barrier\n"); g_errors++; bytes = g_expected_tx; }
    g_expected_tx -= bytes;
    if (g_expected_tx == 0 && !g_late_loads) g_phases_done.fetch_add(1);
}
static inline void mbar_wait(u64 *, unsigned parity) {
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
static inline void bulk_store(void *gmem_dst, const void *smem_src, unsigned bytes) {
    std::lock_guard<std::mutex> lock(g_engine);
    if (bytes % 16 || ((size_t)smem_src & 15) || ((size_t)gmem_dst & 15)) { std::fprintf(stderr, "shim: misaligned bulk store\n"); g_errors++; }
    g_open_stores.push_back({gmem_dst, smem_src, bytes});
    if (!g_late_stores) run_copies(g_open_stores);
}
static inline void bulk_commit() {
    std::lock_guard<std::mutex> lock(g_engine);
    for (const Copy &c : g_open_stores) g_committed_stores.push_back(c);
    g_open_stores.clear();
}
static inline void bulk_wait_read() {
    std::lock_guard<std::mutex> lock(g_engine);
    run_copies(g_committed_stores);
}
static inline void bulk_wait_all() { bulk_wait_read(); }
static inline void fence_async_smem() {}

// ---- arithmetic ---------------------------------------------------------------------------------------------------
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, 8); return d; }
static inline int __double2hiint(double d) { long long x; std::memcpy(&x, &d, 8); return (int)(x >> 32); }
static inline int __double2loint(double d) { long long x; std::memcpy(&x, &d, 8); return (int)x; }
"""

DRIVER = r"""
static int run_grid(int grid, const double *lk, const u8 *flags, double *post, double *single, u8 *gt, u8 *status, i64 V) {
    gridDim.x = (unsigned)grid;
    for (int b = 0; b < grid; b++) { // blocks one after the other: they share nothing
        blockIdx.x = (unsigned)b;
        g_loads.clear(), g_open_stores.clear(), g_committed_stores.clear();
        std::vector<std::thread> lanes;
        for (int l = 0; l < 32; l++)
            lanes.emplace_back([=] {
                threadIdx.x = (unsigned)l;
                famseq_es(lk, flags, post, single, gt, status, V);
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
// usage: prog grid late_loads late_stores V in.bin out.bin      in: lk [V][NCOL][3] f64, flags [V]; out: post, single, gt, status
int main(int argc, char **argv) {
    if (argc != 7) return 2;
    const int grid = std::atoi(argv[1]);
    g_late_loads = std::atoi(argv[2]);
    g_late_stores = std::atoi(argv[3]);
    const i64 V = std::atoll(argv[4]);
    const size_t n3 = (size_t)V * S3, n1 = (size_t)V * NCOL;
    auto room = [](size_t bytes) { return std::aligned_alloc(128, (bytes + 127) / 128 * 128 + 128); };
    double *lk = (double *)room(n3 * 8), *post = (double *)room(n3 * 8), *single = (double *)room(n3 * 8);
    u8 *flags = (u8 *)room((size_t)V), *gt = (u8 *)room(n1), *status = (u8 *)room((size_t)V);
    FILE *f = std::fopen(argv[5], "rb");
    if (!f || std::fread(lk, 8, n3, f) != n3 || std::fread(flags, 1, (size_t)V, f) != (size_t)V) return 2;
    std::fclose(f);
    for (size_t k = 0; k < n3; k++) post[k] = single[k] = -1.0;
    std::memset(gt, 77, n1);
    std::memset(status, 77, (size_t)V);
    const int errors = run_grid(grid, lk, flags, post, single, gt, status, V);
    f = std::fopen(argv[6], "wb");
    if (!f) return 2;
    std::fwrite(post, 8, n3, f), std::fwrite(single, 8, n3, f), std::fwrite(gt, 1, n1, f), std::fwrite(status, 1, (size_t)V, f);
    std::fclose(f);
    return errors ? 1 : 0;
}
"""


def build_host_kernel(tmp_path, src: str):
    """The generated source from its arithmetic helpers on (the PTX wrappers before them are the shim's) as host programs:
    (plain, under ThreadSanitizer or None)."""
    start = src.index("// x[0..2] / s, correctly rounded")
    body = src[start:].replace("extern __shared__ __align__(128) unsigned char smem_raw[];", "")
    head = src[:src.index("typedef unsigned int u32;")]  # the #defines of the tile shape (TB, NCOL, S3)
    cpp = str(tmp_path / "es_kernel_host.cpp")
    open(cpp, "w").write(head + SHIM + body + DRIVER)
    plain, tsan = str(tmp_path / "es_kernel_host"), str(tmp_path / "es_kernel_host_tsan")
    flags = ["-O1", "-std=c++20", "-ffp-contract=off", "-pthread", "-w"]
    subprocess.run(["g++"] + flags + ["-o", plain, cpp], check=True)
    if subprocess.run(["g++", "-fsanitize=thread", "-g"] + flags + ["-o", tsan, cpp], capture_output=True).returncode != 0:
        tsan = None
    return plain, tsan


def run_host_kernel(tmp_path, prog, grid, late_loads, late_stores, lk, fl, expect_ok=True):
    V, S = lk.shape[0], lk.shape[1]
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.ascontiguousarray(lk, np.float64).tobytes())
        f.write(np.ascontiguousarray(fl, np.uint8).tobytes())
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66 report_signal_unsafe=0")
    r = subprocess.run([prog, str(grid), str(late_loads), str(late_stores), str(V), inp, out], capture_output=True, text=True, env=env, timeout=600)
    if expect_ok:
        assert r.returncode == 0, f"host kernel exited {r.returncode}: {r.stderr[-2000:]}"
    raw = np.fromfile(out, np.uint8)
    n3 = V * S * 3 * 8
    post = raw[:n3].view(np.float64).reshape(V, S, 3)
    single = raw[n3:2 * n3].view(np.float64).reshape(V, S, 3)
    gt = raw[2 * n3:2 * n3 + V * S].reshape(V, S)
    status = raw[2 * n3 + V * S:]
    return post, single, gt, status


def check(got, want, what):
    post, single, gt, status = got
    assert np.array_equal(status, want["status"].astype(np.uint8)), what
    ok = want["status"] == 0
    assert np.array_equal(post[ok], want["post"][ok]), what
    assert np.array_equal(single[ok], want["single"][ok]), what
    assert np.array_equal(gt[ok], want["gt"][ok].astype(np.uint8)), what
    assert not post[~ok].any() and not single[~ok].any() and (gt[~ok] == 255).all(), what


@pytest.mark.parametrize("name,mrate,cols", [("half_sibs", 1e-7, None), ("ped14", 0.0, None), ("three_wives", 1e-7, None),
                                             ("ped14", 1e-7, [13, 2, 7, 0, 10, 5])])
def test_tile_pipeline_of_the_generated_kernel_against_an_adversarial_copy_engine(name, mrate, cols, tmp_path):
    ped = synth.PEDIGREES[name]()
    cols = ped.sequenced_cols() if cols is None else cols
    S = len(cols)
    V = 32 * 9 + 13  # ten tiles, the last one ragged
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, S + 1)]), V, seed=2718, x_fraction=0.3)
    lk[::7] = np.round(lk[::7])  # certain variants: the LRC gate keeps the pedigree out
    lk[5::11] = 0.0              # a variant that fails before the pedigree is looked at
    lk[64:96:2] = np.round(lk[64:96:2])  # a tile in which half of the lanes sit the pedigree out
    lk[128:160] = np.round(lk[128:160])  # a tile nobody wants the pedigree for: the pipeline point after the rule sets
    if mrate == 0.0:  # Mendel-impossible certain genotypes: the row sum vanishes INSIDE the pedigree, after `single` has left
        kid = next(i for i in range(ped.n) if ped.mids[i] != 0 and ped.ids[i] in [ped.ids[c] for c in cols])
        mother, father = list(ped.ids).index(ped.mids[kid]), list(ped.ids).index(ped.fids[kid])
        col_of = {row: c for c, row in enumerate(cols)}
        bad = np.arange(V) % 5 == 3
        for row, value in ((mother, (1.0, 0.0, 0.0)), (father, (1.0, 0.0, 0.0)), (kid, (0.0, 0.0, 1.0))):
            lk[bad, col_of[row]] = np.array(value)
    prm = fs.Params.default()
    prm.mrate = mrate
    want = O.run(ped, cols, lk, fl, method=O.ES, mrate=mrate)
    assert 0 < (want["status"] != 0).sum() < V
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, params=prm, device=-1) as e:
        src, _ = e.es_kernel()
    plain, tsan = build_host_kernel(tmp_path, src)
    # grids: one block walks all ten tiles; three blocks walk 4 / 3 / 3; more blocks than tiles
    for grid in (1, 3, 16):
        for late_loads, late_stores in ((0, 1), (1, 0)):
            got = run_host_kernel(tmp_path, plain, grid, late_loads, late_stores, lk, fl)
            check(got, want, f"{name} grid={grid} late_loads={late_loads} late_stores={late_stores}")
    if tsan is not None:  # the same under ThreadSanitizer: every shared-memory hand-over must be ordered by a barrier
        got = run_host_kernel(tmp_path, tsan, 3, 0, 1, lk, fl)
        check(got, want, f"{name} under ThreadSanitizer")


def test_the_adversary_notices_a_broken_pipeline(tmp_path):
    """The shim must be able to fail: the same kernel with the wait of the pipeline point removed hands buffer B to the next
    tile's rows while the previous tile's store has not read it -- with late stores the previous tile's posteriors come out wrong."""
    ped = synth.PEDIGREES["half_sibs"]()
    cols = ped.sequenced_cols()
    S = len(cols)
    V = 32 * 6
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, S + 1)]), V, seed=99, x_fraction=0.0)
    want = O.run(ped, cols, lk, fl, method=O.ES)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1) as e:
        src, _ = e.es_kernel()
    needle = "    if (p.lane == 0) bulk_wait_read();\n"
    assert src.count(needle) == 1
    plain, _ = build_host_kernel(tmp_path, src.replace(needle, ""))
    post, single, gt, status = run_host_kernel(tmp_path, plain, 1, 0, 1, lk, fl, expect_ok=False)
    assert not np.array_equal(post, want["post"])


def test_thread_sanitizer_notices_a_missing_syncwarp(tmp_path):
    """... and so must the sanitizer: without the __syncwarp between the lanes' last reads of buffer A and the next tile's copy
    into it (pipeline point), lane 0's copy races with the other lanes' reads."""
    ped = synth.PEDIGREES["half_sibs"]()
    cols = ped.sequenced_cols()
    S = len(cols)
    V = 32 * 6
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, S + 1)]), V, seed=99, x_fraction=0.0)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1) as e:
        src, _ = e.es_kernel()
    needle = "    fence_async_smem(); // order this lane's reads of A before the copy engine's writes to it\n    __syncwarp();\n"
    assert src.count(needle) == 1
    _, tsan = build_host_kernel(tmp_path, src.replace(needle, "    fence_async_smem();\n"))
    if tsan is None:
        pytest.skip("no ThreadSanitizer in this tool chain")
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(lk.tobytes() + fl.tobytes())
    r = subprocess.run([tsan, "1", "0", "1", str(V), inp, out], capture_output=True, text=True,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66"), timeout=600)
    assert r.returncode == 66 and "data race" in r.stderr
