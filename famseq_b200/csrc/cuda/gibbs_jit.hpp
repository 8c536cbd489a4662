// gibbs_jit.hpp -- pedigree-specialised Gibbs sampler: the engine writes CUDA C++ for ONE pedigree (every member,
// parent and child index a literal, the member loop unrolled, genotypes in named registers), compiles it for
// sm_100a with NVRTC when the first large MCMC batch arrives and launches the resulting kernel.  Semantics and random
// stream are those of the table-driven kernel in mcmc_kernel.cu (same chains; posteriors equal to a few ulps); see gibbs_jit.cu.
#pragma once

#include <string>

#include "kernels.hpp"

namespace famseq {

// Per chain and member the sampler caches the outcome of the last evaluation of the member's full conditional as two integer
// draw thresholds: for the first n_p_reg members in registers, for the others in thread-private shared-memory columns
// (see gibbs_jit.cu).
struct GibbsJitConfig {
    int cached = 1;  // 1: cached conditionals + integer draws (generator 1), 0: dense sweeps (generator 2, round 1's kernel)
    int tb = 0;      // chains (threads) per block
    int blocks = 1;  // resident blocks per SM the register budget is sized for
    int n_p_reg = 0; // cached: members whose thresholds live in registers
    // dense: per chain 3 own factors and 3 accumulators per member, placed in ped order: the first n_*_reg in registers, the
    // next n_*_smem in shared memory, the rest in a block-private global scratch (L2)
    int n_acc_reg = 0, n_acc_smem = 0;
    int n_lk_reg = 0, n_lk_smem = 0;
    int prefetch = 1; // dense: own factors read from the scratch are requested this many members ahead
};

// Layout heuristic (FAMSEQ_JIT_CACHED=0/1 picks the generator; FAMSEQ_JIT_TB / _BLOCKS, cached: _PREG, dense: _RACC / _SACC / _RLK / _SLK / _PF).
GibbsJitConfig gibbs_jit_default_config(const McmcParams &P);
GibbsJitConfig gibbs_jit_config(const McmcParams &P, int cached); // the layout of one generator
bool gibbs_jit_generator_forced();                                // FAMSEQ_JIT_CACHED is set: no pilot, no switching

std::string gibbs_jit_source(const McmcParams &P, const GibbsJitConfig &cfg);

// Source -> sm_100a cubin (needs libnvrtc, no device).  `log` receives the compiler output (ptxas -v included).
int gibbs_jit_compile(const std::string &source, std::string &cubin, std::string &log, std::string &err);

// Source generation + compilation for one pedigree; host only, safe to run on a worker thread.
int gibbs_jit_build(const McmcParams &P, const GibbsJitConfig &cfg, std::string &cubin, std::string &log, std::string &err);

struct GibbsJitKernel; // a loaded cubin
int gibbs_jit_load(const McmcParams &P, const GibbsJitConfig &cfg, const std::string &cubin, GibbsJitKernel **out, std::string &err);
void gibbs_jit_unload(GibbsJitKernel *k);
// vote_stats (device, may be null): the cached generator adds {groups redone member by member, groups} of the sampling sweeps
cudaError_t gibbs_jit_launch(GibbsJitKernel *k, const BatchPtrs &B, int burn, int rep, uint64_t seed, int64_t v_offset,
                             int sm_count, cudaStream_t stream, unsigned long long *vote_stats = nullptr);

} // namespace famseq
