// kernels.hpp -- launch interface between the engine (engine.cu) and the three method kernels.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../host/bn_plan.hpp"
#include "../host/es_program.hpp"
#include "../host/mcmc_plan.hpp"
#include "common.cuh"

namespace famseq {

// ---- Elston-Stewart (es_kernel.cu) ----------------------------------------------------------------
struct EsParams {
    RunConstants C;
    EsProgram prog;
};
size_t es_smem_bytes(const EsParams &P, int tb);
int es_pick_block(const EsParams &P, size_t smem_limit);
cudaError_t launch_es(const EsParams &P, const BatchPtrs &B, int tb, cudaStream_t stream);

// ---- Elston-Stewart, nuclear families (es_nuclear_kernel.cu) -------------------------------------------
constexpr int ES_NUCLEAR_MAX_CHILDREN = 5;
struct NuclearParams {
    RunConstants C;
    int32_t n_children;
    int32_t col_father, col_mother;               // input column or -1
    int32_t col_child[ES_NUCLEAR_MAX_CHILDREN];   // children in ped order
    int32_t male_child[ES_NUCLEAR_MAX_CHILDREN];
    int32_t io_probe; // FAMSEQ_ES_IO_PROBE=1: move the tiles but skip the arithmetic (measures the I/O ceiling; results are garbage)
};
cudaError_t launch_es_nuclear(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream);

// ---- Bayesian network (bn_kernel.cu) --------------------------------------------------------------
struct BnParams {
    RunConstants C;
    BnPlan plan;
};
size_t bn_smem_bytes(const BnParams &P);
cudaError_t launch_bn(const BnParams &P, const BatchPtrs &B, int sm_count, cudaStream_t stream);

// ---- Gibbs sampler (mcmc_kernel.cu) ---------------------------------------------------------------
struct McmcParams {
    RunConstants C;
    McmcPlan plan;
};
size_t mcmc_smem_bytes(const McmcParams &P, int tb);
int mcmc_pick_block(const McmcParams &P, size_t smem_limit, size_t smem_per_sm);
// fixup: only variants whose status byte is 2 are processed (what the specialised kernel of gibbs_jit.cu left over).
cudaError_t launch_mcmc(const McmcParams &P, const BatchPtrs &B, int tb, int burn, int rep, uint64_t seed,
                        int64_t v_offset, int sm_count, cudaStream_t stream, bool fixup = false);

} // namespace famseq
