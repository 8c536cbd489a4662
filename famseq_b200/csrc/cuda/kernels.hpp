// kernels.hpp -- launch interface between the engine (engine.cu) and the three method kernels.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../host/bn_plan.hpp"
#include "../host/es_program.hpp"
#include "../host/mcmc_plan.hpp"
#include "common.cuh"

struct fs_phred_fix; // include/famseq_b200.h

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
    int32_t allow_ident; // use the kernel specialised for the identity column map when it applies (FAMSEQ_ES_IDENT=0 turns it off)
    int32_t stream_tiles; // compact input: tiles per block, each prefetched under the one before (0: 4 for trios, 8 otherwise; -1: the one-tile kernel; FAMSEQ_ES_STREAM)
};
// Handles B.pl (compact input, decoded through B.lut) and B.single == nullptr itself.
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
// Launch-time tuning, read from the environment once in fs_create (FAMSEQ_MCMC_BLOCKS, FAMSEQ_MCMC_WGLOBAL).
struct McmcTuning {
    int blocks_cap = 3; // resident blocks per SM when the own factors live in L2
    int wglobal = -1;   // own factors in L2 (1) / shared memory (0) / by pedigree size (-1)
};
// fixup: only variants whose status byte is 2 are processed (what the specialised kernel of gibbs_jit.cu left over);
// *fixup_count (device, may be null) is incremented once per such variant.
cudaError_t launch_mcmc(const McmcParams &P, const BatchPtrs &B, int tb, int burn, int rep, uint64_t seed,
                        int64_t v_offset, int sm_count, cudaStream_t stream, const McmcTuning &tune, bool fixup = false,
                        unsigned long long *fixup_count = nullptr);

// ---- compact input (engine.cu) ----------------------------------------------------------------------
// lk[k] = lut[pl[k]] for k < n: expands fs_run_pl input for the kernels that read FP64 likelihoods.
cudaError_t launch_pl_decode(const uint16_t *pl, const double *lut, double *lk, int64_t n, cudaStream_t stream);

// ---- Phred encoding of posteriors (phred_kernel.cu) --------------------------------------------------------
// out[k] = packed six-digit Phred code of p[k] (include/famseq_b200.h, FS_PHRED_*); values the device does not decide are
// appended to `fixes` (index0 + k, exact double), *n_fixes counts them even beyond `capacity`.
cudaError_t launch_phred_pack(const double *p, uint32_t *out, int64_t n, int64_t index0, struct ::fs_phred_fix *fixes, int64_t capacity,
                              unsigned long long *n_fixes, cudaStream_t stream);

} // namespace famseq
