// es_nuclear_kernel.cu -- dispatch of the nuclear-family Elston-Stewart kernel (es_nuclear_kernel.cuh) by sibship size.  The
// template is instantiated in one translation unit per number of children (es_nuclear_nc1.cu ... nc5.cu: they build in
// parallel; a single unit took four minutes).
#include "kernels.hpp"

namespace famseq {

cudaError_t launch_es_nuclear_nc1(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream);
cudaError_t launch_es_nuclear_nc2(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream);
cudaError_t launch_es_nuclear_nc3(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream);
cudaError_t launch_es_nuclear_nc4(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream);
cudaError_t launch_es_nuclear_nc5(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream);

// Variants per block (= per TMA tile).  Measured on the trio (profiles/es_tb_check.sh), fraction of the HBM peak:
// 32 -> 0.946, 64 -> 0.933, 128 -> 0.904, 256 -> 0.72: many one-warp blocks per SM interleave their load / compute /
// store phases best.  A persistent, double-buffered variant (one block per slot looping over tiles, next tile
// requested before the current one is computed) was slower (0.854): its barriers serialise what the block
// scheduler overlaps for free.
cudaError_t launch_es_nuclear(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream) {
    if (B.V <= 0) return cudaSuccess;
    switch (P.n_children) { // one tile size: 32 variants, one warp per block (see above)
    case 1: return launch_es_nuclear_nc1(P, B, stream);
    case 2: return launch_es_nuclear_nc2(P, B, stream);
    case 3: return launch_es_nuclear_nc3(P, B, stream);
    case 4: return launch_es_nuclear_nc4(P, B, stream);
    case 5: return launch_es_nuclear_nc5(P, B, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace famseq
