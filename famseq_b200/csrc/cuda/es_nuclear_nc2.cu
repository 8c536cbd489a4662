// es_nuclear_nc2.cu -- the nuclear-family Elston-Stewart kernel (es_nuclear_kernel.cuh) for sibships of 2.
#include "es_nuclear_kernel.cuh"

namespace famseq {
cudaError_t launch_es_nuclear_nc2(const NuclearParams &P, const BatchPtrs &B, cudaStream_t stream) { return launch_nc<2, 32>(P, B, stream); }
} // namespace famseq
