// es_jit.hpp -- Elston-Stewart peeling of one loop-free pedigree as generated straight-line code (see es_jit.cu).
#pragma once

#include <string>

#include "kernels.hpp"

namespace famseq {

std::string es_jit_source(const EsParams &P);

// Source generation + NVRTC compilation; host only, safe to run on a worker thread.
int es_jit_build(const EsParams &P, std::string &cubin, std::string &log, std::string &err);

bool es_jit_fits(const EsParams &P, size_t smem_limit); // two tiles of 32 variants must fit in shared memory

struct EsJitKernel; // a loaded cubin
int es_jit_load(const EsParams &P, const std::string &cubin, EsJitKernel **out, std::string &err);
void es_jit_unload(EsJitKernel *k);
cudaError_t es_jit_launch(EsJitKernel *k, const BatchPtrs &B, cudaStream_t stream);

} // namespace famseq
