// mcmc_plan.hpp -- neighbour lists for the single-site Gibbs sampler (family::estGenoProb,
// src/family.cpp:2098-2299): for every member, in ped order, its parents and the (child, other parent)
// pairs whose transmission factors enter its full conditional.
// Shared between host/mcmc_plan.cpp and cuda/mcmc_kernel.cu.
#pragma once

#include <cstdint>

namespace famseq {

constexpr int MCMC_MAX_MEMBERS = 64; // genotype vector packed 2 bits/member in two 64-bit registers
constexpr int MCMC_MAX_LINKS = 512;

struct McmcPlan {
    int32_t n = 0;
    int32_t n_links = 0;
    int8_t mother[MCMC_MAX_MEMBERS]; // -1 for founders
    int8_t father[MCMC_MAX_MEMBERS];
    uint8_t male[MCMC_MAX_MEMBERS];
    int16_t col[MCMC_MAX_MEMBERS];        // input column or -1
    uint16_t link_begin[MCMC_MAX_MEMBERS + 1]; // links of member i: [link_begin[i], link_begin[i+1])
    uint8_t link_child[MCMC_MAX_LINKS];
    uint8_t link_other[MCMC_MAX_LINKS]; // the child's other parent
};

} // namespace famseq
