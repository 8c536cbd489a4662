// mcmc_plan.hpp -- neighbour lists for the single-site Gibbs sampler (family::estGenoProb,
// src/family.cpp:2098-2299): for every member, in ped order, its parents and the (child, other parent)
// pairs whose transmission factors enter its full conditional, packed into one descriptor word per member
// and per link so that a Gibbs step needs one uniform constant load for each.
// Shared between host/mcmc_plan.cpp and cuda/mcmc_kernel.cu.
#pragma once

#include <cstdint>

namespace famseq {

constexpr int MCMC_MAX_MEMBERS = 64; // genotype vector packed 2 bits/member in two 64-bit registers
constexpr int MCMC_MAX_LINKS = 255;

// member descriptor: bits 0-5 mother, 6-11 father, 12 founder, 13 male, 14-21 first link, 22-29 link count
// link descriptor  : bits 0-5 child, 6-11 the child's other parent, 12 child is male
struct McmcPlan {
    int32_t n = 0;
    int32_t n_links = 0;
    uint32_t member[MCMC_MAX_MEMBERS];
    uint16_t link[MCMC_MAX_LINKS + 1];
    int16_t col[MCMC_MAX_MEMBERS]; // input column or -1
};

} // namespace famseq
