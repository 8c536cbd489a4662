// mcmc_plan.hpp -- neighbour lists for the single-site Gibbs sampler (family::estGenoProb,
// src/family.cpp:2098-2299): for every member, in ped order, its parents and the (child, other parent)
// pairs whose transmission factors enter its full conditional, packed into one descriptor word per member
// and per link so that a Gibbs step needs one uniform constant load for each.
// Shared between host/mcmc_plan.cpp and cuda/mcmc_kernel.cu.
#pragma once

#include <cstdint>
#include <string>

namespace famseq {

struct Pedigree; // host/pedigree.hpp

constexpr int MCMC_MAX_MEMBERS = 128; // genotype vector packed 2 bits/member in two (<= 64 members) or four 64-bit registers
constexpr int MCMC_MAX_LINKS = 255;

// member descriptor: bits 0-6 mother, 7-13 father, 14 founder, 15 male, 16-23 first link, 24-31 link count
// link descriptor  : bits 0-6 child, 7-13 the child's other parent, 14 child is male
struct McmcPlan {
    int32_t n = 0;
    int32_t n_links = 0;
    uint32_t member[MCMC_MAX_MEMBERS];
    uint16_t link[MCMC_MAX_LINKS + 1];
    int16_t col[MCMC_MAX_MEMBERS]; // input column or -1
};

// Host entry point (mcmc_plan.cpp): builds the neighbour lists.  Returns FS_OK or FS_E_TOO_LARGE.
int build_mcmc_plan(const Pedigree &ped, McmcPlan &out, std::string &err);

} // namespace famseq
