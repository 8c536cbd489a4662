// mcmc_planner.hpp -- host entry point that builds the Gibbs neighbour lists (mcmc_plan.cpp).
#pragma once

#include <string>

#include "mcmc_plan.hpp"
#include "pedigree.hpp"

namespace famseq {

// Returns FS_OK or FS_E_TOO_LARGE.
int build_mcmc_plan(const Pedigree &ped, McmcPlan &out, std::string &err);

} // namespace famseq
