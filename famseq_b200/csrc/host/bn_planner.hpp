// bn_planner.hpp -- host entry point that builds the BN enumeration plan (bn_plan.cpp).
#pragma once

#include <string>

#include "bn_plan.hpp"
#include "pedigree.hpp"

namespace famseq {

// Returns FS_OK or FS_E_TOO_LARGE.
int build_bn_plan(const Pedigree &ped, BnPlan &out, std::string &err);

} // namespace famseq
