// mcmc_plan.cpp -- see mcmc_plan.hpp.
#include <cstring>

#include "../../../include/famseq_b200.h"
#include "mcmc_planner.hpp"

namespace famseq {

int build_mcmc_plan(const Pedigree &ped, McmcPlan &out, std::string &err) {
    McmcPlan p{};
    if (ped.n > MCMC_MAX_MEMBERS) {
        err = "The MCMC kernel supports pedigrees of up to " + std::to_string(MCMC_MAX_MEMBERS) + " members; this one has " +
              std::to_string(ped.n) + ".";
        return FS_E_TOO_LARGE;
    }
    p.n = ped.n;
    int k = 0;
    for (int i = 0; i < ped.n; i++) {
        p.mother[i] = (int8_t)ped.mother[i];
        p.father[i] = (int8_t)ped.father[i];
        p.male[i] = (uint8_t)ped.male[i];
        p.col[i] = (int16_t)ped.col_of[i];
        p.link_begin[i] = (uint16_t)k;
        for (int c : ped.children[i]) { // child order as in family.cpp:2128
            if (k >= MCMC_MAX_LINKS) {
                err = "The MCMC kernel supports at most " + std::to_string(MCMC_MAX_LINKS) + " parent-child links.";
                return FS_E_TOO_LARGE;
            }
            p.link_child[k] = (uint8_t)c;
            p.link_other[k] = (uint8_t)(ped.mother[c] == i ? ped.father[c] : ped.mother[c]);
            k++;
        }
    }
    p.link_begin[ped.n] = (uint16_t)k;
    p.n_links = k;
    out = p;
    return FS_OK;
}

} // namespace famseq
