// mcmc_plan.cpp -- see mcmc_plan.hpp.
#include "../../../include/famseq_b200.h"
#include "mcmc_plan.hpp"
#include "pedigree.hpp"

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
        const int first = k;
        for (int c : ped.children[i]) { // child order as in family.cpp:2128
            if (k >= MCMC_MAX_LINKS) {
                err = "The MCMC kernel supports at most " + std::to_string(MCMC_MAX_LINKS) + " parent-child links.";
                return FS_E_TOO_LARGE;
            }
            const int other = ped.mother[c] == i ? ped.father[c] : ped.mother[c];
            p.link[k++] = (uint16_t)(c | other << 7 | ped.male[c] << 14);
        }
        const bool founder = ped.founder(i);
        p.member[i] = (uint32_t)(founder ? 0 : ped.mother[i]) | (uint32_t)(founder ? 0 : ped.father[i]) << 7 | (uint32_t)(founder ? 1 : 0) << 14 |
                      (uint32_t)ped.male[i] << 15 | (uint32_t)first << 16 | (uint32_t)(k - first) << 24;
        p.col[i] = (int16_t)ped.col_of[i];
    }
    p.n_links = k;
    out = p;
    return FS_OK;
}

} // namespace famseq
