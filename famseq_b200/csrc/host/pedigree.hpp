// pedigree.hpp -- host-side pedigree model: topology, transmission tables, loop detection.
//
// Behaviour follows the reference's family set-up (src/family.cpp:204-350, :383-550) but the
// representation is flat index arrays that the pedigree compilers (es_program.cpp, bn_plan.cpp,
// mcmc_plan.cpp) turn into device programs.  Built once per run.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace famseq {

enum TableKind { TAB_AUTO = 0, TAB_XF = 1, TAB_XM = 2 };

struct Pedigree {
    int n = 0;                       // members
    std::vector<int> mother, father; // ped row of each parent, -1 for founders (family.cpp:291-350)
    std::vector<int> male;           // 1 when gender == 1 (the reference treats everything else as female)
    std::vector<std::vector<int>> children; // ped order
    std::vector<std::vector<int>> spouses;  // order of first joint child
    std::vector<int> cols;           // [s] ped row of every sequenced input column (mapV2P >= 0 entries)
    std::vector<int> col_of;         // [n] input column of a member or -1 (mapP2V)
    bool has_loop = false;           // marriage-node graph has a cycle => Elston-Stewart undefined

    int s() const { return (int)cols.size(); }
    bool founder(int i) const { return mother[i] < 0; }
    int n_founders() const;
};

// Error codes are the FS_E_* values of include/famseq_b200.h; `err` receives the text.
int build_pedigree(int n, const int32_t *id, const int32_t *mother_id, const int32_t *father_id,
                   const int32_t *gender, int s, const int32_t *cols, Pedigree &out, std::string &err);

// Mendelian transmission tables t[g*9 + a*3 + b] = Pr(child g | mother a, father b) for mutation
// rate mu: autosome (family.cpp:447-550), X daughter (:383-416), X son (:418-445).
void build_tables(double mu, double tab[3][27]);

} // namespace famseq
