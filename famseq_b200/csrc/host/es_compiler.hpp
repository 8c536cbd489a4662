// es_compiler.hpp -- host entry point of the Elston-Stewart pedigree compiler (es_program.cpp).
#pragma once

#include <string>

#include "es_program.hpp"
#include "pedigree.hpp"

namespace famseq {

// Returns FS_OK, FS_E_LOOP (pedigree is not peelable) or FS_E_TOO_LARGE; `err` receives the text.
int compile_es_program(const Pedigree &ped, EsProgram &out, std::string &err);

} // namespace famseq
