// ccsd_plan.h — plan builders (host-only) for the CCSD / CCS residual path.
#pragma once
#include "plan.h"

namespace ecw {

struct Sizes {
  int64_t nocc = 0, nvir = 0;
  int rank = 0, world = 1;   // vvvv_p is row-sharded over the packed virtual pair index; heavy GEMMs owner-computes
};

// mode flags mirror the reference keyword arguments of GCC.tupdate/lupdate
// (CCSD.py:248, :419): has_alpha <=> `alpha is not None`, equation <=> `equation=True`.
void build_ccsd_tupdate(Plan& P, const Sizes& z, int has_alpha, int equation);
void build_ccsd_lupdate(Plan& P, const Sizes& z, int has_alpha, int equation);
// general variants: amplitudes not assumed antisymmetric (only the integrals are)
void build_ccsd_tupdate_general(Plan& P, const Sizes& z, int has_alpha, int equation);
void build_ccsd_lupdate_general(Plan& P, const Sizes& z, int has_alpha, int equation);
void build_ccsd_gamma(Plan& P, const Sizes& z);
void build_ccsd_energy(Plan& P, const Sizes& z);

// CCS entry points (ccs_plan.cpp); returns false for an unknown function name
bool build_ccs_plan(Plan& P, const Sizes& z, const std::string& func, int flags);

}  // namespace ecw
