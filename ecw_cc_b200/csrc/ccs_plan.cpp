// ccs_plan.cpp — plan builders for the CCS (singles) ground- and excited-state
// residual path (CCS.py:23-1518).
#include "ccsd_plan.h"

namespace ecw {

bool build_ccs_plan(Plan& P, const Sizes& z, const std::string& func, int flags) {
  (void)P; (void)z; (void)func; (void)flags;
  return false;
}

}  // namespace ecw
