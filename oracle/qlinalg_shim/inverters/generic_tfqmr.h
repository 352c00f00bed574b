// ORACLE / TEST INFRASTRUCTURE ONLY.
// quantum-linalg "inverters/generic_tfqmr.h": used only by
// /root/reference/tests/n11_wilson_test/wilson_test.cpp (solver survey, out of
// scope: SURVEY.md section 2 row 18).  Declared so the driver parses; it reports
// failure without iterating.
#ifndef QLINALG_SHIM_TFQMR
#define QLINALG_SHIM_TFQMR
#include "../blas/generic_vector.h"
#include "inverter_struct.h"
inline inversion_info minv_vector_tfqmr(complex<double>*, complex<double>*, int, int, double,
                                        matrix_op_cplx, void*, inversion_verbose_struct* = 0)
{
  inversion_info invif; invif.name = "TFQMR (not provided by the oracle shim)"; return invif;
}
#endif
