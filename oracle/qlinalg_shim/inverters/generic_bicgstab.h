// ORACLE / TEST INFRASTRUCTURE ONLY.
// quantum-linalg "inverters/generic_bicgstab.h": used only by
// /root/reference/tests/n11_wilson_test/wilson_test.cpp:118 (a solver survey,
// outside the hot path).  Provided as BiCGstab(1) so that driver compiles.
#ifndef QLINALG_SHIM_BICGSTAB
#define QLINALG_SHIM_BICGSTAB
#include "generic_bicgstab_l.h"
inline inversion_info minv_vector_bicgstab(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                           matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  return minv_vector_bicgstab_l(phi, phi0, size, max_iter, eps, 1, matrix_vector, extra_info, verb);
}
#endif
