// quantum-mg on B200 -- plain BiCGstab on device vectors (quantum-linalg "inverters/generic_bicgstab.h"; the solver
// survey of /root/reference/tests/n11_wilson_test/wilson_test.cpp:185).  As in the oracle's restatement it is BiCGstab(1).
#ifndef QMG_B200_BICGSTAB
#define QMG_B200_BICGSTAB
#include "generic_bicgstab_l.h"
inline inversion_info minv_vector_bicgstab(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                           matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  inversion_info invif = minv_vector_bicgstab_l(phi, phi0, size, max_iter, eps, 1, matrix_vector, extra_info, verb);
  invif.name = "BiCGstab";
  return invif;
}
#endif
