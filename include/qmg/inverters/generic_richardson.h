// quantum-mg on B200 -- relaxed Richardson iteration on device vectors (adaptive-setup relaxation,
// /root/reference/tests/n22_wilson_kcycle_adaptive/wilson_kcycle.cpp:289: 10 iterations, omega 0.33, check every 250).
// x += omega (b - A x); the residual norm is only reduced every check_freq iterations, as the oracle states it
// (oracle/qlinalg_shim/inverters/generic_richardson.h).
#ifndef QMG_B200_RICHARDSON
#define QMG_B200_RICHARDSON

#include "../blas/generic_vector.h"
#include "inverter_struct.h"

inline inversion_info minv_vector_richardson(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, double omega, int check_freq,
                                             matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  inversion_info invif;
  invif.name = "Richardson";
  complex<double>* Ax = allocate_vector<complex<double> >(size);
  const double bsqrt = sqrt(norm2sq(phi0, size));
  bool converged = false;
  int k;
  for (k = 1; k <= max_iter; k++)
  {
    matrix_vector(Ax, phi, extra_info); invif.ops_count++;
    if (check_freq > 0 && k % check_freq == 0)
    {
      const double rsq = diffnorm2sq(Ax, phi0, size);
      qmg_host::say(verb, VERB_DETAIL, "Richardson", "", false, false, k, invif.ops_count, sqrt(rsq) / bsqrt);
      if (sqrt(rsq) < eps * bsqrt) { converged = true; k--; break; }
    }
    // x += omega (b - A x) in one pass
    caxpbypz(omega, phi0, -omega, Ax, phi, size);
  }
  if (k > max_iter) k = max_iter;
  matrix_vector(Ax, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(Ax, phi0, size);
  invif.iter = k;
  invif.success = converged || (sqrt(invif.resSq) < eps * bsqrt);
  qmg_host::say(verb, VERB_SUMMARY, "Richardson", "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);
  deallocate_vector(&Ax);
  return invif;
}

#endif
