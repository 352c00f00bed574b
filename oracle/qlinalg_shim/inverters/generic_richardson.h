// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "inverters/generic_richardson.h".
// Call site: /root/reference/tests/n22_wilson_kcycle_adaptive/wilson_kcycle.cpp:289
//   minv_vector_richardson(x, b, n, max_iter, rel_tol, omega, check_freq, op, extra, verb)
// Algorithm (UNPINNED, defined here):
//   repeat k = 1..max_iter:  r = b - A x (1 op) ; x += omega r ;
//     every check_freq iterations (and never otherwise) test |r| < rel_tol |b|
//   resSq = |b - A x|^2 recomputed (1 op)
#ifndef QLINALG_SHIM_RICHARDSON
#define QLINALG_SHIM_RICHARDSON

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
    zero_vector(Ax, size);
    matrix_vector(Ax, phi, extra_info); invif.ops_count++;
    // Ax <- r = b - Ax
    caxpby(1.0, phi0, -1.0, Ax, size);
    if (check_freq > 0 && k % check_freq == 0)
    {
      double rsq = norm2sq(Ax, size);
      print_verbosity_resid(verb, "Richardson", k, invif.ops_count, sqrt(rsq) / bsqrt);
      if (sqrt(rsq) < eps * bsqrt) { converged = true; k--; break; }
    }
    caxpy(omega, Ax, phi, size);
  }
  if (k > max_iter) k = max_iter;
  zero_vector(Ax, size);
  matrix_vector(Ax, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(Ax, phi0, size);
  invif.iter = k;
  invif.success = converged || (sqrt(invif.resSq) < eps * bsqrt);
  print_verbosity_summary(verb, "Richardson", invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);
  deallocate_vector(&Ax);
  return invif;
}

#endif
