// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "inverters/generic_minres.h":
// relaxed minimal-residual (MR) iteration.  Call site fixing the signature:
// /root/reference/multigrid/stateful_multigrid.h:860
//   minv_vector_minres(x, b, n, max_iter, rel_tol, omega, op, extra[, verb]).
// Algorithm (solver internals are UNPINNED -- quantum-linalg is absent; this
// restatement is the definition both the CPU oracle and the GPU path follow):
//   r = b - A x                                  (1 op)
//   repeat k = 1..max_iter:
//     p = A r                                    (1 op)
//     alpha = <p|r> / <p|p>
//     x += omega alpha r ;  r -= omega alpha p
//     stop if |r| < rel_tol |b|
//   resSq = |b - A x|^2 recomputed               (1 op)
#ifndef QLINALG_SHIM_MINRES
#define QLINALG_SHIM_MINRES

#include "../blas/generic_vector.h"
#include "inverter_struct.h"

inline inversion_info minv_vector_minres(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, double omega,
                                         matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  inversion_info invif;
  invif.name = "MR";
  complex<double>* r = allocate_vector<complex<double> >(size);
  complex<double>* p = allocate_vector<complex<double> >(size);
  const double bsqrt = sqrt(norm2sq(phi0, size));

  zero_vector(p, size);
  matrix_vector(p, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, p, r, size);
  double rsq = norm2sq(r, size);

  int k = 0;
  bool converged = false;
  if (max_iter <= 0 || sqrt(rsq) < eps * bsqrt) converged = (sqrt(rsq) < eps * bsqrt);
  else for (k = 1; k <= max_iter; k++)
  {
    zero_vector(p, size);
    matrix_vector(p, r, extra_info); invif.ops_count++;
    complex<double> alpha = dot(p, r, size) / norm2sq(p, size);
    caxpy(omega * alpha, r, phi, size);
    caxpy(-omega * alpha, p, r, size);
    rsq = norm2sq(r, size);
    print_verbosity_resid(verb, "MR", k, invif.ops_count, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
  }
  if (k > max_iter) k = max_iter;

  zero_vector(p, size);
  matrix_vector(p, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(p, phi0, size);
  invif.iter = k;
  invif.success = converged;
  print_verbosity_summary(verb, "MR", invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  deallocate_vector(&r);
  deallocate_vector(&p);
  return invif;
}

#endif
