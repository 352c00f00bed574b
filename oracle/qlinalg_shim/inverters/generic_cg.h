// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "inverters/generic_cg.h".
// Call sites: /root/reference/multigrid/stateful_multigrid.h:928 (cg), :960 (cg_restart),
//   tests/n03_gauge_laplace_test/gauged_laplace.cpp:83.
// Algorithm (UNPINNED, defined here): textbook CG for Hermitian positive A,
//   r = b - A x (1 op), p = r, then per iteration Ap = A p (1 op),
//   alpha = <r|r>/<p|Ap>, x += alpha p, r -= alpha Ap, stop if |r| < rel_tol |b|,
//   beta = <r'|r'>/<r|r>, p = r + beta p;  resSq recomputed at exit (1 op).
// Restarted flavour: run CG in bursts of `restart_freq` iterations from the
// current x until converged or max_iter total iterations.
#ifndef QLINALG_SHIM_CG
#define QLINALG_SHIM_CG

#include "../blas/generic_vector.h"
#include "inverter_struct.h"

inline inversion_info minv_vector_cg(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                     matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  inversion_info invif;
  invif.name = "CG";
  complex<double>* r = allocate_vector<complex<double> >(size);
  complex<double>* p = allocate_vector<complex<double> >(size);
  complex<double>* Ap = allocate_vector<complex<double> >(size);
  const double bsqrt = sqrt(norm2sq(phi0, size));

  zero_vector(Ap, size);
  matrix_vector(Ap, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, Ap, r, size);
  copy_vector(p, r, size);
  double rsq = norm2sq(r, size);

  int k = 0;
  bool converged = sqrt(rsq) < eps * bsqrt;
  if (!converged) for (k = 1; k <= max_iter; k++)
  {
    zero_vector(Ap, size);
    matrix_vector(Ap, p, extra_info); invif.ops_count++;
    double alpha = rsq / real(dot(p, Ap, size));
    caxpy(alpha, p, phi, size);
    caxpy(-alpha, Ap, r, size);
    double rsqNew = norm2sq(r, size);
    print_verbosity_resid(verb, "CG", k, invif.ops_count, sqrt(rsqNew) / bsqrt);
    if (sqrt(rsqNew) < eps * bsqrt) { rsq = rsqNew; converged = true; break; }
    double beta = rsqNew / rsq;
    rsq = rsqNew;
    cxpay(r, beta, p, size);
  }
  if (k > max_iter) k = max_iter;

  zero_vector(Ap, size);
  matrix_vector(Ap, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(Ap, phi0, size);
  invif.iter = k;
  invif.success = converged;
  print_verbosity_summary(verb, "CG", invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  deallocate_vector(&r);
  deallocate_vector(&p);
  deallocate_vector(&Ap);
  return invif;
}

inline inversion_info minv_vector_cg_restart(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                             matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  inversion_info invif, total;
  total.name = "Restarted CG";
  const double bsqrt = sqrt(norm2sq(phi0, size));
  inversion_verbose_struct quiet;
  if (verb != 0) { quiet = *verb; if (quiet.verbosity == VERB_SUMMARY || quiet.verbosity == VERB_RESTART_DETAIL) quiet.verbosity = VERB_NONE; }
  do
  {
    int burst = max_iter - total.iter < restart_freq ? max_iter - total.iter : restart_freq;
    invif = minv_vector_cg(phi, phi0, size, burst, eps, matrix_vector, extra_info, &quiet);
    total.iter += invif.iter;
    total.ops_count += invif.ops_count;
    total.resSq = invif.resSq;
    print_verbosity_restart(verb, total.name, total.iter, total.ops_count, sqrt(total.resSq) / bsqrt);
  } while (total.iter < max_iter && !invif.success && sqrt(invif.resSq) > eps * bsqrt);
  total.success = invif.success || sqrt(invif.resSq) <= eps * bsqrt;
  print_verbosity_summary(verb, total.name, total.success, total.iter, total.ops_count, sqrt(total.resSq) / bsqrt);
  return total;
}

#endif
