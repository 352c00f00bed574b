// quantum-mg on B200 -- conjugate gradients on device vectors (normal-equation coarsest solves,
// /root/reference/multigrid/stateful_multigrid.h:921-971; tests/n03, n04, n17, n21).
// Iteration as stated by the oracle (oracle/qlinalg_shim/inverters/generic_cg.h).
#ifndef QMG_B200_CG
#define QMG_B200_CG

#include "generic_gcr.h"

inline inversion_info minv_vector_cg(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                     matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  using namespace qmg_host;
  inversion_info invif;
  invif.name = "CG";
  complex<double>* r = allocate_vector<complex<double> >(size);
  complex<double>* p = allocate_vector<complex<double> >(size);
  complex<double>* Ap = allocate_vector<complex<double> >(size);
  const double bsqrt = sqrt(norm2sq(phi0, size));

  matrix_vector(Ap, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, Ap, r, size);
  copy_vector(p, r, size);
  double rsq = norm2sq(r, size);

  int k = 0;
  bool converged = sqrt(rsq) < eps * bsqrt;
  if (!converged) for (k = 1; k <= max_iter; k++)
  {
    matrix_vector(Ap, p, extra_info); invif.ops_count++;
    const double alpha = rsq / real(dot(p, Ap, size));
    double rsqNew = 0.0;
    QMG_CHK(qmg_update_xr_norm(alpha, 0.0, P(p), P(Ap), P(phi), P(r), size, &rsqNew));
    say(verb, VERB_DETAIL, "CG", "", false, false, k, invif.ops_count, sqrt(rsqNew) / bsqrt);
    if (sqrt(rsqNew) < eps * bsqrt) { rsq = rsqNew; converged = true; break; }
    const double beta = rsqNew / rsq;
    rsq = rsqNew;
    cxpay(r, beta, p, size);
  }
  if (k > max_iter) k = max_iter;

  matrix_vector(Ap, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(Ap, phi0, size);
  invif.iter = k;
  invif.success = converged;
  say(verb, VERB_SUMMARY, "CG", "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  deallocate_vector(&r);
  deallocate_vector(&p);
  deallocate_vector(&Ap);
  return invif;
}

inline inversion_info minv_vector_cg_restart(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                             matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  return qmg_host::restarted("Restarted CG", phi0, size, max_iter, eps, restart_freq, verb,
    [&](int burst, inversion_verbose_struct* quiet, qmg_host::SolveHints*) { return minv_vector_cg(phi, phi0, size, burst, eps, matrix_vector, extra_info, quiet); });
}

#endif
