// quantum-mg on B200 -- relaxed minimal-residual smoother on device vectors.
// Signature from /root/reference/multigrid/stateful_multigrid.h:860:
//   minv_vector_minres(x, b, n, max_iter, rel_tol, omega, op, extra[, verb]).
// quantum-linalg (where the reference takes it from) is un-vendored; the iteration is the one the
// oracle states (oracle/qlinalg_shim/inverters/generic_minres.h) so that iteration counts compare:
//   r = b - A x ; repeat { p = A r ; alpha = <p|r>/<p|p> ; x += w alpha r ; r -= w alpha p } ; true residual.
// Per iteration: one operator apply, one fused (dot, norm) pass and one fused (x, r, |r|^2) update
// -- 3 launches and 2 scalar read-backs instead of 1 + 5 BLAS sweeps.
#ifndef QMG_B200_MINRES
#define QMG_B200_MINRES

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

  matrix_vector(p, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, p, r, size);
  double rsq = norm2sq(r, size);

  int k = 0;
  bool converged = false;
  if (max_iter <= 0 || sqrt(rsq) < eps * bsqrt) converged = (sqrt(rsq) < eps * bsqrt);
  else for (k = 1; k <= max_iter; k++)
  {
    matrix_vector(p, r, extra_info); invif.ops_count++;
    // alpha = omega <Ar|r> / <Ar|Ar> ; x += alpha r ; r -= alpha A r ; |r|^2 -- alpha is formed on the device, one host wait
    // (x is updated from the old r inside the same thread)
    double step[4];
    QMG_CHK(qmg_step_xr_norm(omega, qmg_host::P(r), qmg_host::P(p), qmg_host::P(phi), qmg_host::P(r), size, step));
    rsq = step[0];
    qmg_host::say(verb, VERB_DETAIL, "MR", "", false, false, k, invif.ops_count, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
  }
  if (k > max_iter) k = max_iter;

  matrix_vector(p, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(p, phi0, size);
  invif.iter = k;
  invif.success = converged;
  qmg_host::say(verb, VERB_SUMMARY, "MR", "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  deallocate_vector(&r);
  deallocate_vector(&p);
  return invif;
}

#endif
