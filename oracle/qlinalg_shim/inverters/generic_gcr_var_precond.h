// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "inverters/generic_gcr_var_precond.h":
// flexible (variably preconditioned) GCR, the outer Krylov method of the K-cycle.
// Call sites: /root/reference/multigrid/stateful_multigrid.h:980,987 and
//   tests/n13_wilson_kcycle/wilson_kcycle.cpp:459.  Preconditioner signature from
//   stateful_multigrid.h:734: void f(lhs, rhs, size, extra, verb), lhs overwritten.
// Algorithm (UNPINNED, defined here):
//   r = b - A x (1 op); z = M^{-1} r; p_0 = z; Ap_0 = A z (1 op)
//   iteration k: alpha = <Ap_k|r>/<Ap_k|Ap_k>; x += alpha p_k; r -= alpha Ap_k;
//     stop if |r| < rel_tol |b|;  z = M^{-1} r; Az = A z (1 op);
//     p_{k+1} = z + sum_i beta_i p_i, Ap_{k+1} = Az + sum_i beta_i Ap_i,
//     beta_i = -<Ap_i|Az>/<Ap_i|Ap_i>
//   resSq = |b - A x|^2 recomputed (1 op)
// ops_count counts applications of A by this solver only (not those inside M^{-1}).
#ifndef QLINALG_SHIM_GCR_VAR_PRECOND
#define QLINALG_SHIM_GCR_VAR_PRECOND

#include <vector>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

inline inversion_info minv_vector_gcr_var_precond(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                                  matrix_op_cplx matrix_vector, void* extra_info,
                                                  precond_op_cplx precond_matrix_vector, void* precond_info,
                                                  inversion_verbose_struct* verb = 0)
{
  inversion_info invif;
  invif.name = "VPGCR";
  inversion_verbose_struct verb_prec;
  shuffle_verbosity_precond(&verb_prec, verb);

  complex<double>* r = allocate_vector<complex<double> >(size);
  complex<double>* z = allocate_vector<complex<double> >(size);
  complex<double>* Az = allocate_vector<complex<double> >(size);
  std::vector<complex<double>*> p, Ap;
  std::vector<double> ApNormSq;
  const double bsqrt = sqrt(norm2sq(phi0, size));

  zero_vector(Az, size);
  matrix_vector(Az, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, Az, r, size);
  double rsq = norm2sq(r, size);

  int k = 0;
  bool converged = sqrt(rsq) < eps * bsqrt;
  if (!converged && max_iter > 0)
  {
    zero_vector(z, size);
    precond_matrix_vector(z, r, size, precond_info, &verb_prec);
    p.push_back(allocate_vector<complex<double> >(size));
    Ap.push_back(allocate_vector<complex<double> >(size));
    copy_vector(p[0], z, size);
    zero_vector(Ap[0], size);
    matrix_vector(Ap[0], p[0], extra_info); invif.ops_count++;
    for (k = 1; k <= max_iter; k++)
    {
      const int c = k - 1;
      ApNormSq.push_back(norm2sq(Ap[c], size));
      complex<double> alpha = dot(Ap[c], r, size) / ApNormSq[c];
      caxpy(alpha, p[c], phi, size);
      caxpy(-alpha, Ap[c], r, size);
      rsq = norm2sq(r, size);
      print_verbosity_resid(verb, "VPGCR", k, invif.ops_count, sqrt(rsq) / bsqrt);
      if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
      if (k == max_iter) break;

      zero_vector(z, size);
      precond_matrix_vector(z, r, size, precond_info, &verb_prec);
      zero_vector(Az, size);
      matrix_vector(Az, z, extra_info); invif.ops_count++;
      p.push_back(allocate_vector<complex<double> >(size));
      Ap.push_back(allocate_vector<complex<double> >(size));
      copy_vector(p[k], z, size);
      copy_vector(Ap[k], Az, size);
      for (int i = 0; i < k; i++)
      {
        complex<double> beta = -dot(Ap[i], Az, size) / ApNormSq[i];
        caxpy(beta, p[i], p[k], size);
        caxpy(beta, Ap[i], Ap[k], size);
      }
    }
  }
  if (k > max_iter) k = max_iter;

  zero_vector(Az, size);
  matrix_vector(Az, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(Az, phi0, size);
  invif.iter = k;
  invif.success = converged;
  print_verbosity_summary(verb, "VPGCR", invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  for (size_t i = 0; i < p.size(); i++) { deallocate_vector(&p[i]); deallocate_vector(&Ap[i]); }
  deallocate_vector(&r);
  deallocate_vector(&z);
  deallocate_vector(&Az);
  return invif;
}

inline inversion_info minv_vector_gcr_var_precond_restart(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                                          matrix_op_cplx matrix_vector, void* extra_info,
                                                          precond_op_cplx precond_matrix_vector, void* precond_info,
                                                          inversion_verbose_struct* verb = 0)
{
  inversion_info invif, total;
  total.name = "Restarted VPGCR";
  const double bsqrt = sqrt(norm2sq(phi0, size));
  inversion_verbose_struct quiet;
  if (verb != 0) { quiet = *verb; if (quiet.verbosity == VERB_SUMMARY || quiet.verbosity == VERB_RESTART_DETAIL) quiet.verbosity = VERB_NONE; }
  do
  {
    int burst = max_iter - total.iter < restart_freq ? max_iter - total.iter : restart_freq;
    invif = minv_vector_gcr_var_precond(phi, phi0, size, burst, eps, matrix_vector, extra_info, precond_matrix_vector, precond_info, &quiet);
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
