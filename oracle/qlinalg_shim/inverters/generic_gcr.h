// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "inverters/generic_gcr.h".
// Call sites: /root/reference/multigrid/stateful_multigrid.h:915 (gcr), :947 (gcr_restart),
//   tests/n18_rbjacobi_stencil_test/rbjacobi_stencil_test.cpp:154.
// Algorithm (UNPINNED, defined here): GCR, Saad "Iterative Methods" Alg. 6.21,
// full (untruncated) orthogonalisation of A p against all previous A p_i:
//   r = b - A x (1 op); p_0 = r; Ap_0 = A p_0 (1 op)
//   iteration k: alpha = <Ap_k|r>/<Ap_k|Ap_k>; x += alpha p_k; r -= alpha Ap_k;
//     stop if |r| < rel_tol |b|;  Ar = A r (1 op);
//     p_{k+1} = r + sum_i beta_i p_i, Ap_{k+1} = Ar + sum_i beta_i Ap_i,
//     beta_i = -<Ap_i|Ar>/<Ap_i|Ap_i> (classical Gram-Schmidt against the stored set)
//   resSq = |b - A x|^2 recomputed (1 op)
// Restarted flavour: bursts of restart_freq iterations from the current x,
// tolerance always relative to |b|.
#ifndef QLINALG_SHIM_GCR
#define QLINALG_SHIM_GCR

#include <vector>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

inline inversion_info minv_vector_gcr(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                      matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  inversion_info invif;
  invif.name = "GCR";
  complex<double>* r = allocate_vector<complex<double> >(size);
  complex<double>* Ar = allocate_vector<complex<double> >(size);
  std::vector<complex<double>*> p, Ap;
  std::vector<double> ApNormSq;
  const double bsqrt = sqrt(norm2sq(phi0, size));

  zero_vector(Ar, size);
  matrix_vector(Ar, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, Ar, r, size);
  double rsq = norm2sq(r, size);

  int k = 0;
  bool converged = sqrt(rsq) < eps * bsqrt;
  if (!converged && max_iter > 0)
  {
    p.push_back(allocate_vector<complex<double> >(size));
    Ap.push_back(allocate_vector<complex<double> >(size));
    copy_vector(p[0], r, size);
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
      print_verbosity_resid(verb, "GCR", k, invif.ops_count, sqrt(rsq) / bsqrt);
      if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
      if (k == max_iter) break;

      zero_vector(Ar, size);
      matrix_vector(Ar, r, extra_info); invif.ops_count++;
      p.push_back(allocate_vector<complex<double> >(size));
      Ap.push_back(allocate_vector<complex<double> >(size));
      copy_vector(p[k], r, size);
      copy_vector(Ap[k], Ar, size);
      for (int i = 0; i < k; i++)
      {
        complex<double> beta = -dot(Ap[i], Ar, size) / ApNormSq[i];
        caxpy(beta, p[i], p[k], size);
        caxpy(beta, Ap[i], Ap[k], size);
      }
    }
  }
  if (k > max_iter) k = max_iter;

  zero_vector(Ar, size);
  matrix_vector(Ar, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(Ar, phi0, size);
  invif.iter = k;
  invif.success = converged;
  print_verbosity_summary(verb, "GCR", invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  for (size_t i = 0; i < p.size(); i++) { deallocate_vector(&p[i]); deallocate_vector(&Ap[i]); }
  deallocate_vector(&r);
  deallocate_vector(&Ar);
  return invif;
}

inline inversion_info minv_vector_gcr_restart(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                              matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  inversion_info invif, total;
  total.name = "Restarted GCR";
  const double bsqrt = sqrt(norm2sq(phi0, size));
  inversion_verbose_struct quiet;
  if (verb != 0) { quiet = *verb; if (quiet.verbosity == VERB_SUMMARY || quiet.verbosity == VERB_RESTART_DETAIL) quiet.verbosity = VERB_NONE; }
  do
  {
    int burst = max_iter - total.iter < restart_freq ? max_iter - total.iter : restart_freq;
    invif = minv_vector_gcr(phi, phi0, size, burst, eps, matrix_vector, extra_info, &quiet);
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
