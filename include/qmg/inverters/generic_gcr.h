// quantum-mg on B200 -- GCR on device vectors (coarsest-level solver of the K-cycle,
// /root/reference/multigrid/stateful_multigrid.h:915,947; tests/n18_rbjacobi_stencil_test/rbjacobi_stencil_test.cpp:154).
// Iteration as stated by the oracle (oracle/qlinalg_shim/inverters/generic_gcr.h; quantum-linalg itself is
// un-vendored): untruncated GCR, classical Gram-Schmidt of A r against every stored A p_i.
// Device formulation: the k projections are ONE multi-dot pass over the stored A p_i, the two basis updates are
// two multi-axpy passes, and (alpha, x, r, |r|^2) is one fused (dot, norm) pass plus one fused update.
#ifndef QMG_B200_GCR
#define QMG_B200_GCR

#include <vector>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

namespace qmg_host {

// Shared body of GCR and flexible (variably preconditioned) GCR.
inline inversion_info gcr_core(const char* name, complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                               matrix_op_cplx matrix_vector, void* extra_info,
                               precond_op_cplx precond, void* precond_info, inversion_verbose_struct* verb)
{
  inversion_info invif;
  invif.name = name;
  inversion_verbose_struct verb_prec = precond_view(verb);
  complex<double>* r = allocate_vector<complex<double> >(size);
  complex<double>* z = precond ? allocate_vector<complex<double> >(size) : 0;
  complex<double>* scratch = allocate_vector<complex<double> >(size);
  std::vector<complex<double>*> p, Ap;
  std::vector<double> ApNormSq;
  const double bsqrt = sqrt(norm2sq(phi0, size));

  matrix_vector(scratch, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, scratch, r, size);
  double rsq = norm2sq(r, size);

  int k = 0;
  bool converged = sqrt(rsq) < eps * bsqrt;
  if (!converged && max_iter > 0)
  {
    p.push_back(allocate_vector<complex<double> >(size));
    Ap.push_back(allocate_vector<complex<double> >(size));
    if (precond) { zero_vector(p[0], size); precond(p[0], r, size, precond_info, &verb_prec); }
    else copy_vector(p[0], r, size);
    matrix_vector(Ap[0], p[0], extra_info); invif.ops_count++;
    for (k = 1; k <= max_iter; k++)
    {
      const int c = k - 1;
      // alpha = <Ap|r> / <Ap|Ap> formed on the device; x += alpha p ; r -= alpha Ap ; |r|^2 : one host wait per step
      double step[4];
      QMG_CHK(qmg_step_xr_norm(1.0, P(p[c]), P(Ap[c]), P(phi), P(r), size, step));
      rsq = step[0];
      ApNormSq.push_back(step[3]);
      say(verb, VERB_DETAIL, name, "", false, false, k, invif.ops_count, sqrt(rsq) / bsqrt);
      if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
      if (k == max_iter) break;

      // next direction: d = M^-1 r (or r), A d straight into its slot, then project out the stored set
      p.push_back(allocate_vector<complex<double> >(size));
      Ap.push_back(allocate_vector<complex<double> >(size));
      complex<double>* dir = r;
      if (precond) { zero_vector(z, size); precond(z, r, size, precond_info, &verb_prec); dir = z; }
      matrix_vector(Ap[k], dir, extra_info); invif.ops_count++;
      std::vector<double> beta(2 * k);
      std::vector<const qmg_cplx*> ptrs(k);
      for (int i = 0; i < k; i++) ptrs[i] = P(Ap[i]);
      QMG_CHK(qmg_multi_dot(ptrs.data(), k, P(Ap[k]), size, beta.data()));
      for (int i = 0; i < k; i++) { beta[2 * i] = -beta[2 * i] / ApNormSq[i]; beta[2 * i + 1] = -beta[2 * i + 1] / ApNormSq[i]; }
      QMG_CHK(qmg_multi_axpyz(beta.data(), ptrs.data(), k, P(Ap[k]), P(Ap[k]), size));
      for (int i = 0; i < k; i++) ptrs[i] = P(p[i]);
      QMG_CHK(qmg_multi_axpyz(beta.data(), ptrs.data(), k, P(dir), P(p[k]), size));
    }
  }
  if (k > max_iter) k = max_iter;

  matrix_vector(scratch, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(scratch, phi0, size);
  invif.iter = k;
  invif.success = converged;
  say(verb, VERB_SUMMARY, name, "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  for (size_t i = 0; i < p.size(); i++) { deallocate_vector(&p[i]); deallocate_vector(&Ap[i]); }
  deallocate_vector(&r);
  deallocate_vector(&scratch);
  if (z != 0) deallocate_vector(&z);
  return invif;
}

// Bursts of restart_freq iterations of `one_burst` from the current iterate; tolerance stays relative to |b|.
template <class Burst>
inline inversion_info restarted(const char* name, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                inversion_verbose_struct* verb, Burst one_burst)
{
  inversion_info invif, total;
  total.name = name;
  const double bsqrt = sqrt(norm2sq(phi0, size));
  inversion_verbose_struct quiet = burst_view(verb);
  do
  {
    const int left = max_iter - total.iter;
    invif = one_burst(left < restart_freq ? left : restart_freq, &quiet);
    total.iter += invif.iter;
    total.ops_count += invif.ops_count;
    total.resSq = invif.resSq;
    say(verb, VERB_RESTART_DETAIL, name, " Restart", false, false, total.iter, total.ops_count, sqrt(total.resSq) / bsqrt);
  } while (total.iter < max_iter && !invif.success && sqrt(invif.resSq) > eps * bsqrt);
  total.success = invif.success || sqrt(invif.resSq) <= eps * bsqrt;
  say(verb, VERB_SUMMARY, name, "", true, total.success, total.iter, total.ops_count, sqrt(total.resSq) / bsqrt);
  return total;
}

} // namespace qmg_host

inline inversion_info minv_vector_gcr(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                      matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  return qmg_host::gcr_core("GCR", phi, phi0, size, max_iter, eps, matrix_vector, extra_info, 0, 0, verb);
}

inline inversion_info minv_vector_gcr_restart(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                              matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  return qmg_host::restarted("Restarted GCR", phi0, size, max_iter, eps, restart_freq, verb,
    [&](int burst, inversion_verbose_struct* quiet) { return minv_vector_gcr(phi, phi0, size, burst, eps, matrix_vector, extra_info, quiet); });
}

#endif
