// quantum-mg on B200 -- GCR on device vectors (coarsest-level solver of the K-cycle,
// /root/reference/multigrid/stateful_multigrid.h:915,947; tests/n18_rbjacobi_stencil_test/rbjacobi_stencil_test.cpp:154).
// Iteration as stated by the oracle (oracle/qlinalg_shim/inverters/generic_gcr.h; quantum-linalg itself is
// un-vendored): untruncated GCR, classical Gram-Schmidt of A r against every stored A p_i.
// Device formulation (qmg_gcr_orthogonalize + qmg_krylov_step): the k projections are ONE multi-dot pass over the stored
// A p_i whose results stay on the device, the coefficients are formed there, both basis updates are ONE pass that also
// forms <Ap_k|r> and |Ap_k|^2, and (alpha, x, r, |r|^2) is one fused update -- one host wait per iteration.
#ifndef QMG_B200_GCR
#define QMG_B200_GCR

#include <vector>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

namespace qmg_host {

// A preconditioner that overwrites its whole output (the K-cycle registers StatefulMultigridMG::mg_preconditioner here)
// needs no zeroed output vector; any other callback still gets one, as quantum-linalg's callers expect.
inline precond_op_cplx& overwriting_precond() { static precond_op_cplx f = 0; return f; }

// Shared body of GCR and flexible (variably preconditioned) GCR.  `hints` (inverter_struct.h) as in minres_core: from a
// zero start r0 = b without applying A to zero, x is written by the first step instead of being zeroed and read, |b|^2
// can come from the caller, and a solve that ends converged skips the true-residual apply nobody reads.
inline inversion_info gcr_core(const char* name, complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                               matrix_op_cplx matrix_vector, void* extra_info,
                               precond_op_cplx precond, void* precond_info, inversion_verbose_struct* verb, SolveHints* hints = 0)
{
  inversion_info invif;
  invif.name = name;
  inversion_verbose_struct verb_prec = precond_view(verb);
  const int flags = (hints != 0 && eps < 1.0 && max_iter > 0) ? hints->flags : 0;
  const bool zero_start = (flags & SOLVE_ZERO_START) != 0;
  const bool zero_for_precond = (precond != 0 && precond != overwriting_precond());
  int executed = 0;
  complex<double>* r = allocate_vector<complex<double> >(size);
  complex<double>* z = precond ? allocate_vector<complex<double> >(size) : 0;
  complex<double>* scratch = 0;
  std::vector<complex<double>*> p, Ap;
  double bsq = (hints != 0 && hints->bnorm2 >= 0.0) ? hints->bnorm2 : norm2sq(phi0, size);
  const double bsqrt = sqrt(bsq);
  complex<double>* r_in = r;
  double rsq;

  if (zero_start)
  {
    invif.ops_count++;           // the reference's A.0
    r_in = phi0; rsq = bsq;
  }
  else
  {
    if (hints != 0 && (hints->flags & SOLVE_ZERO_START)) zero_vector(phi, size);
    scratch = allocate_vector<complex<double> >(size);
    matrix_vector(scratch, phi, extra_info); invif.ops_count++; executed++;
    caxpbyz(1.0, phi0, -1.0, scratch, r, size);
    rsq = norm2sq(r, size);
  }

  int k = 0;
  bool converged = sqrt(rsq) < eps * bsqrt;
  bool stepped = false;
  // device scratch of the orthogonalisation: <Ap_i|Ap_k> of the current step, |Ap_i|^2 of every step so far (the coefficients
  // beta_i = -<Ap_i|Ap_k> / |Ap_i|^2 are formed on the device, qmg_gcr_orthogonalize: one host wait per iteration, not two)
  double* dev_dots = 0; double* dev_apn = 0;
  if (!converged && max_iter > 0)
  {
    long cap = max_iter < 64 ? max_iter + 1 : 64;          // grown on demand: an unrestarted solve may be allowed 10^8 iterations
    dev_dots = allocate_vector<double>(2 * cap);
    dev_apn = allocate_vector<double>(cap);
    p.push_back(allocate_vector<complex<double> >(size));
    Ap.push_back(allocate_vector<complex<double> >(size));
    if (precond) { if (zero_for_precond) zero_vector(p[0], size); precond(p[0], r_in, size, precond_info, &verb_prec); }
    else copy_vector(p[0], r_in, size);
    matrix_vector(Ap[0], p[0], extra_info); invif.ops_count++; executed++;
    bool dots_ready = false;
    for (k = 1; k <= max_iter; k++)
    {
      const int c = k - 1;
      // alpha = <Ap|r> / <Ap|Ap> formed on the device; x += alpha p ; r -= alpha Ap ; |r|^2 : the one host wait of the step
      double step[5];
      QMG_CHK(qmg_krylov_step(1.0, P(p[c]), P(Ap[c]), (zero_start && k == 1) ? 0 : P(phi), P(phi), P(r_in), P(r), 0, size,
                              dots_ready ? QMG_STEP_DOTS_READY : 0, step, dev_apn + c));
      r_in = r;
      stepped = true;
      rsq = step[0];
      say(verb, VERB_DETAIL, name, "", false, false, k, invif.ops_count, sqrt(rsq) / bsqrt);
      if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
      if (k == max_iter) break;

      // next direction: d = M^-1 r (or r), A d straight into its slot, then project out the stored set on the device
      p.push_back(allocate_vector<complex<double> >(size));
      Ap.push_back(allocate_vector<complex<double> >(size));
      complex<double>* dir = r;
      if (precond) { if (zero_for_precond) zero_vector(z, size); precond(z, r, size, precond_info, &verb_prec); dir = z; }
      matrix_vector(Ap[k], dir, extra_info); invif.ops_count++; executed++;
      if (k + 1 > cap)
      {
        double* grown = allocate_vector<double>(2 * cap);
        copy_vector(grown, dev_apn, cap);
        deallocate_vector(&dev_apn); deallocate_vector(&dev_dots);
        cap *= 2;
        dev_apn = grown;
        dev_dots = allocate_vector<double>(2 * cap);
      }
      std::vector<const qmg_cplx*> aps(k), ps(k);
      for (int i = 0; i < k; i++) { aps[i] = P(Ap[i]); ps[i] = P(p[i]); }
      QMG_CHK(qmg_gcr_orthogonalize(aps.data(), ps.data(), k, P(Ap[k]), P(dir), P(p[k]), P(r), size, dev_dots, dev_apn));
      dots_ready = true;
    }
  }
  if (k > max_iter) k = max_iter;
  // a zero start that took no step (b = 0, or no iterations allowed) still owes the caller its x = 0
  if (zero_start && !stepped) zero_vector(phi, size);

  invif.ops_count++;
  if ((flags & SOLVE_NO_FINAL_RESIDUAL) && converged) invif.resSq = rsq;     // converged: nothing downstream reads the true residual
  else
  {
    if (scratch == 0) scratch = allocate_vector<complex<double> >(size);
    matrix_vector(scratch, phi, extra_info); executed++;
    invif.resSq = diffnorm2sq(scratch, phi0, size);
  }
  invif.iter = k;
  invif.success = converged;
  say(verb, VERB_SUMMARY, name, "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);
  if (hints != 0) hints->executed += executed;

  for (size_t i = 0; i < p.size(); i++) { deallocate_vector(&p[i]); deallocate_vector(&Ap[i]); }
  if (dev_dots != 0) deallocate_vector(&dev_dots);
  if (dev_apn != 0) deallocate_vector(&dev_apn);
  deallocate_vector(&r);
  if (scratch != 0) deallocate_vector(&scratch);
  if (z != 0) deallocate_vector(&z);
  return invif;
}

// Bursts of restart_freq iterations of `one_burst` from the current iterate; tolerance stays relative to |b|.
// hints: the zero start only holds for the first burst; |b|^2 is computed once for all of them.
template <class Burst>
inline inversion_info restarted(const char* name, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                inversion_verbose_struct* verb, Burst one_burst, SolveHints* hints = 0)
{
  inversion_info invif, total;
  total.name = name;
  SolveHints local((hints != 0) ? hints->flags : 0, (hints != 0) ? hints->bnorm2 : -1.0);
  if (local.bnorm2 < 0.0) local.bnorm2 = norm2sq(phi0, size);
  const double bsqrt = sqrt(local.bnorm2);
  inversion_verbose_struct quiet = burst_view(verb);
  do
  {
    const int left = max_iter - total.iter;
    invif = one_burst(left < restart_freq ? left : restart_freq, &quiet, &local);
    local.flags &= ~SOLVE_ZERO_START;
    total.iter += invif.iter;
    total.ops_count += invif.ops_count;
    total.resSq = invif.resSq;
    say(verb, VERB_RESTART_DETAIL, name, " Restart", false, false, total.iter, total.ops_count, sqrt(total.resSq) / bsqrt);
  } while (total.iter < max_iter && !invif.success && sqrt(invif.resSq) > eps * bsqrt);
  total.success = invif.success || sqrt(invif.resSq) <= eps * bsqrt;
  say(verb, VERB_SUMMARY, name, "", true, total.success, total.iter, total.ops_count, sqrt(total.resSq) / bsqrt);
  if (hints != 0) hints->executed += local.executed;
  return total;
}

// GCR or flexible GCR, restarted or not (restart_freq == -1), with hints: what the K-cycle calls
inline inversion_info gcr_solve(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                matrix_op_cplx matrix_vector, void* extra_info, precond_op_cplx precond, void* precond_info,
                                inversion_verbose_struct* verb, SolveHints* hints)
{
  const char* nm = precond ? "VPGCR" : "GCR";
  if (restart_freq == -1) return gcr_core(nm, phi, phi0, size, max_iter, eps, matrix_vector, extra_info, precond, precond_info, verb, hints);
  return restarted(precond ? "Restarted VPGCR" : "Restarted GCR", phi0, size, max_iter, eps, restart_freq, verb,
    [&](int burst, inversion_verbose_struct* quiet, SolveHints* h) {
      return gcr_core(nm, phi, phi0, size, burst, eps, matrix_vector, extra_info, precond, precond_info, quiet, h); }, hints);
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
    [&](int burst, inversion_verbose_struct* quiet, qmg_host::SolveHints* h) {
      return qmg_host::gcr_core("GCR", phi, phi0, size, burst, eps, matrix_vector, extra_info, 0, 0, quiet, h); });
}

#endif
