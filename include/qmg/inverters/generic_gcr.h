// quantum-mg on B200 -- GCR on device vectors (coarsest-level solver of the K-cycle,
// /root/reference/multigrid/stateful_multigrid.h:915,947; tests/n18_rbjacobi_stencil_test/rbjacobi_stencil_test.cpp:154).
// Iteration as stated by the oracle (oracle/qlinalg_shim/inverters/generic_gcr.h; quantum-linalg itself is
// un-vendored): untruncated GCR, classical Gram-Schmidt of A r against every stored A p_i.
// Device formulation (qmg_gcr_orthogonalize + qmg_krylov_step): the k projections are ONE multi-dot pass over the stored
// A p_i whose results stay on the device, the coefficients are formed there, the A p basis update also forms <Ap_k|r> and
// |Ap_k|^2, and (alpha, r, |r|^2) is one fused update -- one host wait per iteration; x is formed once per solve.
#ifndef QMG_B200_GCR
#define QMG_B200_GCR

#include <vector>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

namespace qmg_host {

// A preconditioner that overwrites its whole output (the K-cycle registers StatefulMultigridMG::mg_preconditioner here)
// needs no zeroed output vector; any other callback still gets one, as quantum-linalg's callers expect.
inline precond_op_cplx& overwriting_precond() { static precond_op_cplx f = 0; return f; }

// A preconditioner that knows A z of the z it returns (the K-cycle does: its post-smoother's recurrence residual r' satisfies
// A z = rhs - r') can spare the flexible solver the operator apply that follows every preconditioner call.  The solver posts
// a request -- where A z should go, for which operator -- before the call; a callback that can serve it fills `out` and sets
// `valid`, any other ignores it and the solver applies the operator itself.  Requests nest with the solves: whoever posts
// one restores the previous one after the call, and the K-cycle takes the pending request off the board before it starts
// the solves one level down.
struct PrecondAzRequest { complex<double>* out; matrix_op_cplx op; void* op_data; bool valid; };
inline PrecondAzRequest*& precond_az_request() { static PrecondAzRequest* r = 0; return r; }

// Shared body of GCR and flexible (variably preconditioned) GCR.  `hints` (inverter_struct.h) as in minres_core: from a
// zero start r0 = b without applying A to zero, |b|^2 can come from the caller, and a solve that ends converged skips the
// true-residual apply nobody reads.
//
// Storage is GMRES-style: the raw directions d_k (the residuals themselves, or the preconditioner's outputs) are kept, not
// their orthogonalised combinations p_k = d_k + sum_i beta_ki p_i.  Per iteration only the A p basis is orthogonalised and the
// residual recurrence r -= alpha_k A p_k runs -- every convergence decision is taken on the same numbers as in textbook GCR,
// bit for bit -- and x = x0 + sum_k alpha_k p_k = x0 + sum_j c_j d_j is formed ONCE at the end, c = U alpha with the small
// unit upper-triangular U the beta's define (host, K^3 flops, K <= restart length).  That removes the p-basis update (k + 2
// vector passes) and the x update (3 passes) from every iteration; the unpreconditioned solver does not even copy the
// residual into a direction: each step writes its new residual into a fresh vector and the old one IS d_k.
inline inversion_info gcr_core(const char* name, complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                               matrix_op_cplx matrix_vector, void* extra_info,
                               precond_op_cplx precond, void* precond_info, inversion_verbose_struct* verb, SolveHints* hints = 0)
{
  typedef complex<double> cplx;
  inversion_info invif;
  invif.name = name;
  inversion_verbose_struct verb_prec = precond_view(verb);
  const int flags = (hints != 0 && eps < 1.0 && max_iter > 0) ? hints->flags : 0;
  const bool zero_start = (flags & SOLVE_ZERO_START) != 0;
  const bool zero_for_precond = (precond != 0 && precond != overwriting_precond());
  int executed = 0;
  std::vector<cplx*> owned;                 // every vector allocated here
  std::vector<cplx*> d, Ap;                 // raw directions (may alias residual vectors or phi0) and the orthogonalised A p basis
  std::vector<cplx> alpha;
  std::vector<double> apn;
  struct Alloc { std::vector<cplx*>& o; int n; cplx* operator()() { cplx* v = allocate_vector<cplx>(n); o.push_back(v); return v; } } fresh = { owned, size };
  // d = M^-1 r and A d: from the preconditioner itself when it can say (see PrecondAzRequest), else by an apply
  auto precond_then_apply = [&](cplx* dk, cplx* r, cplx* Adk)
  {
    PrecondAzRequest req = { Adk, matrix_vector, extra_info, false };
    PrecondAzRequest* pending = precond_az_request();
    precond_az_request() = &req;
    precond(dk, r, size, precond_info, &verb_prec);
    precond_az_request() = pending;
    invif.ops_count++;
    if (!req.valid) { matrix_vector(Adk, dk, extra_info); executed++; }
  };
  cplx* scratch = 0;
  double bsq = (hints != 0 && hints->bnorm2 >= 0.0) ? hints->bnorm2 : norm2sq(phi0, size);
  const double bsqrt = sqrt(bsq);
  cplx* cur_r;
  double rsq;

  if (zero_start)
  {
    invif.ops_count++;           // the reference's A.0
    cur_r = phi0; rsq = bsq;
  }
  else
  {
    if (hints != 0 && (hints->flags & SOLVE_ZERO_START)) zero_vector(phi, size);
    scratch = fresh();
    cur_r = fresh();
    matrix_vector(scratch, phi, extra_info); invif.ops_count++; executed++;
    caxpbyz(1.0, phi0, -1.0, scratch, cur_r, size);
    rsq = norm2sq(cur_r, size);
  }

  int k = 0;
  bool converged = sqrt(rsq) < eps * bsqrt;
  // device scratch: row k holds <Ap_i|Ap_k>, i < k, of the orthogonalisation of direction k; |Ap_i|^2 of every step so far
  double* dev_dots = 0; double* dev_apn = 0;
  long cap = 0;
  if (!converged && max_iter > 0)
  {
    cap = max_iter < 32 ? max_iter + 1 : 32;          // grown on demand: an unrestarted solve may be allowed 10^8 iterations
    dev_dots = allocate_vector<double>(2 * cap * cap);
    dev_apn = allocate_vector<double>(cap);
    cplx* dk = cur_r;
    Ap.push_back(fresh());
    if (precond) { dk = fresh(); if (zero_for_precond) zero_vector(dk, size); precond_then_apply(dk, cur_r, Ap[0]); }
    else { matrix_vector(Ap[0], dk, extra_info); invif.ops_count++; executed++; }
    d.push_back(dk);
    bool dots_ready = false;
    for (k = 1; k <= max_iter; k++)
    {
      const int c = k - 1;
      // alpha = <Ap|r> / <Ap|Ap> formed on the device; r' = r - alpha Ap ; |r'|^2 : the one host wait of the iteration.
      // Unpreconditioned, r' goes into a fresh vector: the old residual is direction d_c and must stay.
      cplx* next_r = precond ? (cur_r == phi0 ? fresh() : cur_r) : fresh();
      double step[5];
      QMG_CHK(qmg_krylov_step(1.0, 0, P(Ap[c]), 0, 0, P(cur_r), P(next_r), 0, size, QMG_STEP_R_ONLY | (dots_ready ? QMG_STEP_DOTS_READY : 0), step, dev_apn + c));
      alpha.push_back(cplx(step[1], step[2]) / step[3]);
      apn.push_back(step[3]);
      cur_r = next_r;
      rsq = step[0];
      say(verb, VERB_DETAIL, name, "", false, false, k, invif.ops_count, sqrt(rsq) / bsqrt);
      if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
      if (k == max_iter) break;

      // next direction: d = M^-1 r (or r itself), A d straight into its slot, then project the stored A p_i out of it on the device
      dk = cur_r;
      Ap.push_back(fresh());
      if (precond) { dk = fresh(); if (zero_for_precond) zero_vector(dk, size); precond_then_apply(dk, cur_r, Ap[k]); }
      else { matrix_vector(Ap[k], dk, extra_info); invif.ops_count++; executed++; }
      d.push_back(dk);
      if (k + 1 > cap)
      {
        double* gd = allocate_vector<double>(8 * cap * cap);        // (2 cap)^2 rows x columns, 2 doubles each
        double* ga = allocate_vector<double>(2 * cap);
        for (long row = 1; row < k; row++) copy_vector(gd + row * 4 * cap, dev_dots + row * 2 * cap, 2 * row);
        copy_vector(ga, dev_apn, cap);
        deallocate_vector(&dev_dots); deallocate_vector(&dev_apn);
        dev_dots = gd; dev_apn = ga; cap *= 2;
      }
      std::vector<const qmg_cplx*> aps(k);
      for (int i = 0; i < k; i++) aps[i] = P(Ap[i]);
      QMG_CHK(qmg_gcr_orthogonalize(aps.data(), 0, k, P(Ap[k]), 0, 0, P(cur_r), size, dev_dots + (long)k * 2 * cap, dev_apn));
      dots_ready = true;
    }
  }
  if (k > max_iter) k = max_iter;

  // x = x0 + sum_j c_j d_j,  c = U alpha,  U[:,k] = e_k + sum_{i<k} beta_ki U[:,i],  beta_ki = -<Ap_i|Ap_k> / |Ap_i|^2
  const int K = (int)alpha.size();
  if (zero_start) zero_vector(phi, size);      // (also the zero start that took no step: b = 0)
  if (K > 0)
  {
    std::vector<double> hd((size_t)K * 2 * cap, 0.0);
    if (K > 1) QMG_CHK(qmg_memcpy_d2h(hd.data(), dev_dots, sizeof(double) * (size_t)K * 2 * cap));
    std::vector<std::vector<cplx> > U(K, std::vector<cplx>(K, cplx(0.0, 0.0)));
    for (int kk = 0; kk < K; kk++)
    {
      U[kk][kk] = 1.0;
      for (int i = 0; i < kk; i++)
      {
        const cplx beta = -cplx(hd[(size_t)kk * 2 * cap + 2 * i], hd[(size_t)kk * 2 * cap + 2 * i + 1]) / apn[i];
        for (int j = 0; j <= i; j++) U[j][kk] += beta * U[j][i];
      }
    }
    std::vector<double> coef(2 * K);
    std::vector<const qmg_cplx*> ds(K);
    for (int j = 0; j < K; j++)
    {
      cplx cj(0.0, 0.0);
      for (int kk = j; kk < K; kk++) cj += U[j][kk] * alpha[kk];
      coef[2 * j] = cj.real(); coef[2 * j + 1] = cj.imag();
      ds[j] = P(d[j]);
    }
    QMG_CHK(qmg_multi_axpy(coef.data(), ds.data(), K, P(phi), size));
  }

  invif.ops_count++;
  if ((flags & SOLVE_NO_FINAL_RESIDUAL) && converged) invif.resSq = rsq;     // converged: nothing downstream reads the true residual
  else
  {
    if (scratch == 0) scratch = fresh();
    matrix_vector(scratch, phi, extra_info); executed++;
    invif.resSq = diffnorm2sq(scratch, phi0, size);
  }
  invif.iter = k;
  invif.success = converged;
  say(verb, VERB_SUMMARY, name, "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);
  if (hints != 0) hints->executed += executed;

  for (size_t i = 0; i < owned.size(); i++) deallocate_vector(&owned[i]);
  if (dev_dots != 0) deallocate_vector(&dev_dots);
  if (dev_apn != 0) deallocate_vector(&dev_apn);
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
