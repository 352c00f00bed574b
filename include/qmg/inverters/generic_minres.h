// quantum-mg on B200 -- relaxed minimal-residual smoother on device vectors.
// Signature from /root/reference/multigrid/stateful_multigrid.h:860:
//   minv_vector_minres(x, b, n, max_iter, rel_tol, omega, op, extra[, verb]).
// quantum-linalg (where the reference takes it from) is un-vendored; the iteration is the one the
// oracle states (oracle/qlinalg_shim/inverters/generic_minres.h) so that iteration counts compare:
//   r = b - A x ; repeat { p = A r ; alpha = <p|r>/<p|p> ; x += w alpha r ; r -= w alpha p } ; true residual.
// Per iteration: one operator apply, one fused (dot, norm) pass and one fused (x, r, |r|^2) update
// -- 3 launches and 1 scalar read-back instead of 1 + 5 BLAS sweeps.
//
// The K-cycle calls the core below with SolveHints (inverter_struct.h): from a zero start the initial residual is the
// right-hand side itself (no A.0 apply, no copy, no zeroing of x: the first step WRITES x), the final true-residual
// apply whose result the K-cycle never reads is skipped, and the last permitted iteration updates x only.  Every value
// that is still computed is formed by the same kernels in the same order, so x is bit-identical to the unhinted call;
// ops_count keeps reporting what the reference would have counted, hints->executed what was launched.
#ifndef QMG_B200_MINRES
#define QMG_B200_MINRES

#include <limits>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

namespace qmg_host {

inline inversion_info minres_core(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, double omega,
                                  matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb, SolveHints* hints)
{
  inversion_info invif;
  invif.name = "MR";
  // the shortcuts assume the solve cannot stop before its first step (a relative tolerance below one) and takes one
  const int flags = (hints != 0 && eps < 1.0 && max_iter > 0) ? hints->flags : 0;
  const bool zero_start = (flags & SOLVE_ZERO_START) != 0;
  int executed = 0;
  // a caller that takes the residual over gets it in its own vector: the recurrence runs there
  complex<double>* rout = (hints != 0 && flags != 0) ? hints->residual_out : 0;
  complex<double>* r = (rout != 0) ? rout : allocate_vector<complex<double> >(size);
  complex<double>* p = allocate_vector<complex<double> >(size);
  double bsq = (hints != 0 && hints->bnorm2 >= 0.0) ? hints->bnorm2 : -1.0;
  const complex<double>* r_in = r;
  double rsq;

  if (zero_start)
  {
    // r = b - A 0 = b exactly; the reference still counts the apply
    invif.ops_count++;
    r_in = phi0;
    rsq = bsq;       // unknown (< 0) until the first step returns |b|^2; the entry test below cannot fire for eps < 1
  }
  else
  {
    // a caller that promised a zero start did not have to zero x: do it here when the shortcut does not apply
    if (hints != 0 && (hints->flags & SOLVE_ZERO_START)) zero_vector(phi, size);
    if (bsq < 0.0) bsq = norm2sq(phi0, size);
    matrix_vector(p, phi, extra_info); invif.ops_count++; executed++;
    caxpbyz(1.0, phi0, -1.0, p, r, size);
    rsq = norm2sq(r, size);
  }
  double bsqrt = bsq >= 0.0 ? sqrt(bsq) : std::numeric_limits<double>::quiet_NaN();

  int k = 0;
  bool converged = false;
  bool acc_done = false;
  complex<double>* acc = (hints != 0) ? hints->accumulate_into : 0;

  // MR(2) from a zero start in two passes (see qmg_mr2_gram): q1 = A r0, p2 = A q1, one pass of dot products, one update.
  // Taken only where the two-step loop below would run both steps anyway (a tolerance the first step cannot meet).
  if (zero_start && max_iter == 2 && (flags & SOLVE_TWO_STEP_MR) && (flags & SOLVE_LAST_X_ONLY) && (flags & SOLVE_NO_FINAL_RESIDUAL))
  {
    typedef complex<double> cplx;
    complex<double>* q1 = p;
    complex<double>* p2 = allocate_vector<complex<double> >(size);
    matrix_vector(q1, const_cast<complex<double>*>(phi0), extra_info);
    matrix_vector(p2, q1, extra_info);
    double g[9];
    QMG_CHK(qmg_mr2_gram(P(phi0), P(q1), P(p2), size, g));
    const cplx a(g[0], g[1]), c(g[3], g[4]), d(g[5], g[6]);
    const double b = g[2], e = g[7], f = g[8];
    const cplx a1 = omega * a / b;
    const double r1sq = f - 2.0 * std::real(std::conj(a1) * a) + std::norm(a1) * b;
    const cplx q2r1 = a - a1 * b - std::conj(a1) * c + std::norm(a1) * d;
    const double q2q2 = b - 2.0 * std::real(std::conj(a1) * d) + std::norm(a1) * e;
    const bool usable = (b > 0.0) && (q2q2 > 0.0) && !(r1sq < eps * eps * f);
    if (usable)
    {
      const cplx a2 = omega * q2r1 / q2q2;
      const cplx cx0 = a1 + a2, cx1 = -a1 * a2, cr1 = -(a1 + a2), cr2 = a1 * a2;
      const double vx0[2] = { cx0.real(), cx0.imag() }, vx1[2] = { cx1.real(), cx1.imag() }, vr1[2] = { cr1.real(), cr1.imag() }, vr2[2] = { cr2.real(), cr2.imag() };
      QMG_CHK(qmg_mr2_update(vx0, vx1, vr1, vr2, P(phi0), P(q1), P(p2), acc != 0 ? P(acc) : 0, P(acc != 0 ? acc : phi), rout != 0 ? P(rout) : 0, size));
      deallocate_vector(&p2);
      invif.ops_count += 3;        // what the reference counts: two iterations and the final residual (A.0 was counted above)
      invif.iter = 2; invif.success = false; invif.resSq = r1sq;
      say(verb, VERB_SUMMARY, "MR", "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq / f));
      hints->executed += 2;
      if (rout != 0) hints->residual_valid = true; else deallocate_vector(&r);
      deallocate_vector(&p);
      return invif;
    }
    deallocate_vector(&p2);      // (a breakdown, or a first step that already meets the tolerance: the step-by-step loop decides)
  }

  if (!zero_start && (max_iter <= 0 || sqrt(rsq) < eps * bsqrt)) converged = (sqrt(rsq) < eps * bsqrt);
  else for (k = 1; k <= max_iter; k++)
  {
    matrix_vector(p, const_cast<complex<double>*>(r_in), extra_info); invif.ops_count++; executed++;
    // alpha = omega <Ar|r> / <Ar|Ar> ; x += alpha r ; r -= alpha A r ; |r|^2 -- alpha is formed on the device, one host wait
    // (x is updated from the old r inside the same thread)
    const bool first = zero_start && k == 1;
    const bool x_only = (flags & SOLVE_LAST_X_ONLY) && k == max_iter;
    const bool want_b = first && bsq < 0.0 && !x_only;
    double step[5];
    // the last step of a smoother: x only -- or, when the caller takes the residual over, x and r without the norm of r
    const int last_flag = x_only ? (rout != 0 ? QMG_STEP_NO_NORM : QMG_STEP_X_ONLY) : 0;
    QMG_CHK(qmg_krylov_step(omega, P(r_in), P(p), first ? 0 : P(phi), P(x_only && acc != 0 ? acc : phi), P(r_in), P(r),
                            x_only && acc != 0 ? P(acc) : 0, size, (want_b ? QMG_STEP_WANT_RNORM : 0) | last_flag, step, 0));
    r_in = r;
    if (x_only) { acc_done = (acc != 0); break; }     // nobody reads this residual's norm: no reduction, no host wait
    if (want_b) { bsq = step[4]; bsqrt = sqrt(bsq); }
    rsq = step[0];
    say(verb, VERB_DETAIL, "MR", "", false, false, k, invif.ops_count, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) { converged = true; break; }
  }
  if (k > max_iter) k = max_iter;
  if (acc != 0 && !acc_done) cxpy(phi, acc, size);

  invif.ops_count++;
  if (flags & SOLVE_NO_FINAL_RESIDUAL) invif.resSq = rsq;      // the recursive residual (of the last step that formed one)
  else
  {
    matrix_vector(p, phi, extra_info); executed++;
    invif.resSq = diffnorm2sq(p, phi0, size);
  }
  invif.iter = k;
  invif.success = converged;
  say(verb, VERB_SUMMARY, "MR", "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);
  if (hints != 0) hints->executed += executed;

  // r holds b - A x of the returned x whenever at least one step ran (every step updates x and r together) or the solve started
  // from an explicit residual; from a zero start without a step x = 0 and the residual is b itself, which was never copied
  if (rout != 0) hints->residual_valid = (r_in == r);
  else deallocate_vector(&r);
  deallocate_vector(&p);
  return invif;
}

} // namespace qmg_host

inline inversion_info minv_vector_minres(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, double omega,
                                         matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  return qmg_host::minres_core(phi, phi0, size, max_iter, eps, omega, matrix_vector, extra_info, verb, 0);
}

#endif
