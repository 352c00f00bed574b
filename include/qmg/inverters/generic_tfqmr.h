// quantum-mg on B200 -- transpose-free QMR (Freund 1993, the weighted-free variant) on device vectors
// (quantum-linalg "inverters/generic_tfqmr.h"; the solver survey of /root/reference/tests/n11_wilson_test/wilson_test.cpp:241).
// Not on the hot path; quantum-linalg is un-vendored and the oracle's shim only declares this solver, so it is checked by
// its explicit residual, not by iteration-count parity.
#ifndef QMG_B200_TFQMR
#define QMG_B200_TFQMR
#include <cmath>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

inline inversion_info minv_vector_tfqmr(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                        matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  typedef complex<double> cd;
  inversion_info invif;
  invif.name = "TFQMR";
  cd* r0 = allocate_vector<cd>(size); cd* w = allocate_vector<cd>(size); cd* y = allocate_vector<cd>(size);
  cd* y_next = allocate_vector<cd>(size); cd* v = allocate_vector<cd>(size); cd* d = allocate_vector<cd>(size);
  cd* Ay = allocate_vector<cd>(size); cd* Ay_next = allocate_vector<cd>(size);
  const double bsqrt = sqrt(norm2sq(phi0, size));

  matrix_vector(Ay, phi, extra_info); invif.ops_count++;
  caxpbyz(1.0, phi0, -1.0, Ay, r0, size);                 // r0 = b - A x
  copy_vector(w, r0, size); copy_vector(y, r0, size);
  matrix_vector(Ay, y, extra_info); invif.ops_count++;
  copy_vector(v, Ay, size);
  zero_vector(d, size);
  double tau = sqrt(norm2sq(r0, size)), theta = 0.0, eta_re = 0.0;
  cd eta = 0.0, rho = dot(r0, r0, size);
  (void)eta_re;
  int k = 0;
  bool converged = tau < eps * bsqrt;
  while (!converged && k < max_iter)
  {
    const cd sigma = dot(r0, v, size);
    const cd alpha = rho / sigma;
    caxpbyz(1.0, y, -alpha, v, y_next, size);              // y_{2k} = y_{2k-1} - alpha v
    matrix_vector(Ay_next, y_next, extra_info); invif.ops_count++;
    for (int half = 0; half < 2 && !converged; half++)
    {
      cd* yy = half == 0 ? y : y_next; cd* Ayy = half == 0 ? Ay : Ay_next;
      caxpy(-alpha, Ayy, w, size);                         // w -= alpha A y
      // d = y + (theta^2 eta / alpha) d
      caxpby(1.0, yy, theta * theta * eta / alpha, d, size);
      theta = sqrt(norm2sq(w, size)) / tau;
      const double c = 1.0 / sqrt(1.0 + theta * theta);
      tau = tau * theta * c;
      eta = c * c * alpha;
      caxpy(eta, d, phi, size);                            // x += eta d
      k++;
      qmg_host::say(verb, VERB_DETAIL, "TFQMR", "", false, false, k, invif.ops_count, tau * sqrt((double)(k + 1)) / bsqrt);
      // tau sqrt(m + 1) bounds the true residual norm: only check it properly when the bound says so
      if (tau * sqrt((double)(k + 1)) < eps * bsqrt)
      {
        matrix_vector(Ay, phi, extra_info); invif.ops_count++;
        if (sqrt(diffnorm2sq(Ay, phi0, size)) < eps * bsqrt) converged = true;
        else if (half == 0) { matrix_vector(Ay, y, extra_info); invif.ops_count++; }     // Ay was borrowed: restore A y
      }
      if (k >= max_iter) break;
    }
    if (converged || k >= max_iter) break;
    const cd rho_next = dot(r0, w, size);
    const cd beta = rho_next / rho;
    rho = rho_next;
    // y_{2k+1} = w + beta y_{2k};  v = A y_{2k+1} + beta (A y_{2k} + beta v)
    caxpbyz(1.0, w, beta, y_next, y, size);
    matrix_vector(Ay, y, extra_info); invif.ops_count++;
    caxpby(1.0, Ay_next, beta, v, size);                   // v = A y_{2k} + beta v
    caxpby(1.0, Ay, beta, v, size);                        // v = A y_{2k+1} + beta (...)
  }
  matrix_vector(Ay, phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(Ay, phi0, size);
  invif.iter = k;
  invif.success = sqrt(invif.resSq) < eps * bsqrt;
  qmg_host::say(verb, VERB_SUMMARY, "TFQMR", "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);
  deallocate_vector(&r0); deallocate_vector(&w); deallocate_vector(&y); deallocate_vector(&y_next);
  deallocate_vector(&v); deallocate_vector(&d); deallocate_vector(&Ay); deallocate_vector(&Ay_next);
  return invif;
}
#endif
