// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "inverters/generic_bicgstab_l.h".
// Call site: /root/reference/tests/n13_wilson_kcycle/wilson_kcycle.cpp:359
//   minv_vector_bicgstab_l(x, b, n, max_iter, rel_tol, L, op, extra[, verb])
// (null-vector generation only; setup path, not the K-cycle itself).
// Algorithm (UNPINNED, defined here): BiCGstab(L) of Sleijpen & Fokkema,
// ETNA 1 (1993) Alg. 3.1, shadow residual = initial residual, convergence
// tested once per L-step sweep, `iter` counts BiCG steps (L per sweep).
#ifndef QLINALG_SHIM_BICGSTAB_L
#define QLINALG_SHIM_BICGSTAB_L

#include <vector>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

inline inversion_info minv_vector_bicgstab_l(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int L,
                                             matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  typedef complex<double> cd;
  inversion_info invif;
  invif.name = "BiCGstab-L";
  std::vector<cd*> r(L + 1), u(L + 1);
  for (int i = 0; i <= L; i++)
  {
    r[i] = allocate_vector<cd>(size); zero_vector(r[i], size);
    u[i] = allocate_vector<cd>(size); zero_vector(u[i], size);
  }
  cd* rtilde = allocate_vector<cd>(size);
  const double bsqrt = sqrt(norm2sq(phi0, size));

  matrix_vector(r[0], phi, extra_info); invif.ops_count++;   // r[0] was zeroed above
  caxpby(1.0, phi0, -1.0, r[0], size);
  copy_vector(rtilde, r[0], size);

  cd rho0 = 1.0, alpha = 0.0, omega = 1.0;
  std::vector<cd> gamma(L + 1), gamma_p(L + 1), gamma_pp(L + 1), sigma(L + 1);
  std::vector<std::vector<cd> > tau(L + 1, std::vector<cd>(L + 1));

  int k = 0;
  double rsq = norm2sq(r[0], size);
  bool converged = sqrt(rsq) < eps * bsqrt;
  while (!converged && k < max_iter)
  {
    rho0 = -omega * rho0;
    for (int j = 0; j < L; j++)
    {
      cd rho1 = dot(rtilde, r[j], size);
      cd beta = alpha * rho1 / rho0;
      rho0 = rho1;
      for (int i = 0; i <= j; i++) caxpby(1.0, r[i], -beta, u[i], size);
      zero_vector(u[j + 1], size);
      matrix_vector(u[j + 1], u[j], extra_info); invif.ops_count++;
      alpha = rho0 / dot(rtilde, u[j + 1], size);
      for (int i = 0; i <= j; i++) caxpy(-alpha, u[i + 1], r[i], size);
      zero_vector(r[j + 1], size);
      matrix_vector(r[j + 1], r[j], extra_info); invif.ops_count++;
      caxpy(alpha, u[0], phi, size);
    }
    for (int j = 1; j <= L; j++)
    {
      for (int i = 1; i < j; i++)
      {
        tau[i][j] = dot(r[i], r[j], size) / sigma[i];
        caxpy(-tau[i][j], r[i], r[j], size);
      }
      sigma[j] = norm2sq(r[j], size);
      gamma_p[j] = dot(r[j], r[0], size) / sigma[j];
    }
    gamma[L] = gamma_p[L];
    omega = gamma[L];
    for (int j = L - 1; j >= 1; j--)
    {
      gamma[j] = gamma_p[j];
      for (int i = j + 1; i <= L; i++) gamma[j] -= tau[j][i] * gamma[i];
    }
    for (int j = 1; j < L; j++)
    {
      gamma_pp[j] = gamma[j + 1];
      for (int i = j + 1; i < L; i++) gamma_pp[j] += tau[j][i] * gamma[i + 1];
    }
    caxpy(gamma[1], r[0], phi, size);
    caxpy(-gamma_p[L], r[L], r[0], size);
    caxpy(-gamma[L], u[L], u[0], size);
    for (int j = 1; j < L; j++)
    {
      caxpy(-gamma[j], u[j], u[0], size);
      caxpy(gamma_pp[j], r[j], phi, size);
      caxpy(-gamma_p[j], r[j], r[0], size);
    }
    k += L;
    rsq = norm2sq(r[0], size);
    print_verbosity_resid(verb, "BiCGstab-L", k, invif.ops_count, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) converged = true;
  }

  zero_vector(u[0], size);
  matrix_vector(u[0], phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(u[0], phi0, size);
  invif.iter = k;
  invif.success = converged;
  print_verbosity_summary(verb, "BiCGstab-L", invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  for (int i = 0; i <= L; i++) { deallocate_vector(&r[i]); deallocate_vector(&u[i]); }
  deallocate_vector(&rtilde);
  return invif;
}

#endif
