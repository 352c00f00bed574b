// quantum-mg on B200 -- BiCGstab-L on device vectors (null-vector generation of the K-cycle setup,
// /root/reference/tests/n13_wilson_kcycle/wilson_kcycle.cpp:359: L = 6, 500 iterations, tol 5e-5).
// Sleijpen-Fokkema BiCGstab(L) with the MR part done by modified Gram-Schmidt, as stated by the oracle
// (oracle/qlinalg_shim/inverters/generic_bicgstab_l.h); iter advances by L per outer sweep.
//
// A sweep written call by call moves 280 vectors through HBM beside its 2 L applies (L = 6) -- twice the bytes of the applies
// on the Wilson fine level, which is why the set-up took longer than the solve.  The default path below issues the same
// floating-point operations per element in the same order through three fused kernels (csrc/qmg_blas.cu: qmg_bicgstab_replay,
// _mgs, _finish; 148 vector passes): in the BiCG part only the top vectors r_j, u_j are updated step by step (they feed the
// applies and the dot products), the lower ones and x are brought up to date in ONE pass before the MR part; the modified
// Gram-Schmidt runs right-looking with its coefficients formed on the device (no host wait inside); x, r_0 and |r_0|^2 come
// out of one pass.  QMG_BICGSTAB_FUSED=0 / qmg_set_bicgstab_fused(0) (or L > qmg_bicgstab_max_l()) selects the call-by-call sequence; the two differ only
// by the summation order of the Gram-Schmidt dot products.
#ifndef QMG_B200_BICGSTAB_L
#define QMG_B200_BICGSTAB_L

#include <vector>
#include "../blas/generic_vector.h"
#include "inverter_struct.h"

namespace qmg_host {
inline bool bicgstab_fused(int L) { return qmg_get_bicgstab_fused() != 0 && L <= qmg_bicgstab_max_l(); }
}

inline inversion_info minv_vector_bicgstab_l(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int L,
                                             matrix_op_cplx matrix_vector, void* extra_info, inversion_verbose_struct* verb = 0)
{
  typedef complex<double> cd;
  inversion_info invif;
  invif.name = "BiCGstab-L";
  std::vector<cd*> r(L + 1), u(L + 1);
  for (int i = 0; i <= L; i++)
  {
    r[i] = allocate_vector<cd>(size);
    u[i] = allocate_vector<cd>(size); zero_vector(u[i], size);
  }
  cd* rtilde = allocate_vector<cd>(size);
  const double bsqrt = sqrt(norm2sq(phi0, size));

  matrix_vector(r[0], phi, extra_info); invif.ops_count++;
  caxpby(1.0, phi0, -1.0, r[0], size);
  copy_vector(rtilde, r[0], size);

  cd rho0 = 1.0, alpha = 0.0, omega = 1.0;
  std::vector<cd> gamma(L + 1), gamma_p(L + 1), gamma_pp(L + 1);
  std::vector<double> sigma(L + 1);
  std::vector<std::vector<cd> > tau(L + 1, std::vector<cd>(L + 1));

  const bool fused = qmg_host::bicgstab_fused(L);
  int k = 0;
  double rsq = norm2sq(r[0], size);
  bool converged = sqrt(rsq) < eps * bsqrt;
  while (!converged && k < max_iter)
  {
    rho0 = -omega * rho0;
    // BiCG part
    std::vector<double> al(2 * L), be(2 * L);
    for (int j = 0; j < L; j++)
    {
      const cd rho1 = dot(rtilde, r[j], size);
      const cd beta = alpha * rho1 / rho0;
      rho0 = rho1;
      // fused: only the top vector now, the lower ones (i < j) in the replay pass after the loop
      for (int i = fused ? j : 0; i <= j; i++) caxpby(1.0, r[i], -beta, u[i], size);
      matrix_vector(u[j + 1], u[j], extra_info); invif.ops_count++;
      alpha = rho0 / dot(rtilde, u[j + 1], size);
      for (int i = fused ? j : 0; i <= j; i++) caxpy(-alpha, u[i + 1], r[i], size);
      matrix_vector(r[j + 1], r[j], extra_info); invif.ops_count++;
      if (!fused) caxpy(alpha, u[0], phi, size);
      al[2 * j] = alpha.real(); al[2 * j + 1] = alpha.imag(); be[2 * j] = beta.real(); be[2 * j + 1] = beta.imag();
    }
    if (fused) QMG_CHK(qmg_bicgstab_replay(L, (qmg_cplx* const*)r.data(), (qmg_cplx* const*)u.data(), qmg_host::P(phi), al.data(), be.data(), size));
    // MR part: modified Gram-Schmidt on r[1..L]
    if (fused)
    {
      const int stride = 2 * qmg_bicgstab_max_l() + 2;
      std::vector<double> sums((size_t)L * stride);
      QMG_CHK(qmg_bicgstab_mgs(L, (qmg_cplx* const*)r.data(), size, sums.data()));
      for (int j = 1; j <= L; j++)
      {
        const double* row = &sums[(size_t)(j - 1) * stride];
        sigma[j] = row[0];
        gamma_p[j] = cd(row[1], row[2]) / sigma[j];
        for (int m = j + 1; m <= L; m++) tau[j][m] = cd(row[3 + 2 * (m - j - 1)], row[4 + 2 * (m - j - 1)]) / sigma[j];
      }
    }
    else
    for (int j = 1; j <= L; j++)
    {
      for (int i = 1; i < j; i++)
      {
        tau[i][j] = dot(r[i], r[j], size) / sigma[i];
        caxpy(-tau[i][j], r[i], r[j], size);
      }
      double d[3];
      QMG_CHK(qmg_dot_norm(qmg_host::P(r[j]), qmg_host::P(r[0]), size, d));
      sigma[j] = d[2];
      gamma_p[j] = cd(d[0], d[1]) / sigma[j];
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
    // updates: x += gamma_1 r_0 + sum gamma''_j r_j ; r_0 -= sum gamma'_j r_j ; u_0 -= sum gamma_j u_j
    {
      std::vector<double> cx, cr, cu;
      std::vector<const qmg_cplx*> px, pr, pu;
      cx.push_back(gamma[1].real()); cx.push_back(gamma[1].imag()); px.push_back(qmg_host::P(r[0]));
      for (int j = 1; j < L; j++) { cx.push_back(gamma_pp[j].real()); cx.push_back(gamma_pp[j].imag()); px.push_back(qmg_host::P(r[j])); }
      for (int j = 1; j <= L; j++) { cr.push_back(-gamma_p[j].real()); cr.push_back(-gamma_p[j].imag()); pr.push_back(qmg_host::P(r[j])); }
      if (fused) QMG_CHK(qmg_bicgstab_finish(L, (qmg_cplx* const*)r.data(), qmg_host::P(phi), cx.data(), cr.data(), size, &rsq));      // x, r_0, |r_0|^2 in one pass
      else
      {
        QMG_CHK(qmg_multi_axpy(cx.data(), px.data(), (int)px.size(), qmg_host::P(phi), size));
        QMG_CHK(qmg_multi_axpy(cr.data(), pr.data(), (int)pr.size(), qmg_host::P(r[0]), size));
      }
      for (int j = 1; j <= L; j++) { cu.push_back(-gamma[j].real()); cu.push_back(-gamma[j].imag()); pu.push_back(qmg_host::P(u[j])); }
      QMG_CHK(qmg_multi_axpy(cu.data(), pu.data(), (int)pu.size(), qmg_host::P(u[0]), size));
    }
    k += L;
    if (!fused) rsq = norm2sq(r[0], size);
    qmg_host::say(verb, VERB_DETAIL, "BiCGstab-L", "", false, false, k, invif.ops_count, sqrt(rsq) / bsqrt);
    if (sqrt(rsq) < eps * bsqrt) converged = true;
  }

  matrix_vector(u[0], phi, extra_info); invif.ops_count++;
  invif.resSq = diffnorm2sq(u[0], phi0, size);
  invif.iter = k;
  invif.success = converged;
  qmg_host::say(verb, VERB_SUMMARY, "BiCGstab-L", "", true, invif.success, invif.iter, invif.ops_count, sqrt(invif.resSq) / bsqrt);

  for (int i = 0; i <= L; i++) { deallocate_vector(&r[i]); deallocate_vector(&u[i]); }
  deallocate_vector(&rtilde);
  return invif;
}

#endif
