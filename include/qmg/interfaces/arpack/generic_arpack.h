// quantum-mg on B200 -- the eigensolver interface of quantum-linalg ("interfaces/arpack/generic_arpack.h", class
// arpack_dcn) for the one use the hot path has for it: deflating the coarsest-level NORMAL-equation solve
// (/root/reference/multigrid/stateful_multigrid.h:613-696 gets num_low / num_high eigenpairs of M^dag M with
// ncv = 3 nev, tol 1e-5).  ARPACK and gfortran are not in this image, so the class is restated for HERMITIAN operators
// as a thick-restart Lanczos process on device vectors:
//   * the Krylov basis lives in HBM; every new vector is orthogonalised against the whole basis with ONE fused
//     multi-dot and ONE multi-axpy pass (twice: classical Gram-Schmidt with re-orthogonalisation);
//   * the projected ncv x ncv matrix is diagonalised on the host (cyclic Jacobi for complex Hermitian matrices);
//   * the wanted Ritz vectors plus a few neighbours are kept across restarts (thick restart), convergence is the
//     Lanczos residual estimate |beta_m s_i(m)| <= tol |theta_i|.
// A non-Hermitian operator (the Wilson operator itself: the spectra printed by tests/n10, n12, n13 :485) is refused
// with an error -- those need an Arnoldi process, which nothing on the path calls.
#ifndef QMG_B200_ARPACK
#define QMG_B200_ARPACK

#include <algorithm>
#include <cmath>
#include <complex>
#include <iostream>
#include <random>
#include <vector>

#include "../../blas/generic_vector.h"
#include "../../inverters/inverter_struct.h"

namespace qmg_host {

// Eigen-decomposition of a small complex Hermitian matrix A (m x m, row major) by cyclic Jacobi rotations:
// on return evals ascending, evecs column j (evecs[i * m + j]) the eigenvector of evals[j].
inline void hermitian_eig(std::vector<std::complex<double> > A, int m, std::vector<double>& evals, std::vector<std::complex<double> >& evecs)
{
  typedef std::complex<double> cd;
  std::vector<cd> V((size_t)m * m, 0.0);
  for (int i = 0; i < m; i++) V[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 60; sweep++)
  {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) (i == j ? diag : off) += std::norm(A[(size_t)i * m + j]);
    if (off <= 1e-30 * (diag > 0 ? diag : 1.0)) break;
    for (int p = 0; p < m - 1; p++)
      for (int q = p + 1; q < m; q++)
      {
        const cd apq = A[(size_t)p * m + q];
        const double g = std::abs(apq);
        if (g < 1e-300) continue;
        const double app = A[(size_t)p * m + p].real(), aqq = A[(size_t)q * m + q].real();
        // rotate the (p, q) plane: phase first (makes the off-diagonal real), then a real Jacobi rotation
        const cd phase = apq / g;
        const double tau = (aqq - app) / (2.0 * g);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        // columns p, q of A and V:  [p', q'] = [p, q] * [[c, s phase], [-s conj(phase), c]]
        for (int k = 0; k < m; k++)
        {
          const cd akp = A[(size_t)k * m + p], akq = A[(size_t)k * m + q];
          A[(size_t)k * m + p] = c * akp - s * std::conj(phase) * akq;
          A[(size_t)k * m + q] = s * phase * akp + c * akq;
          const cd vkp = V[(size_t)k * m + p], vkq = V[(size_t)k * m + q];
          V[(size_t)k * m + p] = c * vkp - s * std::conj(phase) * vkq;
          V[(size_t)k * m + q] = s * phase * vkp + c * vkq;
        }
        // rows p, q of A with the adjoint rotation
        for (int k = 0; k < m; k++)
        {
          const cd apk = A[(size_t)p * m + k], aqk = A[(size_t)q * m + k];
          A[(size_t)p * m + k] = c * apk - s * phase * aqk;
          A[(size_t)q * m + k] = s * std::conj(phase) * apk + c * aqk;
        }
      }
  }
  std::vector<int> order(m);
  for (int i = 0; i < m; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * m + a].real() < A[(size_t)b * m + b].real(); });
  evals.resize(m); evecs.assign((size_t)m * m, 0.0);
  for (int j = 0; j < m; j++)
  {
    evals[j] = A[(size_t)order[j] * m + order[j]].real();
    for (int i = 0; i < m; i++) evecs[(size_t)i * m + j] = V[(size_t)i * m + order[j]];
  }
}

} // namespace qmg_host

class arpack_dcn
{
public:
  enum arpack_spectrum_piece { ARPACK_NONE, ARPACK_LARGEST_MAGNITUDE, ARPACK_SMALLEST_MAGNITUDE, ARPACK_LARGEST_REAL, ARPACK_SMALLEST_REAL, ARPACK_LARGEST_IMAGINARY, ARPACK_SMALLEST_IMAGINARY };

  arpack_dcn(int n, int maxitr, double tol, matrix_op_cplx op, void* extra) : n(n), maxitr(maxitr), tol(tol), op(op), extra(extra), nev(0), ncv(0), ready(false), restarts(0), ops(0) { }
  arpack_dcn(int n, int maxitr, double tol, matrix_op_cplx op, void* extra, int nev, int ncv) : n(n), maxitr(maxitr), tol(tol), op(op), extra(extra), nev(nev), ncv(ncv), ready(false), restarts(0), ops(0) { }
  ~arpack_dcn() { release(); }

  // compute nev eigenpairs at the requested end of the spectrum of a HERMITIAN operator (kept in the object)
  bool prepare_eigensystem(arpack_spectrum_piece piece, int nev_in, int ncv_in)
  {
    typedef std::complex<double> cd;
    release();
    nev = nev_in; ncv = ncv_in;
    const bool want_low = (piece == ARPACK_SMALLEST_REAL || piece == ARPACK_SMALLEST_MAGNITUDE);
    if (piece == ARPACK_NONE || piece == ARPACK_LARGEST_IMAGINARY || piece == ARPACK_SMALLEST_IMAGINARY)
    { std::cout << "[QMG-ERROR]: arpack_dcn on B200 serves Hermitian operators: only the real ends of the spectrum exist.\n"; return false; }
    if (nev < 1 || ncv < nev + 2 || ncv > n) { if (ncv > n) ncv = n; if (nev < 1 || ncv < nev + 1) { std::cout << "[QMG-ERROR]: arpack_dcn needs 1 <= nev and nev + 2 <= ncv <= n.\n"; return false; } }
    const int m = ncv;
    std::vector<cd*> V(m + 1);
    for (int i = 0; i <= m; i++) V[i] = allocate_vector<cd>(n);
    cd* w = allocate_vector<cd>(n);
    auto cleanup = [&]() { for (int i = 0; i <= m; i++) deallocate_vector(&V[i]); deallocate_vector(&w); };

    // Hermitian?  <y|A x> must equal conj(<x|A y>) for random x, y
    std::mt19937 gen(20171337u);
    gaussian(V[0], n, gen); gaussian(V[1], n, gen);
    op(w, V[0], extra); const cd yAx = dot(V[1], w, n);
    op(w, V[1], extra); const cd xAy = dot(V[0], w, n);
    ops += 2;
    if (std::abs(yAx - std::conj(xAy)) > 1e-8 * (std::abs(yAx) + std::abs(xAy) + 1e-300))
    {
      std::cout << "[QMG-ERROR]: arpack_dcn on B200 is a Lanczos process and needs a Hermitian operator (normal-equation stencils); the operator given is not.\n";
      cleanup(); return false;
    }

    std::vector<cd> T((size_t)m * m, 0.0);
    normalize(V[0], n);
    int k = 0;                      // basis vectors 0..k-1 are locked Ritz vectors, V[k] is the current Lanczos vector
    std::vector<double> theta; std::vector<cd> S;
    bool done = false;
    for (restarts = 0; restarts < maxitr && !done; restarts++)
    {
      double beta_m = 0.0;
      for (int j = k; j < m; j++)
      {
        op(w, V[j], extra); ops++;
        // project out the whole basis (this also produces column j of T), twice for numerical orthogonality
        std::vector<double> c(2 * (j + 1)), c2(2 * (j + 1));
        std::vector<const qmg_cplx*> ptrs(j + 1);
        for (int i = 0; i <= j; i++) ptrs[i] = qmg_host::P(V[i]);
        QMG_CHK(qmg_multi_dot(ptrs.data(), j + 1, qmg_host::P(w), n, c.data()));
        for (int i = 0; i < 2 * (j + 1); i++) c2[i] = -c[i];
        QMG_CHK(qmg_multi_axpy(c2.data(), ptrs.data(), j + 1, qmg_host::P(w), n));
        std::vector<double> d(2 * (j + 1));
        QMG_CHK(qmg_multi_dot(ptrs.data(), j + 1, qmg_host::P(w), n, d.data()));
        for (int i = 0; i < 2 * (j + 1); i++) c2[i] = -d[i];
        QMG_CHK(qmg_multi_axpy(c2.data(), ptrs.data(), j + 1, qmg_host::P(w), n));
        for (int i = 0; i <= j; i++)
        {
          const cd hij(c[2 * i] + d[2 * i], c[2 * i + 1] + d[2 * i + 1]);
          T[(size_t)i * m + j] = hij;
          T[(size_t)j * m + i] = std::conj(hij);
        }
        T[(size_t)j * m + j] = T[(size_t)j * m + j].real();
        const double beta = sqrt(norm2sq(w, n));
        if (j + 1 < m) { T[(size_t)(j + 1) * m + j] = beta; T[(size_t)j * m + (j + 1)] = beta; }
        else beta_m = beta;
        if (beta < 1e-14) { gaussian(w, n, gen); }      // invariant subspace: continue with a fresh direction (beta stays 0 in T)
        caxy(1.0 / (beta < 1e-14 ? sqrt(norm2sq(w, n)) : beta), w, V[j + 1], n);
      }
      qmg_host::hermitian_eig(T, m, theta, S);
      // wanted Ritz pairs: indices at the requested end
      std::vector<int> want(nev);
      for (int i = 0; i < nev; i++) want[i] = want_low ? i : m - 1 - i;
      done = true;
      for (int i = 0; i < nev; i++)
      {
        const double res = beta_m * std::abs(S[(size_t)(m - 1) * m + want[i]]);
        if (res > tol * std::max(std::fabs(theta[want[i]]), 1e-300)) done = false;
      }
      // keep the wanted pairs and (for faster convergence) their neighbours: thick restart
      const int keep = done ? nev : std::min(m - 2, nev + std::max(1, (m - nev) / 2));
      std::vector<int> sel(keep);
      for (int i = 0; i < keep; i++) sel[i] = want_low ? i : m - 1 - i;
      std::vector<cd*> Y(keep);
      for (int i = 0; i < keep; i++)
      {
        Y[i] = allocate_vector<cd>(n);
        std::vector<double> coef(2 * m); std::vector<const qmg_cplx*> ptrs(m);
        for (int j = 0; j < m; j++) { const cd sji = S[(size_t)j * m + sel[i]]; coef[2 * j] = sji.real(); coef[2 * j + 1] = sji.imag(); ptrs[j] = qmg_host::P(V[j]); }
        zero_vector(Y[i], n);
        QMG_CHK(qmg_multi_axpy(coef.data(), ptrs.data(), m, qmg_host::P(Y[i]), n));
      }
      if (done)
      {
        evals_.resize(nev); evecs_.resize(nev);
        for (int i = 0; i < nev; i++) { evals_[i] = theta[sel[i]]; evecs_[i] = Y[i]; }
        break;
      }
      // restart: basis = kept Ritz vectors + the residual direction; T = diag(theta) bordered by beta_m s_i(m)
      for (int i = 0; i < keep; i++) { copy_vector(V[i], Y[i], n); deallocate_vector(&Y[i]); }
      copy_vector(V[keep], V[m], n);
      std::fill(T.begin(), T.end(), cd(0.0));
      for (int i = 0; i < keep; i++)
      {
        T[(size_t)i * m + i] = theta[sel[i]];
        const cd b = beta_m * S[(size_t)(m - 1) * m + sel[i]];
        T[(size_t)keep * m + i] = b;
        T[(size_t)i * m + keep] = std::conj(b);
      }
      k = keep;
    }
    cleanup();
    if (!done) { std::cout << "[QMG-ERROR]: arpack_dcn: Lanczos did not converge in " << maxitr << " restarts.\n"; release(); return false; }
    which_ = piece;
    ready = true;
    return true;
  }

  // ARPACK's own default for the Krylov dimension when the caller gives none (tests/n12_wilson_eigenvalue_test/wilson_test.cpp:208)
  bool prepare_eigensystem(arpack_spectrum_piece piece, int nev_in) { return prepare_eigensystem(piece, nev_in, std::min(n, std::max(2 * nev_in + 1, 20))); }
  // what the reference reads after a failure (n12 :210): the ARPACK return codes have no meaning here
  struct arpack_solve_info { int znaupd_code, zneupd_code, nconv, niter; bool is_error; };
  arpack_solve_info get_solve_info() const { arpack_solve_info i; i.znaupd_code = ready ? 0 : -9999; i.zneupd_code = i.znaupd_code; i.nconv = ready ? nev : 0; i.niter = restarts; i.is_error = !ready; return i; }

  // copy the eigenpairs computed by prepare_eigensystem (evecs: nev device vectors of the caller)
  bool get_eigensystem(std::complex<double>* evals, std::complex<double>** evecs, arpack_spectrum_piece)
  {
    if (!ready) return false;
    for (int i = 0; i < nev; i++) { evals[i] = evals_[i]; if (evecs != 0) copy_vector(evecs[i], evecs_[i], n); }
    return true;
  }
  bool get_eigensystem(std::complex<double>* evals, arpack_spectrum_piece piece) { return get_eigensystem(evals, 0, piece); }
  // the whole spectrum is a dense problem (tests/n13 :485 prints it for tiny lattices); not a Lanczos job
  bool get_entire_eigensystem(std::complex<double>*, arpack_spectrum_piece) { std::cout << "[QMG-ERROR]: arpack_dcn::get_entire_eigensystem is not available on B200.\n"; return false; }
  bool get_entire_eigensystem(std::complex<double>*, std::complex<double>**, arpack_spectrum_piece) { std::cout << "[QMG-ERROR]: arpack_dcn::get_entire_eigensystem is not available on B200.\n"; return false; }
  int get_restarts() const { return restarts; }
  int get_ops_count() const { return ops; }

private:
  arpack_dcn(const arpack_dcn&); arpack_dcn& operator=(const arpack_dcn&);
  void release() { for (size_t i = 0; i < evecs_.size(); i++) deallocate_vector(&evecs_[i]); evecs_.clear(); evals_.clear(); ready = false; }
  int n, maxitr; double tol; matrix_op_cplx op; void* extra; int nev, ncv; bool ready; int restarts, ops;
  arpack_spectrum_piece which_;
  std::vector<double> evals_; std::vector<std::complex<double>*> evecs_;
};

#endif
