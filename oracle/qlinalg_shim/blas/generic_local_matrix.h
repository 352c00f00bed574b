// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// Clean-room stand-in for quantum-linalg's batched small-matrix routines.
// Matrices are row-major nrow x ncol blocks, one per site, sites contiguous
// (the LatticeColorMatrix layout of /root/reference/lattice/lattice.h:121).
// Semantics restated from the reference's call sites, cited per function.

#ifndef QLINALG_SHIM_GENERIC_LOCAL_MATRIX
#define QLINALG_SHIM_GENERIC_LOCAL_MATRIX

#include <cmath>
#include <complex>

// stencil/stencil_2d.h:1055  y = M x   (per site)
template <typename T> inline void cMATxy(const T* M, const T* x, T* y, int nsites, int nrow, int ncol)
{
  for (int s = 0; s < nsites; s++)
    for (int r = 0; r < nrow; r++)
    {
      T acc = 0;
      const T* row = M + ((long)s * nrow + r) * ncol;
      for (int c = 0; c < ncol; c++) acc += row[c] * x[(long)s * ncol + c];
      y[(long)s * nrow + r] = acc;
    }
}

// stencil/stencil_2d.h:675  y += M x   (per site)
template <typename T> inline void cMATxpy(const T* M, const T* x, T* y, int nsites, int nrow, int ncol)
{
  for (int s = 0; s < nsites; s++)
    for (int r = 0; r < nrow; r++)
    {
      T acc = 0;
      const T* row = M + ((long)s * nrow + r) * ncol;
      for (int c = 0; c < ncol; c++) acc += row[c] * x[(long)s * ncol + c];
      y[(long)s * nrow + r] += acc;
    }
}

// operators/dwf.h:106  one matrix shared by every site: y = M x
template <typename T> inline void cMAT_single_xy(const T* M, const T* x, T* y, int nsites, int nrow, int ncol)
{
  for (int s = 0; s < nsites; s++)
    for (int r = 0; r < nrow; r++)
    {
      T acc = 0;
      for (int c = 0; c < ncol; c++) acc += M[r * ncol + c] * x[(long)s * ncol + c];
      y[(long)s * nrow + r] = acc;
    }
}

// stencil/stencil_2d.h:1097  out = in^dagger per site
template <typename T> inline void cMATcopy_conjtrans_square(const T* in, T* out, int nsites, int nc)
{
  for (int s = 0; s < nsites; s++)
    for (int r = 0; r < nc; r++)
      for (int c = 0; c < nc; c++)
        out[((long)s * nc + r) * nc + c] = std::conj(in[((long)s * nc + c) * nc + r]);
}

// operators/coarse.h:810  in-place dagger per site
template <typename T> inline void cMATconjtrans_square(T* M, int nsites, int nc)
{
  for (int s = 0; s < nsites; s++)
  {
    T* m = M + (long)s * nc * nc;
    for (int r = 0; r < nc; r++)
    {
      m[r * nc + r] = std::conj(m[r * nc + r]);
      for (int c = r + 1; c < nc; c++)
      {
        T a = m[r * nc + c], b = m[c * nc + r];
        m[r * nc + c] = std::conj(b);
        m[c * nc + r] = std::conj(a);
      }
    }
  }
}

// stencil/stencil_2d.h:1564  Z = X * Y per site
template <typename T> inline void cMATxtMATyMATz_square(const T* X, const T* Y, T* Z, int nsites, int nc)
{
  for (int s = 0; s < nsites; s++)
  {
    const T* x = X + (long)s * nc * nc;
    const T* y = Y + (long)s * nc * nc;
    T* z = Z + (long)s * nc * nc;
    for (int r = 0; r < nc; r++)
      for (int c = 0; c < nc; c++)
      {
        T acc = 0;
        for (int k = 0; k < nc; k++) acc += x[r * nc + k] * y[k * nc + c];
        z[r * nc + c] = acc;
      }
  }
}

// stencil/stencil_2d.h:1536  M = Q R per site (modified Gram-Schmidt on columns,
// Q unitary, R upper triangular with real positive diagonal).
template <typename T> inline void cMATx_do_qr_square(const T* M, T* Q, T* R, int nsites, int nc)
{
  for (int s = 0; s < nsites; s++)
  {
    const T* m = M + (long)s * nc * nc;
    T* q = Q + (long)s * nc * nc;
    T* r = R + (long)s * nc * nc;
    for (int i = 0; i < nc * nc; i++) { q[i] = m[i]; r[i] = 0; }
    for (int j = 0; j < nc; j++)
    {
      for (int i = 0; i < j; i++)
      {
        T d = 0;
        for (int k = 0; k < nc; k++) d += std::conj(q[k * nc + i]) * q[k * nc + j];
        r[i * nc + j] = d;
        for (int k = 0; k < nc; k++) q[k * nc + j] -= d * q[k * nc + i];
      }
      double nrm = 0;
      for (int k = 0; k < nc; k++) nrm += std::norm(q[k * nc + j]);
      nrm = std::sqrt(nrm);
      r[j * nc + j] = nrm;
      for (int k = 0; k < nc; k++) q[k * nc + j] /= nrm;
    }
  }
}

// stencil/stencil_2d.h:1537  Minv = R^{-1} Q^dagger per site (back substitution).
template <typename T> inline void cMATqr_do_xinv_square(const T* Q, const T* R, T* Minv, int nsites, int nc)
{
  for (int s = 0; s < nsites; s++)
  {
    const T* q = Q + (long)s * nc * nc;
    const T* r = R + (long)s * nc * nc;
    T* x = Minv + (long)s * nc * nc;
    for (int c = 0; c < nc; c++)      // column c of the inverse solves R x = Q^dagger e_c
      for (int i = nc - 1; i >= 0; i--)
      {
        T acc = std::conj(q[c * nc + i]);
        for (int k = i + 1; k < nc; k++) acc -= r[i * nc + k] * x[k * nc + c];
        x[i * nc + c] = acc / r[i * nc + i];
      }
  }
}

#endif
