// quantum-mg on B200 -- batched nc x nc site-matrix routines with quantum-linalg's signatures
// ("blas/generic_local_matrix.h"; call sites stencil/stencil_2d.h:675,1055,1097,1536-1564, operators/dwf.h:106).
// Row-major blocks, one per site; all pointers are device memory.
#ifndef QMG_B200_GENERIC_LOCAL_MATRIX
#define QMG_B200_GENERIC_LOCAL_MATRIX

#include "generic_vector.h"

inline void cMATxy(const qmg_cd* M, const qmg_cd* x, qmg_cd* y, long nsites, int nrow, int ncol)
{ (void)ncol; QMG_CHK(qmg_cmat_xy(qmg_host::P(M), qmg_host::P(x), qmg_host::P(y), nsites, nrow, 0)); }
inline void cMATxpy(const qmg_cd* M, const qmg_cd* x, qmg_cd* y, long nsites, int nrow, int ncol)
{ (void)ncol; QMG_CHK(qmg_cmat_xy(qmg_host::P(M), qmg_host::P(x), qmg_host::P(y), nsites, nrow, 1)); }
inline void cMAT_single_xy(const qmg_cd* M, const qmg_cd* x, qmg_cd* y, long nsites, int nrow, int ncol)
{ (void)ncol; QMG_CHK(qmg_cmat_single_xy(qmg_host::P(M), qmg_host::P(x), qmg_host::P(y), nsites, nrow)); }
inline void cMATcopy_conjtrans_square(const qmg_cd* in, qmg_cd* out, long nsites, int nc)
{ QMG_CHK(qmg_cmat_conjtrans(qmg_host::P(in), qmg_host::P(out), nsites, nc)); }
inline void cMATconjtrans_square(qmg_cd* M, long nsites, int nc)
{ QMG_CHK(qmg_cmat_conjtrans(qmg_host::P(M), qmg_host::P(M), nsites, nc)); }
inline void cMATxtMATyMATz_square(const qmg_cd* X, const qmg_cd* Y, qmg_cd* Z, long nsites, int nc)
{ QMG_CHK(qmg_cmat_mul(qmg_host::P(X), qmg_host::P(Y), qmg_host::P(Z), nsites, nc)); }
// The reference inverts through a batched QR pair (stencil_2d.h:1536-1537): M = Q R (Q unitary, R upper triangular, as
// anyone who reads them expects), then Minv = R^-1 Q^dag.  (Stencil2D::build_rbjacobi_stencil itself goes through
// qmg_build_rbjacobi, whose inverse is the pivoted Gauss-Jordan kernel.)
inline void cMATx_do_qr_square(const qmg_cd* M, qmg_cd* Q, qmg_cd* R, long nsites, int nc)
{ QMG_CHK(qmg_cmat_qr(qmg_host::P(M), qmg_host::P(Q), qmg_host::P(R), nsites, nc)); }
inline void cMATqr_do_xinv_square(const qmg_cd* Q, const qmg_cd* R, qmg_cd* Minv, long nsites, int nc)
{ QMG_CHK(qmg_cmat_qr_inverse(qmg_host::P(Q), qmg_host::P(R), qmg_host::P(Minv), nsites, nc)); }
inline void cMATinverse_square(const qmg_cd* M, qmg_cd* Minv, long nsites, int nc)
{ QMG_CHK(qmg_cmat_inverse(qmg_host::P(M), qmg_host::P(Minv), nsites, nc)); }

#endif
