// quantum-mg on B200 -- periodic nearest-neighbour gather in the even-odd layout
// (/root/reference/cshift/cshift_2d.h:45-235): lhs(x) = rhs(x + dir), written on the parity opposite
// to each selected SOURCE parity.  The stencil kernels never call this (they read neighbours in place);
// it is kept for drivers and for the known-answer test tests/n00_cshift.
#ifndef QMG_B200_CSHIFT_2D
#define QMG_B200_CSHIFT_2D

#include <iostream>
#include "../lattice/lattice.h"
#include "blas/generic_vector.h"

enum qmg_cshift_dir
{
  QMG_CSHIFT_FROM_0 = 1,
  QMG_CSHIFT_FROM_XP1 = 2, QMG_CSHIFT_FROM_YP1 = 3, QMG_CSHIFT_FROM_XM1 = 4, QMG_CSHIFT_FROM_YM1 = 5,
  QMG_CSHIFT_FROM_XP2 = 6, QMG_CSHIFT_FROM_YP2 = 7, QMG_CSHIFT_FROM_XM2 = 8, QMG_CSHIFT_FROM_YM2 = 9,
  QMG_CSHIFT_FROM_XP1YP1 = 10, QMG_CSHIFT_FROM_XM1YP1 = 11, QMG_CSHIFT_FROM_XM1YM1 = 12, QMG_CSHIFT_FROM_XP1YM1 = 13,
};

enum qmg_eo
{
  QMG_EO_FROM_EVEN = 1,
  QMG_EO_FROM_ODD = 2,
  QMG_EO_FROM_EVENODD = 3,
};

inline void cshift(complex<double>* lhs, complex<double>* rhs, qmg_cshift_dir cdir, qmg_eo eo, int dof_per_site, Lattice2D* lat)
{
  if (cdir == QMG_CSHIFT_FROM_0)
  {
    // the reference copies volume/2 ELEMENTS in place, ignoring dof_per_site (cshift_2d.h:58,147); kept as is
    const long half = lat->get_volume() / 2;
    if (eo & QMG_EO_FROM_EVEN) copy_vector(lhs, rhs, half);
    if (eo & QMG_EO_FROM_ODD) copy_vector(lhs + half, rhs + half, half);
    return;
  }
  if (cdir > QMG_CSHIFT_FROM_YM1)
  {
    std::cout << "[QMG-ERROR]: Distance-2 and diagonal cshifts are not supported.\n";   // cshift_2d.h:120-129
    return;
  }
  if (lat->get_volume() == 1) return;
  QMG_CHK(qmg_cshift(qmg_host::P(lhs), qmg_host::P(rhs), (int)cdir, (int)eo, dof_per_site, lat->get_dim_mu(0), lat->get_dim_mu(1)));
}
inline void cshift_from_even(complex<double>* lhs, complex<double>* rhs, qmg_cshift_dir cdir, int dof, Lattice2D* lat) { cshift(lhs, rhs, cdir, QMG_EO_FROM_EVEN, dof, lat); }
inline void cshift_from_odd(complex<double>* lhs, complex<double>* rhs, qmg_cshift_dir cdir, int dof, Lattice2D* lat) { cshift(lhs, rhs, cdir, QMG_EO_FROM_ODD, dof, lat); }

#endif
