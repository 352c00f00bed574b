// quantum-mg on B200 -- per-time-slice reductions and wall sources with the reference's names
// (/root/reference/reductions/reductions.h): the correlator measurements of tests/n15, n16, n20.
// Vectors are device memory; each reduction is one kernel (one CTA per row y) and returns Y numbers to the host.
#ifndef QMG_B200_REDUCTIONS
#define QMG_B200_REDUCTIONS

#include <complex>
#include <random>
#include <vector>

#include "../blas/generic_vector.h"
#include "../lattice/lattice.h"

#ifndef PI
#define PI 3.14159265358979323846
#endif

// sum[y] = sum_{x,c} |cv(x,y,c)|^2   (reductions.h:24-42)
inline void norm2sq_cv_timeslice(double* sum, complex<double>* cv, Lattice2D* lat)
{ QMG_CHK(qmg_timeslice_reduce(0, qmg_host::P(cv), 0, lat->get_dim_mu(0), lat->get_dim_mu(1), lat->get_nc(), sum)); }
// sum[y] = sum_{x,c} Re conj(cv1) cv2   (reductions.h:47-66)
inline void redot_cv_timeslice(double* sum, complex<double>* cv1, complex<double>* cv2, Lattice2D* lat)
{ QMG_CHK(qmg_timeslice_reduce(1, qmg_host::P(cv1), qmg_host::P(cv2), lat->get_dim_mu(0), lat->get_dim_mu(1), lat->get_nc(), sum)); }
// sum[y] = sum_{x,c} conj(cv1) cv2   (reductions.h:71-90)
inline void dot_cv_timeslice(complex<double>* sum, complex<double>* cv1, complex<double>* cv2, Lattice2D* lat)
{ QMG_CHK(qmg_timeslice_reduce(2, qmg_host::P(cv1), qmg_host::P(cv2), lat->get_dim_mu(0), lat->get_dim_mu(1), lat->get_nc(), reinterpret_cast<double*>(sum))); }

// Real gaussian numbers on one time slice and one colour, zero elsewhere (reductions.h:93-161).  The draws are made on
// the host in colour-vector index order, exactly as the reference's loop consumes its generator, and uploaded.
inline void gaussian_wall_source(complex<double>* cv, int timeslice, int color, Lattice2D* lat, std::mt19937& generator, double deviation = 1.0, double mean = 0.0)
{
  if (timeslice >= lat->get_dim_mu(lat->get_nd() - 1)) { std::cout << "[QMG-ERROR]: Cannot create gaussian wall source for t < Nt.\n"; return; }
  const int nc = lat->get_nc();
  if (color >= nc) { std::cout << "[QMG-ERROR]: Cannot create gaussian wall source for color < Nc.\n"; return; }
  const int xh = lat->get_dim_mu(0) / 2, Y = lat->get_dim_mu(1);
  std::vector<complex<double> > h((size_t)lat->get_size_cv(), 0.0);
  std::normal_distribution<> dist(0.0, deviation);
  // index order = even parity first: row `timeslice` of each parity half is xh consecutive sites
  for (int p = 0; p < 2; p++)
    for (int k = 0; k < xh; k++)
      h[((size_t)(timeslice + p * Y) * xh + k) * nc + color] = complex<double>(mean + dist(generator), 0.0);
  qmg_host::upload(cv, h.data(), (long)h.size());
}

#endif
