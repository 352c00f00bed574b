// quantum-mg on B200 -- U(1) gauge-field utilities with the reference's names (/root/reference/u1/u1_utils.h).
// Fields live in device memory: `gauge_field` is 2 V complex links, `phase_field` 2 V real angles, both in the
// nc = 1 gauge layout (lattice.h gauge_coord_to_index).  Observables, gauge transformation, APE smearing and the
// heatbath are single fused kernels of libqmg_b200.so (csrc/qmg_gauge.cu); file I/O and the instanton
// constructors, whose definitions are host loops over (x, y), stage through the host.
#ifndef QMG_B200_U1_UTILS
#define QMG_B200_U1_UTILS

#include <cmath>
#include <complex>
#include <fstream>
#include <random>
#include <string>
#include <vector>

#include "../blas/generic_vector.h"
#include "../lattice/lattice.h"
#include "../cshift/cshift_2d.h"

#ifndef PI
#define PI 3.14159265358979323846
#endif

enum qmg_gauge_create_type
{
  GAUGE_LOAD = 0,
  GAUGE_RANDOM = 1,
  GAUGE_UNIT = 2
};

namespace qmg_host {
inline bool u1_lattice_ok(Lattice2D* lat)
{
  if (lat->get_nc() == 1) return true;
  std::cout << "[QMG-ERROR]: U1 gauge functions require Nc = 1 lattice.\n";
  return false;
}
inline void upload_real(double* dev, const double* host, long n) { check(qmg_memcpy_h2d(dev, host, sizeof(double) * (size_t)n), "upload"); }
inline void download_real(double* host, const double* dev, long n) { check(qmg_memcpy_d2h(host, dev, sizeof(double) * (size_t)n), "download"); }
// visit the links in file order: x outer, y, mu inner (u1_utils.h:53-63)
template <class F> inline void for_each_link_in_file_order(Lattice2D* lat, F f)
{
  const int X = lat->get_dim_mu(0), Y = lat->get_dim_mu(1);
  for (int x = 0; x < X; x++)
    for (int y = 0; y < Y; y++)
      for (int mu = 0; mu < 2; mu++) f(lat->gauge_coord_to_index(x, y, 0, 0, mu), x, y, mu);
}
}

// double-field flavours of the quantum-linalg calls the drivers make on phase fields (n13 :203, :212)
// (zero_vector / copy_vector on real fields: the templates of blas/generic_vector.h)
inline void polar_vector(const double* phases, complex<double>* out, long n) { QMG_CHK(qmg_polar_vector(phases, qmg_host::P(out), n)); }

// ---- file format: one phase per line, x outer, y, mu inner (u1_utils.h:38-168)
inline void read_phase_u1(double* phase_field, Lattice2D* lat, std::string input_file)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  std::vector<double> h((size_t)lat->get_size_gauge(), 0.0);
  std::ifstream in(input_file.c_str());
  qmg_host::for_each_link_in_file_order(lat, [&](int idx, int, int, int) { double ph = 0.0; in >> ph; h[idx] = ph; });
  qmg_host::upload_real(phase_field, h.data(), (long)h.size());
}
inline void read_gauge_u1(complex<double>* gauge_field, Lattice2D* lat, std::string input_file)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  std::vector<complex<double> > h((size_t)lat->get_size_gauge());
  std::ifstream in(input_file.c_str());
  qmg_host::for_each_link_in_file_order(lat, [&](int idx, int, int, int) { double ph = 0.0; in >> ph; h[idx] = std::polar(1.0, ph); });
  qmg_host::upload(gauge_field, h.data(), (long)h.size());
}
inline void write_gauge_u1(double* phase_field, Lattice2D* lat, std::string output_file)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  std::vector<double> h((size_t)lat->get_size_gauge());
  qmg_host::download_real(h.data(), phase_field, (long)h.size());
  std::ofstream out(output_file.c_str(), std::ios::trunc);
  out.setf(std::ios_base::fixed, std::ios_base::floatfield);
  out.precision(20);
  qmg_host::for_each_link_in_file_order(lat, [&](int idx, int, int, int) { out << h[idx] << "\n"; });
}
inline void write_gauge_u1(complex<double>* gauge_field, Lattice2D* lat, std::string output_file)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  std::vector<complex<double> > h((size_t)lat->get_size_gauge());
  qmg_host::download(h.data(), gauge_field, (long)h.size());
  std::ofstream out(output_file.c_str(), std::ios::trunc);
  out.setf(std::ios_base::fixed, std::ios_base::floatfield);
  out.precision(20);
  qmg_host::for_each_link_in_file_order(lat, [&](int idx, int, int, int) { out << std::arg(h[idx]) << "\n"; });
}

// ---- field constructors (u1_utils.h:172-238)
inline void unit_gauge_u1(complex<double>* gauge_field, Lattice2D* lat)
{ if (qmg_host::u1_lattice_ok(lat)) constant_vector(gauge_field, 1.0, lat->get_size_gauge()); }
inline void rand_gauge_u1(complex<double>* gauge_field, Lattice2D* lat, std::mt19937& generator)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  random_uniform(gauge_field, lat->get_size_gauge(), generator, -PI, PI);
  polar(gauge_field, lat->get_size_gauge());
}
inline void gauss_gauge_u1(complex<double>* gauge_field, Lattice2D* lat, std::mt19937& generator, double beta)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  beta = std::fabs(beta);
  if (beta == 0) { rand_gauge_u1(gauge_field, lat, generator); return; }
  gaussian(gauge_field, lat->get_size_gauge(), generator, 1.0 / sqrt(beta));
  polar(gauge_field, lat->get_size_gauge());
}
inline void rand_trans_u1(complex<double>* gauge_trans, Lattice2D* lat, std::mt19937& generator)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  random_uniform(gauge_trans, lat->get_size_cm(), generator, -PI, PI);
  polar(gauge_trans, lat->get_size_cm());
}

// ---- fused kernels
inline void apply_gauge_trans_u1(complex<double>* gauge_field, complex<double>* gauge_trans, Lattice2D* lat)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  QMG_CHK(qmg_u1_gauge_transform(qmg_host::P(gauge_field), qmg_host::P(gauge_trans), lat->get_dim_mu(0), lat->get_dim_mu(1)));
}
// As the reference computes it: the y staples are accumulated on the x links (u1_utils.h:352,:372 add to `smeared_field`
// where `smeared_field + size_cm` is meant) -- kept, because drivers compare against the reference's numbers.
inline void apply_ape_smear_u1(complex<double>* smeared_field, complex<double>* gauge_field, Lattice2D* lat, double alpha, int n_iter)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  QMG_CHK(qmg_u1_ape_smear(qmg_host::P(smeared_field), qmg_host::P(gauge_field), lat->get_dim_mu(0), lat->get_dim_mu(1), alpha, n_iter, 0));
}
// The smearing the reference's comments describe (every link with its own two staples).
inline void apply_ape_smear_textbook_u1(complex<double>* smeared_field, complex<double>* gauge_field, Lattice2D* lat, double alpha, int n_iter)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  QMG_CHK(qmg_u1_ape_smear(qmg_host::P(smeared_field), qmg_host::P(gauge_field), lat->get_dim_mu(0), lat->get_dim_mu(1), alpha, n_iter, 1));
}
inline double get_noncompact_action_u1(double* phase_field, double beta, Lattice2D* lat)
{
  if (!qmg_host::u1_lattice_ok(lat)) return -50;
  double s = 0.0;
  QMG_CHK(qmg_u1_noncompact_action(phase_field, lat->get_dim_mu(0), lat->get_dim_mu(1), beta, &s));
  return s;
}
inline complex<double> get_plaquette_u1(complex<double>* gauge_field, Lattice2D* lat)
{
  if (!qmg_host::u1_lattice_ok(lat)) return -50;
  double r[4];
  QMG_CHK(qmg_u1_plaquette(qmg_host::P(gauge_field), lat->get_dim_mu(0), lat->get_dim_mu(1), r));
  return complex<double>(r[0], r[1]);
}
inline double get_topo_u1(complex<double>* gauge_field, Lattice2D* lat)
{
  if (!qmg_host::u1_lattice_ok(lat)) return -50.1;
  double r[4];
  QMG_CHK(qmg_u1_plaquette(qmg_host::P(gauge_field), lat->get_dim_mu(0), lat->get_dim_mu(1), r));
  return r[2];
}
// The reference's body is an unfinished loop that never terminates (u1_utils.h:533-537); nothing calls it.
inline void lorentz_gauge_fix_u1(complex<double>*, Lattice2D* lat, const double, const double, const int)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  std::cout << "[QMG-WARNING]: lorentz_gauge_fix_u1 is not implemented (the reference's version does not terminate).\n";
}

// ---- instantons: host loops over (x, y) in the reference (u1_utils.h:545-600)
inline void create_instanton_u1(complex<double>* gauge_field, Lattice2D* lat, double Q, const int x0, const int y0)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  const int X = lat->get_dim_mu(0), Y = lat->get_dim_mu(1);
  std::vector<complex<double> > h((size_t)lat->get_size_gauge());
  qmg_host::download(h.data(), gauge_field, (long)h.size());
  for (int x = 0; x < X; x++)
    for (int y = 0; y < Y; y++)
    {
      const double rx = x - X / 2 + 0.5, ry = y - Y / 2 + 0.5, r2 = rx * rx + ry * ry;
      const int xs = (x - X / 2 + x0 + 3 * X) % X, ys = (y - Y / 2 + y0 + 3 * Y) % Y;
      h[lat->gauge_coord_to_index(xs, ys, 0, 0, 0)] *= std::polar(1.0, Q * ry / r2);
      h[lat->gauge_coord_to_index(xs, ys, 0, 0, 1)] *= std::polar(1.0, -Q * rx / r2);
    }
  qmg_host::upload(gauge_field, h.data(), (long)h.size());
}
inline void create_noncompact_instanton_u1(double* phase_field, Lattice2D* lat, double Q)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  const int X = lat->get_dim_mu(0), Y = lat->get_dim_mu(1);
  std::vector<double> h((size_t)lat->get_size_gauge());
  qmg_host::download_real(h.data(), phase_field, (long)h.size());
  for (int x = 0; x < X; x++)
    for (int y = 0; y < Y; y++)
    {
      h[lat->gauge_coord_to_index(x, y, 0, 0, 0)] += -Q * 3.1415926535 * y / (X * Y);
      if (y == Y - 1) h[lat->gauge_coord_to_index(x, y, 0, 0, 1)] += Q * 3.1415926535 * x / X;
    }
  qmg_host::upload_real(phase_field, h.data(), (long)h.size());
}

// ---- non-compact heatbath (u1_utils.h:607-667).  The device sweep updates four independent link subsets per update
// from a counter-based generator keyed by two words drawn from `generator`; successive calls draw fresh keys.
inline void heatbath_noncompact_update(double* phase_field, Lattice2D* lat, double beta, int n_update, std::mt19937& generator)
{
  if (!qmg_host::u1_lattice_ok(lat)) return;
  const unsigned long long hi = generator(), lo = generator();
  QMG_CHK(qmg_u1_heatbath(phase_field, lat->get_dim_mu(0), lat->get_dim_mu(1), beta, n_update, (hi << 32) | lo, 0ULL));
}

#endif
