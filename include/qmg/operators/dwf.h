// quantum-mg on B200 -- Shamir domain-wall operator in 2D, nc = 2 Ls dof per site stored as dense site blocks
// (/root/reference/operators/dwf.h:20-296): Ls copies of the Wilson operator (clover 3 w, w = 1) on the 2x2 block
// diagonal, -P_+- between adjacent s slices, m P_+- wrap-around, shift = M5.  Applying it is the generic stencil
// kernel with a large nc (the same kernel as the coarse operator).
#ifndef QMG_B200_DWF
#define QMG_B200_DWF

#include <vector>
#include "../stencil/stencil_2d.h"

template <int Ls>
struct Dwf2D : public Stencil2D
{
protected:
  Dwf2D(Dwf2D const&);
  Dwf2D& operator=(Dwf2D const&);
  complex<double> mass;
  double M5;

public:
  Dwf2D(Lattice2D* in_lat, complex<double> mass, complex<double>* gauge_links, double M5 = -1.0)
    : Stencil2D(in_lat, QMG_PIECE_CLOVER_HOPPING, M5, 0.0, 0.0), mass(mass), M5(M5)
  {
    if (lat->get_nc() != 2 * Ls) { std::cout << "[QMG-ERROR]: Dwf2D only supports Nc = 2 Ls.\n"; return; }
    update_links(gauge_links);
  }
  ~Dwf2D() { }

  void update_links(complex<double>* gauge_links)
  {
    QMG_CHK(qmg_fill_dwf(lat->get_dim_mu(0), lat->get_dim_mu(1), Ls, 1.0, mass.real(), mass.imag(),
                         qmg_host::P(gauge_links), qmg_host::P(clover), qmg_host::P(hopping)));
    free_derived_stencils();
    generated = true;
  }

  static int get_dof() { return 2 * Ls; }
  static chirality_state has_chirality() { return QMG_CHIRAL_YES; }

  // gamma_5 (x) reflection in s: out[2 i + a] = (+1, -1)_a in[2 (Ls - 1 - i) + a]   (dwf.h:63-68,110-114)
  virtual void gamma5(complex<double>* g5_vec, complex<double>* vec)
  {
    double a[2 * Ls]; int pick[2 * Ls];
    for (int i = 0; i < Ls; i++) { a[2 * i] = 1.0; a[2 * i + 1] = -1.0; pick[2 * i] = 2 * (Ls - 1 - i); pick[2 * i + 1] = 2 * (Ls - 1 - i) + 1; }
    caxy_shuffle_pattern(a, pick, 2 * Ls, vec, g5_vec, lat->get_volume());
  }
  virtual void gamma5(complex<double>* vec)
  {
    complex<double>* tmp = scratch_extra();
    gamma5(tmp, vec);
    copy_vector(vec, tmp, lat->get_size_cv());
  }
  // the reference leaves the projections empty for this operator (dwf.h:117-147)
  virtual void chiral_projection(complex<double>*, bool) { }
  virtual void chiral_projection_copy(complex<double>*, complex<double>*, bool) { }
  virtual void chiral_projection_both(complex<double>*, complex<double>*) { }
  virtual QMGDefaultChirality get_default_chirality() { return QMG_CHIRALITY_GAMMA_5; }
};

static inline Stencil2D* createDwfLs(Lattice2D* in_lat, complex<double> mass, complex<double>* gauge_links, int Ls, double M5 = -1.0)
{
  switch (Ls)
  {
    case 2: return new Dwf2D<2>(in_lat, mass, gauge_links, M5);
    case 4: return new Dwf2D<4>(in_lat, mass, gauge_links, M5);
    case 6: return new Dwf2D<6>(in_lat, mass, gauge_links, M5);
    case 8: return new Dwf2D<8>(in_lat, mass, gauge_links, M5);
    case 12: return new Dwf2D<12>(in_lat, mass, gauge_links, M5);
    case 16: return new Dwf2D<16>(in_lat, mass, gauge_links, M5);
    case 24: return new Dwf2D<24>(in_lat, mass, gauge_links, M5);
    case 32: return new Dwf2D<32>(in_lat, mass, gauge_links, M5);
    default:
      std::cout << "[QMG-ERROR]: Unsupported Ls " << Ls << " for domain wall operator. Add a template to dwf.h.\n";
      return nullptr;
  }
}

#endif
