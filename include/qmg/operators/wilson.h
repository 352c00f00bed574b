// quantum-mg on B200 -- 2D Wilson operator, two spin components per site
// (/root/reference/operators/wilson.h:19-226):
//   clover = 2 w 1,  H_{+x} = 1/2 [[-w, 1],[ 1,-w]] U_x(x),      H_{+y} = 1/2 [[-w,-i],[ i,-w]] U_y(x),
//   H_{-x} = 1/2 [[-w,-1],[-1,-w]] U_x*(x - x^),                 H_{-y} = 1/2 [[-w, i],[-i,-w]] U_y*(x - y^),  shift = mass.
// The link fill is one kernel over the gauge field (qmg_fill_wilson) instead of 16 strided BLAS sweeps and 2 cshifts.
#ifndef QMG_B200_WILSON
#define QMG_B200_WILSON

#include "../stencil/stencil_2d.h"

struct Wilson2D : public Stencil2D
{
protected:
  Wilson2D(Wilson2D const&);
  Wilson2D& operator=(Wilson2D const&);
  double wilson_coeff;

  // out[s][i] = scale[i] in[s][pick[i]] on the two spin components
  void spin_map(double s0, double s1, int p0, int p1, complex<double>* in, complex<double>* out)
  {
    const double scale[2] = { s0, s1 }; const int pick[2] = { p0, p1 };
    caxy_shuffle_pattern(scale, pick, 2, in, out, lat->get_volume());
  }

public:
  Wilson2D(Lattice2D* in_lat, complex<double> mass, complex<double>* gauge_links, double wilson_coeff = 1.0)
    : Stencil2D(in_lat, QMG_PIECE_CLOVER_HOPPING, mass, 0.0, 0.0), wilson_coeff(wilson_coeff)
  {
    if (lat->get_nc() != 2) { std::cout << "[QMG-ERROR]: Wilson2D only supports Nc = 2.\n"; return; }
    update_links(gauge_links);
  }
  ~Wilson2D() { }

  // gauge_links: 2 V complex links of the nc = 1 lattice, device memory (wilson.h:153-226)
  void update_links(complex<double>* gauge_links)
  {
    QMG_CHK(qmg_fill_wilson(lat->get_dim_mu(0), lat->get_dim_mu(1), wilson_coeff, qmg_host::P(gauge_links), qmg_host::P(clover), qmg_host::P(hopping)));
    free_derived_stencils();
    generated = true;
  }

  // B200 extension (not in the reference): the five stored 2 x 2 blocks of a site are functions of four U(1) links, so the
  // whole-operator apply can read the links (96 instead of 384 bytes per site from HBM) and rebuild the block elements on the fly
  // with the arithmetic of the fill -- the output has the same bits as the stored-block apply.  The operator takes its own copy
  // of gauge_links (the caller keeps ownership of its array, as in the constructor) and CHECKS that the stored blocks equal the
  // regenerated ones exactly first (false and nothing changes otherwise, e.g. after the n18 mutation of the clover).  Anything
  // that edits the blocks afterwards (update_links, clear_stencils) drops it again; writing through the public pointers must be
  // followed by disable_matrix_free_apply().  Pieces, dagger / rbjacobi / Schur variants keep reading the stored blocks.
  bool enable_matrix_free_apply(complex<double>* gauge_links)
  {
    disable_matrix_free_apply();
    if (lat->get_nc() != 2 || clover == 0 || hopping == 0 || swap_dagger || swap_rbjacobi || swap_rbj_dagger) return false;
    const long V = lat->get_volume();
    complex<double>* copy = allocate_vector<complex<double> >(2 * V);
    copy_vector(copy, gauge_links, 2 * V);
    complex<double>* halo = 0;
    if (qmg_comm_active())
    {
      // on a y-slab row -1 of U_y lives on the lower rank: fetched once
      const long row = lat->get_dim_mu(0);
      halo = allocate_vector<complex<double> >(row);
      complex<double>* unused = allocate_vector<complex<double> >(row);
      QMG_CHK(qmg_halo_exchange(qmg_host::P(copy + V), lat->get_dim_mu(0), lat->get_dim_mu(1), 1, qmg_host::P(halo), qmg_host::P(unused)));
      deallocate_vector(&unused);
    }
    mf_gauge = copy; mf_gauge_halo_ym = halo; mf_w = wilson_coeff;
    qmg_stencil_desc d = describe();
    double dev[2] = {1.0, 0.0};
    QMG_CHK(qmg_wilson_mf_deviation(&d, dev));
    if (dev[0] != 0.0) { disable_matrix_free_apply(); return false; }
    return true;
  }

  static int get_dof(int i = 0) { (void)i; return 2; }
  static chirality_state has_chirality() { return QMG_CHIRAL_YES; }

  // gamma_5 = diag(1, -1) on the spin index
  virtual void gamma5(complex<double>* vec) { spin_map(1.0, -1.0, 0, 1, vec, vec); }
  virtual void gamma5(complex<double>* g5_vec, complex<double>* vec) { spin_map(1.0, -1.0, 0, 1, vec, g5_vec); }
  // upper component = "up", lower = "down"
  virtual void chiral_projection(complex<double>* vector, bool is_up) { spin_map(is_up ? 1.0 : 0.0, is_up ? 0.0 : 1.0, 0, 1, vector, vector); }
  virtual void chiral_projection_copy(complex<double>* orig, complex<double>* dest, bool is_up) { spin_map(is_up ? 1.0 : 0.0, is_up ? 0.0 : 1.0, 0, 1, orig, dest); }
  virtual void chiral_projection_both(complex<double>* orig_to_up, complex<double>* down)
  {
    spin_map(0.0, 1.0, 0, 1, orig_to_up, down);
    spin_map(1.0, 0.0, 0, 1, orig_to_up, orig_to_up);
  }
  // sigma_1 swaps the two components
  virtual void sigma1(complex<double>* vec)
  {
    complex<double>* tmp = scratch_extra();
    spin_map(1.0, 1.0, 1, 0, vec, tmp);
    copy_vector(vec, tmp, lat->get_size_cv());
  }
  virtual void sigma1(complex<double>* s1_vec, complex<double>* vec) { spin_map(1.0, 1.0, 1, 0, vec, s1_vec); }
  virtual QMGDefaultChirality get_default_chirality() { return QMG_CHIRALITY_GAMMA_5; }
};

#endif
