// quantum-mg on B200 -- 2D staggered operator, one dof per site, hopping only
// (/root/reference/operators/staggered.h:21-259):
//   H_{+x} = -1/2 U_x,  H_{+y} = -1/2 eta U_y,  H_{-x} = +1/2 U_x*(x - x^),  H_{-y} = +1/2 eta U_y*(x - y^),
//   eta = 1 - 2 (x mod 2),  shift = mass;  gamma_5 = +1 on even, -1 on odd sites.
#ifndef QMG_B200_STAGGERED
#define QMG_B200_STAGGERED

#include "../stencil/stencil_2d.h"

struct Staggered2D : public Stencil2D
{
protected:
  Staggered2D(Staggered2D const&);
  Staggered2D& operator=(Staggered2D const&);
  complex<double>* tmp_eo_space;
  long half() const { return lat->get_size_cv() / 2; }

public:
  Staggered2D(Lattice2D* in_lat, complex<double> mass, complex<double>* gauge_links)
    : Stencil2D(in_lat, QMG_PIECE_HOPPING, mass, 0.0, 0.0), tmp_eo_space(0)
  {
    if (lat->get_nc() != 1) { std::cout << "[QMG-ERROR]: Staggered2D only supports Nc = 1.\n"; return; }
    update_links(gauge_links);
  }
  ~Staggered2D() { if (tmp_eo_space != 0) deallocate_vector(&tmp_eo_space); }

  void update_links(complex<double>* gauge_links)
  {
    QMG_CHK(qmg_fill_staggered(lat->get_dim_mu(0), lat->get_dim_mu(1), qmg_host::P(gauge_links), qmg_host::P(hopping)));
    free_derived_stencils();
    generated = true;
  }

  static int get_dof(int i = 0) { (void)i; return 1; }
  static chirality_state has_chirality() { return QMG_CHIRAL_YES; }

  virtual void gamma5(complex<double>* vec) { cax(-1.0, vec + half(), half()); }
  virtual void gamma5(complex<double>* g5_vec, complex<double>* vec) { copy_vector(g5_vec, vec, half()); caxy(-1.0, vec + half(), g5_vec + half(), half()); }
  // chirality is the site parity
  virtual void chiral_projection(complex<double>* vector, bool is_up) { zero_vector(is_up ? vector + half() : vector, half()); }
  virtual void chiral_projection_copy(complex<double>* orig, complex<double>* dest, bool is_up)
  {
    const long keep = is_up ? 0 : half(), kill = is_up ? half() : 0;
    zero_vector(dest + kill, half());
    copy_vector(dest + keep, orig + keep, half());
  }
  virtual void chiral_projection_both(complex<double>* orig_to_up, complex<double>* down)
  {
    zero_vector(down, half());
    copy_vector(down + half(), orig_to_up + half(), half());
    zero_vector(orig_to_up + half(), half());
  }
  virtual QMGDefaultChirality get_default_chirality() { return QMG_CHIRALITY_GAMMA_5; }

  // even-odd preconditioned normal system (m^2 - D_eo D_oe) on the even sites (staggered.h:188-242)
  void prepare_b(complex<double>* b_new, complex<double>* b)
  {
    zero_vector(b_new, half());
    apply_M_eo(b_new, b);
    caxpby(shift, b, complex<double>(-1.0), b_new, half());
  }
  void apply_eo_prec_M(complex<double>* lhs, complex<double>* rhs)
  {
    if (tmp_eo_space == 0) tmp_eo_space = allocate_vector<complex<double> >(lat->get_size_cv());
    launch(QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ODD_ROWS_ONLY, 15, tmp_eo_space, rhs);
    launch(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_EVEN_ROWS_ONLY, 15, tmp_eo_space, tmp_eo_space);
    caxpbyz(shift * shift, rhs, complex<double>(-1.0), tmp_eo_space, lhs, half());
  }
  void reconstruct_x(complex<double>* x, complex<double>* b)
  {
    zero_vector(x + half(), half());
    apply_M_oe(x, x);
    caxpby(1.0 / shift, b + half(), -1.0 / shift, x + half(), half());
  }
};

inline void apply_eo_staggered_2D_M(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{ ((Staggered2D*)extra_data)->apply_eo_prec_M(lhs, rhs); }

// eta_y phase factor for one ColorMatrix element (staggered.h:253-259); a host callback for arb_local_function_vector
inline void staggered_set_eta_y(int i, complex<double>& elem, void* extra_data)
{
  Lattice2D* lat = (Lattice2D*)extra_data;
  int x, y, c1, c2;
  lat->cm_index_to_coord(i, x, y, c1, c2);
  elem *= (double)(1.0 - 2.0 * (x % 2));
}

#endif
