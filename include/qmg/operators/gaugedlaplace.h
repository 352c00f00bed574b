// quantum-mg on B200 -- gauged Laplace operator, one dof per site
// (/root/reference/operators/gaugedlaplace.h:18-215): clover = 4, H_mu = -U (conjugated and shifted for -mu), shift = m^2.
#ifndef QMG_B200_GAUGED_LAPLACE
#define QMG_B200_GAUGED_LAPLACE

#include "../stencil/stencil_2d.h"

struct GaugedLaplace2D : public Stencil2D
{
protected:
  GaugedLaplace2D(GaugedLaplace2D const&);
  GaugedLaplace2D& operator=(GaugedLaplace2D const&);
  complex<double>* tmp_eo_space;
  long half() const { return lat->get_size_cv() / 2; }

public:
  GaugedLaplace2D(Lattice2D* in_lat, complex<double> mass_sq, complex<double>* gauge_links)
    : Stencil2D(in_lat, QMG_PIECE_CLOVER_HOPPING, mass_sq, 0.0, 0.0), tmp_eo_space(0)
  {
    if (lat->get_nc() != 1) { std::cout << "[QMG-ERROR]: GaugedLaplace2D only supports Nc = 1.\n"; return; }
    update_links(gauge_links);
  }
  ~GaugedLaplace2D() { if (tmp_eo_space != 0) deallocate_vector(&tmp_eo_space); }

  void update_links(complex<double>* gauge_links)
  {
    QMG_CHK(qmg_fill_laplace(lat->get_dim_mu(0), lat->get_dim_mu(1), qmg_host::P(gauge_links), qmg_host::P(clover), qmg_host::P(hopping)));
    free_derived_stencils();
    generated = true;
  }

  static int get_dof(int i = 0) { (void)i; return 1; }
  static chirality_state has_chirality() { return QMG_CHIRAL_NO; }
  virtual void chiral_projection(complex<double>*, bool) { }
  virtual void chiral_projection_copy(complex<double>*, complex<double>*, bool) { }
  virtual void chiral_projection_both(complex<double>*, complex<double>*) { }
  virtual QMGDefaultChirality get_default_chirality() { return QMG_CHIRALITY_NONE; }

  // even-odd preconditioned system ((4 + m^2)^2 - D_eo D_oe) on the even sites (gaugedlaplace.h:154-205)
  void prepare_b(complex<double>* b_new, complex<double>* b)
  {
    zero_vector(b_new, half());
    apply_M_eo(b_new, b);
    caxpby(4.0 + shift, b, complex<double>(-1.0), b_new, half());
  }
  void apply_eo_prec_M(complex<double>* lhs, complex<double>* rhs)
  {
    if (tmp_eo_space == 0) tmp_eo_space = allocate_vector<complex<double> >(lat->get_size_cv());
    launch(QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ODD_ROWS_ONLY, 15, tmp_eo_space, rhs);
    launch(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_EVEN_ROWS_ONLY, 15, tmp_eo_space, tmp_eo_space);
    caxpbyz((4.0 + shift) * (4.0 + shift), rhs, complex<double>(-1.0), tmp_eo_space, lhs, half());
  }
  void reconstruct_x(complex<double>* x, complex<double>* b)
  {
    zero_vector(x + half(), half());
    apply_M_oe(x, x);
    caxpby(1.0 / (4.0 + shift), b + half(), -1.0 / (4.0 + shift), x + half(), half());
  }
};

inline void apply_eo_gauge_laplace_2D_M(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{ ((GaugedLaplace2D*)extra_data)->apply_eo_prec_M(lhs, rhs); }

#endif
