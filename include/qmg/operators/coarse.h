// quantum-mg on B200 -- CoarseOperator2D: the Galerkin coarse stencil R A P of a nearest-neighbour fine stencil
// (/root/reference/operators/coarse.h:29-899).  The reference finds the 5 dense nc_c x nc_c blocks per coarse site by
// 9 nc_c prolong -> one-direction apply -> restrict probes over the whole fine lattice (:140-444); here one CTA per
// coarse site contracts conj(R_a) A_piece P_b over its aggregate directly (qmg_coarse_build), which evaluates the same
// sums.  Applying the operator is the generic stencil kernel with nc = nc_c (5-point: clover + 4 hopping blocks).
#ifndef QMG_B200_COARSE
#define QMG_B200_COARSE

#include <vector>
#include "../stencil/stencil_2d.h"
#include "../transfer/transfer.h"

// apply_sigma flavours that only exist on a coarse operator whose transfer kept its block factorisations
// (/root/reference/operators/coarse.h:19-25); numbering continues QMGSigmaType
enum QMGSigmaTypeCoarse
{
  QMG_SIGMA_1_L = 6,      // left-apply  L^dag sigma_1 U^-1
  QMG_SIGMA_1_R = 7,      // right-apply U sigma_1 L^-dag
  QMG_SIGMA_1_L_RBJ = 8,  // B^-dag sigma_1^L   (coarsened rbjacobi system)
  QMG_SIGMA_1_R_RBJ = 9,  // B sigma_1^R
};

struct CoarseOperator2D : public Stencil2D
{
protected:
  CoarseOperator2D(CoarseOperator2D const&);
  CoarseOperator2D& operator=(CoarseOperator2D const&);

  Lattice2D* fine_lat;
  bool is_chiral;
  bool use_rbjacobi;
  TransferMG* in_transfer;
  QMGDefaultChirality default_chirality;
  complex<double>* sigma_1_L;   // V nc nc, built on first use by apply_sigma
  complex<double>* sigma_1_R;

  // sigma_1^L = L^dag sigma_1 U^-1 and sigma_1^R = U sigma_1 L^-dag per coarse site, from the factors the transfer saved
  // while (bi-)orthonormalising its blocks: P_raw = P_ortho U, R_raw = R_ortho L^dag (coarse.h:673-731, 771-846).
  // With R = P^dag both factors are the Cholesky factor Sigma and the two matrices coincide.
  void build_sigma_matrices()
  {
    const long cm = lat->get_size_cm(); const long V = lat->get_volume(); const int nc = lat->get_nc();
    complex<double>* U = allocate_vector<complex<double> >(cm);
    complex<double>* Ldag = allocate_vector<complex<double> >(cm);
    if (in_transfer->is_symmetric()) { in_transfer->copy_cholesky(U); copy_vector(Ldag, U, cm); }
    else { in_transfer->copy_LU(Ldag, U); cMATconjtrans_square(Ldag, V, nc); }
    // the sigma_1 block (swap of the two dof halves), repeated over the sites
    std::vector<complex<double> > one_site(nc * nc, 0.0);
    for (int i = 0; i < nc; i++) one_site[i * nc + (i < nc / 2 ? i + nc / 2 : i - nc / 2)] = 1.0;
    complex<double>* s1 = allocate_vector<complex<double> >(cm);
    zero_vector(s1, cm);
    capx_pattern(one_site.data(), nc * nc, s1, V);
    complex<double>* inv = allocate_vector<complex<double> >(cm);
    complex<double>* tmp = allocate_vector<complex<double> >(cm);
    sigma_1_L = allocate_vector<complex<double> >(cm);
    sigma_1_R = allocate_vector<complex<double> >(cm);
    cMATinverse_square(U, inv, V, nc);
    cMATxtMATyMATz_square(Ldag, s1, tmp, V, nc);
    cMATxtMATyMATz_square(tmp, inv, sigma_1_L, V, nc);
    cMATinverse_square(Ldag, inv, V, nc);
    cMATxtMATyMATz_square(U, s1, tmp, V, nc);
    cMATxtMATyMATz_square(tmp, inv, sigma_1_R, V, nc);
    deallocate_vector(&tmp); deallocate_vector(&inv); deallocate_vector(&s1); deallocate_vector(&Ldag); deallocate_vector(&U);
  }

  // per-site dof map out[s][i] = scale[i] in[s][pick[i]]; top half of the dof is "up", bottom half "down"
  void half_map(double s_top, double s_bot, bool swap_halves, complex<double>* in, complex<double>* out)
  {
    const int nc = lat->get_nc();
    std::vector<double> scale(nc); std::vector<int> pick(nc);
    for (int i = 0; i < nc; i++)
    {
      const bool top = i < nc / 2;
      scale[i] = top ? s_top : s_bot;
      pick[i] = swap_halves ? (top ? i + nc / 2 : i - nc / 2) : i;
    }
    caxy_shuffle_pattern(scale.data(), pick.data(), nc, in, out, lat->get_volume());
  }

public:
  enum QMGCoarseBuildStencil
  {
    QMG_COARSE_BUILD_ORIGINAL = 0,
    QMG_COARSE_BUILD_DAGGER = 1,
    QMG_COARSE_BUILD_RBJACOBI = 2,
    QMG_COARSE_BUILD_DAGGER_RBJACOBI = 3,
    QMG_COARSE_BUILD_RBJDAGGER = 4,
    QMG_COARSE_BUILD_ALL = 5,
  };

  // a bare stencil whose blocks the caller fills (coarse.h:76)
  CoarseOperator2D(Lattice2D* in_lat, int pieces, bool is_chiral, QMGDefaultChirality def_chiral = QMG_CHIRALITY_NONE,
                   complex<double> in_shift = 0.0, complex<double> in_eo_shift = 0.0, complex<double> in_dof_shift = 0.0)
    : Stencil2D(in_lat, pieces, in_shift, in_eo_shift, in_dof_shift), fine_lat(0), is_chiral(is_chiral), use_rbjacobi(false),
      in_transfer(0), default_chirality(def_chiral), sigma_1_L(0), sigma_1_R(0)
  { }

  // Galerkin build (coarse.h:90-471)
  CoarseOperator2D(Lattice2D* in_lat, Stencil2D* fine_stencil, Lattice2D* fine_lattice, TransferMG* transfer, bool is_chiral = false,
                   bool use_rbjacobi = false, QMGCoarseBuildStencil build_extra = QMG_COARSE_BUILD_ORIGINAL)
    : Stencil2D(in_lat, QMG_PIECE_CLOVER_HOPPING, 0.0, 0.0, 0.0), fine_lat(fine_lattice), is_chiral(is_chiral), use_rbjacobi(use_rbjacobi),
      in_transfer(transfer), sigma_1_L(0), sigma_1_R(0)
  {
    switch (transfer->get_doubling())
    {
      case QMG_DOUBLE_PROJECTION: default_chirality = QMG_CHIRALITY_GAMMA_5; break;
      case QMG_DOUBLE_OPERATOR: default_chirality = QMG_CHIRALITY_SIGMA_1; break;
      default: default_chirality = QMG_CHIRALITY_NONE; break;
    }
    // coarsen the right-block-Jacobi system instead of the original one: select that link set on the fine stencil
    if (use_rbjacobi) fine_stencil->perform_swap_rbjacobi();
    // only the identity shift is carried over, read while the selection above is active (coarse.h:129-131)
    update_shift(fine_stencil->get_shift());
    qmg_stencil_desc fd = fine_stencil->describe();
    QMG_CHK(qmg_coarse_build(transfer->get_desc(), &fd,
                             reinterpret_cast<const qmg_cplx* const*>(transfer->null_vectors),
                             reinterpret_cast<const qmg_cplx* const*>(transfer->restrict_null_vectors),
                             qmg_host::P(clover), qmg_host::P(hopping)));
    if (use_rbjacobi) fine_stencil->perform_swap_rbjacobi();
    generated = true;

    if (build_extra == QMG_COARSE_BUILD_DAGGER || build_extra == QMG_COARSE_BUILD_DAGGER_RBJACOBI || build_extra == QMG_COARSE_BUILD_ALL)
      build_dagger_stencil();
    if (build_extra == QMG_COARSE_BUILD_RBJACOBI || build_extra == QMG_COARSE_BUILD_DAGGER_RBJACOBI ||
        build_extra == QMG_COARSE_BUILD_RBJDAGGER || build_extra == QMG_COARSE_BUILD_ALL)
      build_rbjacobi_stencil();
    if (build_extra == QMG_COARSE_BUILD_RBJDAGGER || build_extra == QMG_COARSE_BUILD_ALL)
      build_rbj_dagger_stencil();
  }
  ~CoarseOperator2D()
  {
    if (sigma_1_L != 0) deallocate_vector(&sigma_1_L);
    if (sigma_1_R != 0) deallocate_vector(&sigma_1_R);
  }

  static int get_dof(int i = 0) { (void)i; return -1; }
  static chirality_state has_chirality() { return QMG_CHIRAL_UNKNOWN; }

  // gamma_5 = +1 on the top half of the coarse dof, -1 on the bottom half; nothing happens on a non-chiral operator (coarse.h:498-524)
  virtual void gamma5(complex<double>* vec) { if (is_chiral) half_map(1.0, -1.0, false, vec, vec); }
  virtual void gamma5(complex<double>* g5_vec, complex<double>* vec) { if (is_chiral) half_map(1.0, -1.0, false, vec, g5_vec); }
  // sigma_1 swaps the two halves (coarse.h:527-562)
  virtual void sigma1(complex<double>* vec)
  {
    if (lat->get_nc() % 2) return;
    complex<double>* tmp = scratch_extra();
    half_map(1.0, 1.0, true, vec, tmp);
    copy_vector(vec, tmp, lat->get_size_cv());
  }
  virtual void sigma1(complex<double>* s1_vec, complex<double>* vec) { if (lat->get_nc() % 2 == 0) half_map(1.0, 1.0, true, vec, s1_vec); }

  virtual void chiral_projection(complex<double>* vector, bool is_up)
  {
    if (!is_chiral) return;
    if (default_chirality == QMG_CHIRALITY_GAMMA_5) half_map(is_up ? 1.0 : 0.0, is_up ? 0.0 : 1.0, false, vector, vector);
    else if (default_chirality == QMG_CHIRALITY_SIGMA_1)
    {
      complex<double>* tmp = scratch_extra();
      sigma1(tmp, vector);
      caxpby(is_up ? 0.5 : -0.5, tmp, 0.5, vector, lat->get_size_cv());
    }
  }
  virtual void chiral_projection_copy(complex<double>* orig, complex<double>* dest, bool is_up)
  {
    if (!is_chiral) return;
    if (default_chirality == QMG_CHIRALITY_GAMMA_5) half_map(is_up ? 1.0 : 0.0, is_up ? 0.0 : 1.0, false, orig, dest);
    else if (default_chirality == QMG_CHIRALITY_SIGMA_1)
    {
      complex<double>* tmp = scratch_extra();
      sigma1(tmp, orig);
      caxpbyz(is_up ? 0.5 : -0.5, tmp, 0.5, orig, dest, lat->get_size_cv());
    }
  }
  virtual void chiral_projection_both(complex<double>* orig_to_up, complex<double>* down)
  {
    if (!is_chiral) return;
    if (default_chirality == QMG_CHIRALITY_GAMMA_5)
    {
      half_map(0.0, 1.0, false, orig_to_up, down);
      half_map(1.0, 0.0, false, orig_to_up, orig_to_up);
    }
    else if (default_chirality == QMG_CHIRALITY_SIGMA_1)
    {
      complex<double>* tmp = scratch_extra();
      sigma1(tmp, orig_to_up);
      caxpbyz(0.5, orig_to_up, -0.5, tmp, down, lat->get_size_cv());
      caxpy(-1.0, down, orig_to_up, lat->get_size_cv());
    }
  }
  virtual QMGDefaultChirality get_default_chirality() { return default_chirality; }
  using Stencil2D::apply_sigma;

  // sigma_1 as seen through the block factorisations of the transfer (coarse.h:661-894)
  void apply_sigma(complex<double>* output, complex<double>* input, QMGSigmaTypeCoarse type)
  {
    if (in_transfer == 0 || !in_transfer->has_decompositions())
    {
      std::cout << "[QMG-ERROR]: In CoarseOperator2D, cannot apply apply_sigma() if the transfer op does not have factorizations.\n";
      return;
    }
    if (sigma_1_L == 0 || sigma_1_R == 0) build_sigma_matrices();
    const long V = lat->get_volume(); const int nc = lat->get_nc();
    switch (type)
    {
      case QMG_SIGMA_1_L: cMATxy(sigma_1_L, input, output, V, nc, nc); break;
      case QMG_SIGMA_1_R: cMATxy(sigma_1_R, input, output, V, nc, nc); break;
      case QMG_SIGMA_1_L_RBJ:
        if (!built_rbj_dagger)
        {
          std::cout << "[QMG-ERROR]: In apply_sigma, cannot apply QMG_SIGMA_1_L_RBJ without rbjacobi dagger stencil.\n";
          copy_vector(output, input, lat->get_size_cv());
        }
        else
        {
          complex<double>* tmp = scratch_extra();
          cMATxy(sigma_1_L, input, tmp, V, nc, nc);
          cMATxy(rbj_dagger_cinv, tmp, output, V, nc, nc);
        }
        break;
      case QMG_SIGMA_1_R_RBJ:
        if (!built_rbjacobi)
        {
          std::cout << "[QMG-ERROR]: In apply_sigma, cannot apply QMG_SIGMA_1_R_RBJ without rbjacobi stencil.\n";
          copy_vector(output, input, lat->get_size_cv());
        }
        else
        {
          complex<double>* tmp = scratch_extra();
          cMATxy(sigma_1_R, input, tmp, V, nc, nc);
          cMATxy(clover, tmp, output, V, nc, nc);
          caxpy(shift, tmp, output, lat->get_size_cv());
        }
        break;
    }
  }
};

#endif
