// quantum-mg on B200 -- Stencil2D: the generic nearest-neighbour 2D stencil, with the class API of
// /root/reference/stencil/stencil_2d.h (public members :148-210, methods :339-2566, C wrappers :2571-2716).
//
//   out(x) = [clover(x) + shift +- eo_shift +- dof_shift] in(x) + sum_{mu in +x,+y,-x,-y} hopping_mu(x) in(x + mu)
//
// Every array member is a DEVICE pointer in the reference's even-odd layouts; every apply is one launch
// of the fused sm_100a stencil kernel (qmg_stencil_apply, include/qmg_b200.h) instead of the reference's
// 1 zero + 1 clover + 8 cshift + 8 half-volume matvec + 2 axpy sweeps.  The dagger / right-block-Jacobi /
// rbj-dagger variants are second, third and fourth stored link sets selected by the same pointer swaps
// the reference performs (perform_swap_*), so code that swaps by hand (operators/coarse.h:120-123) behaves alike.
// Not thread-safe per object (swaps + shared scratch), like the reference.
#ifndef QMG_B200_STENCIL_2D
#define QMG_B200_STENCIL_2D

#include <complex>
#include <cstdlib>
#include <iostream>
#include <utility>
#include "../lattice/lattice.h"
#include "blas/generic_vector.h"
#include "blas/generic_local_matrix.h"
#include "blas/generic_matrix.h"
#include "../cshift/cshift_2d.h"
#include "inverters/inverter_struct.h"

// offset of direction mu inside a hopping array is mu * lat->get_size_cm()
enum stencil_dir_index
{
  QMG_DIR_INDEX_0 = 0,
  QMG_DIR_INDEX_XP1 = 0, QMG_DIR_INDEX_YP1 = 1, QMG_DIR_INDEX_XM1 = 2, QMG_DIR_INDEX_YM1 = 3,
  QMG_DIR_INDEX_XP2 = 0, QMG_DIR_INDEX_YP2 = 1, QMG_DIR_INDEX_XM2 = 2, QMG_DIR_INDEX_YM2 = 3,
  QMG_DIR_INDEX_XP1YP1 = 0, QMG_DIR_INDEX_XM1YP1 = 1, QMG_DIR_INDEX_XM1YM1 = 2, QMG_DIR_INDEX_XP1YM1 = 3,
};

enum stencil_pieces
{
  QMG_PIECE_CLOVER = 1, QMG_PIECE_HOPPING = 2, QMG_PIECE_TWOLINK = 4, QMG_PIECE_CORNER = 8,
  QMG_PIECE_CLOVER_HOPPING = 3, QMG_PIECE_TWOLINK_CORNER = 12, QMG_PIECE_ALL = 15,
};

enum chirality_state { QMG_CHIRAL_NO = 0, QMG_CHIRAL_YES = 1, QMG_CHIRAL_UNKNOWN = 2 };

enum QMGStencilType
{
  QMG_MATVEC_ORIGINAL = 0,       // M
  QMG_MATVEC_DAGGER = 1,         // M^dagger
  QMG_MATVEC_RIGHT_JACOBI = 2,   // M B^-1, B = clover + shifts
  QMG_MATVEC_RIGHT_SCHUR = 3,    // even-odd Schur complement of M B^-1
  QMG_MATVEC_M_MDAGGER = 4,
  QMG_MATVEC_MDAGGER_M = 5,
  QMG_MATVEC_RBJ_DAGGER = 6,     // (M B^-1)^dagger
  QMG_MATVEC_RBJ_M_MDAGGER = 7,
  QMG_MATVEC_RBJ_MDAGGER_M = 8,
};

enum QMGDefaultChirality { QMG_CHIRALITY_NONE = 0, QMG_CHIRALITY_GAMMA_5 = 1, QMG_CHIRALITY_SIGMA_1 = 2 };

enum QMGSigmaType
{
  QMG_SIGMA_NONE = 0, QMG_SIGMA_DEFAULT = 1, QMG_GAMMA_5 = 2, QMG_SIGMA_1 = 3,
  QMG_GAMMA_5_L_RBJ = 4,   // B gamma_5 -- (rbjacobi only)
  QMG_GAMMA_5_R_RBJ = 5,   // gamma_5 B^-1
};

// matrix_op_cplx instances: lhs = M_type rhs on device vectors, lhs overwritten.
inline void apply_stencil_2D_M(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_piece_clover(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_piece_hopping(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_dagger(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_dagger_M(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_M_dagger(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_rbjacobi(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_rbjacobi_cinv(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_rbjacobi_schur(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_rbj_dagger(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_rbjacobi_MMD(complex<double>* lhs, complex<double>* rhs, void* extra_data);
inline void apply_stencil_2D_M_rbjacobi_MDM(complex<double>* lhs, complex<double>* rhs, void* extra_data);

struct Stencil2D
{
protected:
  Stencil2D(Stencil2D const&);
  Stencil2D& operator=(Stencil2D const&);

  // scratch, allocated on first use (the reference pre-allocates one matrix field and two vectors per stencil)
  complex<double>* extra_cvector;
  complex<double>* eo_cvector;

  std::complex<double> shift_backup, eo_shift_backup, dof_shift_backup;
  bool swap_dagger, swap_rbjacobi, swap_rbj_dagger;
  // link-compressed apply (see enable_gamma5_hermitian_apply)
  bool gamma5_hermitian;
  complex<double>* herm_halo_ym;
  // matrix-free apply (Wilson2D::enable_matrix_free_apply): the operator's own copy of the gauge links its stored blocks were
  // filled from, row -1 of U_y on a y-slab, and the Wilson parameter; 0 = read the stored blocks
  complex<double>* mf_gauge;
  complex<double>* mf_gauge_halo_ym;
  double mf_w;
  bool mf_paused;

  complex<double>* scratch_extra() { if (extra_cvector == 0) extra_cvector = allocate_vector<complex<double> >(lat->get_size_cv()); return extra_cvector; }
  complex<double>* scratch_eo() { if (eo_cvector == 0) eo_cvector = allocate_vector<complex<double> >(lat->get_size_cv()); return eo_cvector; }

  static bool need(bool ok, const char* fn, const char* what)
  {
    if (!ok) std::cout << "[QMG-WARNING]: Tried to call " << fn << ", but " << what << ".\n";
    return ok;
  }
  static void unsupported_pieces(const complex<double>* two, const complex<double>* cor)
  {
    if (two != 0) std::cout << "[QMG-WARNING]: two link stencil not yet supported.\n";
    if (cor != 0) std::cout << "[QMG-WARNING]: corner stencil not yet supported.\n";
  }

public:
  // descriptor of the CURRENTLY selected link set for the kernel layer
  qmg_stencil_desc describe(const complex<double>* cl, const complex<double>* hp, complex<double> s, complex<double> eo, complex<double> dof) const
  {
    qmg_stencil_desc d;
    d.X = lat->get_dim_mu(0); d.Y = lat->get_dim_mu(1); d.nc = lat->get_nc();
    d.clover = qmg_host::P(cl); d.hopping = qmg_host::P(hp);
    d.shift[0] = s.real(); d.shift[1] = s.imag();
    d.eo_shift[0] = eo.real(); d.eo_shift[1] = eo.imag();
    d.dof_shift[0] = dof.real(); d.dof_shift[1] = dof.imag();
    d.halo_ym = 0; d.halo_yp = 0;
    // the link-compressed apply is only valid for the ORIGINAL link set (not while a variant is swapped in)
    const bool original = !swap_dagger && !swap_rbjacobi && !swap_rbj_dagger && (cl == clover || cl == 0) && hp == hopping && hp != 0;
    d.gamma5_hermitian = (gamma5_hermitian && original) ? 1 : 0;
    d.hop_halo_ym = d.gamma5_hermitian ? qmg_host::P(herm_halo_ym) : 0;
    // ... and so is the matrix-free one (it also needs the clover of the set: the kernel layer takes it for whole-operator applies only)
    const bool mf = (mf_gauge != 0) && !mf_paused && original && cl == clover && cl != 0;
    d.wilson_gauge = mf ? qmg_host::P(mf_gauge) : 0;
    d.wilson_w = mf ? mf_w : 0.0;
    d.wilson_gauge_halo_ym = mf ? qmg_host::P(mf_gauge_halo_ym) : 0;
    return d;
  }
  qmg_stencil_desc describe() const { return describe(clover, hopping, shift, eo_shift, dof_shift); }

  // one fused kernel launch on the currently selected link set
  void launch(int pieces, int dir_mask, complex<double>* lhs, complex<double>* rhs)
  {
    qmg_stencil_desc d = describe();
    if (clover == 0) pieces &= ~QMG_APPLY_CLOVER;
    if (hopping == 0) pieces &= ~(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD);
    QMG_CHK(qmg_stencil_apply(&d, pieces, dir_mask, qmg_host::P(lhs), qmg_host::P(rhs)));
  }

  Lattice2D* lat;
  complex<double>* clover;    // V nc nc
  complex<double>* hopping;   // 4 V nc nc, {+x,+y,-x,-y}
  complex<double>* twolink;   // declared by the reference, never applied (stencil_2d.h:925-933)
  complex<double>* corner;
  bool generated;
  complex<double> shift;      // identity
  complex<double> eo_shift;   // + on even sites, - on odd sites
  complex<double> dof_shift;  // + on the top half of the dof, - on the bottom half

  bool built_dagger;
  complex<double>* dagger_clover; complex<double>* dagger_hopping; complex<double>* dagger_twolink; complex<double>* dagger_corner;
  bool built_rbjacobi;
  complex<double>* rbjacobi_clover; complex<double>* rbjacobi_hopping; complex<double>* rbjacobi_twolink; complex<double>* rbjacobi_corner;
  complex<double>* rbjacobi_cinv;
  bool built_rbj_dagger;
  complex<double>* rbj_dagger_clover; complex<double>* rbj_dagger_hopping; complex<double>* rbj_dagger_twolink; complex<double>* rbj_dagger_corner;
  complex<double>* rbj_dagger_cinv;

  Stencil2D(Lattice2D* in_lat, int pieces, complex<double> in_shift = 0.0, complex<double> in_eo_shift = 0.0, complex<double> in_dof_shift = 0.0)
    : extra_cvector(0), eo_cvector(0), lat(in_lat), clover(0), hopping(0), twolink(0), corner(0), generated(false),
      shift(in_shift), eo_shift(in_eo_shift), dof_shift(in_dof_shift),
      built_dagger(false), dagger_clover(0), dagger_hopping(0), dagger_twolink(0), dagger_corner(0),
      built_rbjacobi(false), rbjacobi_clover(0), rbjacobi_hopping(0), rbjacobi_twolink(0), rbjacobi_corner(0), rbjacobi_cinv(0),
      built_rbj_dagger(false), rbj_dagger_clover(0), rbj_dagger_hopping(0), rbj_dagger_twolink(0), rbj_dagger_corner(0), rbj_dagger_cinv(0)
  {
    if (pieces & QMG_PIECE_CLOVER) clover = allocate_vector<complex<double> >(lat->get_size_cm());
    if (pieces & QMG_PIECE_HOPPING) hopping = allocate_vector<complex<double> >(lat->get_size_hopping());
    if (pieces & QMG_PIECE_TWOLINK) twolink = allocate_vector<complex<double> >(lat->get_size_hopping());
    if (pieces & QMG_PIECE_CORNER) corner = allocate_vector<complex<double> >(lat->get_size_corner());
    shift_backup = shift; eo_shift_backup = eo_shift; dof_shift_backup = dof_shift;
    swap_dagger = swap_rbjacobi = swap_rbj_dagger = false;
    gamma5_hermitian = false; herm_halo_ym = 0;
    mf_gauge = 0; mf_gauge_halo_ym = 0; mf_w = 0.0; mf_paused = false;
  }

  // B200 extension (not in the reference): for an operator with D^dag = gamma5 D gamma5 -- Wilson2D and the Galerkin
  // coarsenings built from chirality-doubled null vectors -- the stored backward blocks repeat the neighbours' forward
  // blocks, hopping_{-mu}(x) = gamma5 hopping_{+mu}(x - mu)^dag gamma5.  After this call apply_M on the ORIGINAL link set
  // reads 3 of the 5 blocks per site (clover, +x, +y) and takes the backward hops out of L2.  The relation is CHECKED on
  // the stored blocks first (relative deviation <= tol, else nothing changes and false is returned); anything that edits
  // the links (update_links, clear_stencils, writing through the public pointers) must be followed by a new call.
  bool enable_gamma5_hermitian_apply(double tol = 1e-12)
  {
    gamma5_hermitian = false;
    const int nc = lat->get_nc();
    if (hopping == 0 || nc % 2 != 0 || nc > 32 || swap_dagger || swap_rbjacobi || swap_rbj_dagger) return false;
    {
      // QMG_HERM_MIN_NC: smallest dof count that switches (experiments: nc = 2 has no tile kernel and gains little)
      const char* e = getenv("QMG_HERM_MIN_NC");
      if (e != 0 && nc < atoi(e)) return false;
    }
    qmg_stencil_desc d = describe();
    double dev[2] = {0.0, 0.0};
    QMG_CHK(qmg_stencil_gamma5_deviation(&d, dev));
    if (!(dev[1] > 0.0) || sqrt(dev[0] / dev[1]) > tol) return false;
    if (qmg_comm_active())
    {
      // on a y-slab row -1 of the +y blocks lives on the lower rank: fetched once
      const long row = (long)lat->get_dim_mu(0) * nc * nc;
      if (herm_halo_ym == 0) herm_halo_ym = allocate_vector<complex<double> >(row);
      complex<double>* unused = allocate_vector<complex<double> >(row);
      QMG_CHK(qmg_halo_exchange(qmg_host::P(hopping + lat->get_size_cm()), lat->get_dim_mu(0), lat->get_dim_mu(1), nc * nc,
                                qmg_host::P(herm_halo_ym), qmg_host::P(unused)));
      deallocate_vector(&unused);
    }
    gamma5_hermitian = true;
    return true;
  }
  void disable_gamma5_hermitian_apply() { gamma5_hermitian = false; }
  bool uses_gamma5_hermitian_apply() const { return gamma5_hermitian; }

  // drop the matrix-free apply (the stored blocks are read again)
  void disable_matrix_free_apply()
  {
    if (mf_gauge != 0) deallocate_vector(&mf_gauge);
    if (mf_gauge_halo_ym != 0) deallocate_vector(&mf_gauge_halo_ym);
  }
  bool uses_matrix_free_apply() const { return mf_gauge != 0 && !mf_paused; }
  // read the stored blocks for a while without giving the gauge copy up (returns whether a matrix-free apply is set up at all)
  bool pause_matrix_free_apply(bool pause) { mf_paused = pause; return mf_gauge != 0; }

  virtual ~Stencil2D()
  {
    if (herm_halo_ym != 0) deallocate_vector(&herm_halo_ym);
    disable_matrix_free_apply();
    complex<double>** all[] = { &clover, &hopping, &twolink, &corner, &extra_cvector, &eo_cvector,
                                &dagger_clover, &dagger_hopping, &dagger_twolink, &dagger_corner,
                                &rbjacobi_clover, &rbjacobi_hopping, &rbjacobi_twolink, &rbjacobi_corner, &rbjacobi_cinv,
                                &rbj_dagger_clover, &rbj_dagger_hopping, &rbj_dagger_twolink, &rbj_dagger_corner, &rbj_dagger_cinv };
    for (unsigned i = 0; i < sizeof(all) / sizeof(all[0]); i++) if (*all[i] != 0) deallocate_vector(all[i]);
    built_dagger = built_rbjacobi = built_rbj_dagger = generated = false;
  }

  // drop the derived link sets (after the links changed: operators/wilson.h:211-225)
  void free_derived_stencils()
  {
    complex<double>** der[] = { &dagger_clover, &dagger_hopping, &rbjacobi_clover, &rbjacobi_hopping, &rbjacobi_cinv };
    for (unsigned i = 0; i < sizeof(der) / sizeof(der[0]); i++) if (*der[i] != 0) deallocate_vector(der[i]);
    built_dagger = built_rbjacobi = false;
    gamma5_hermitian = false;      // the links changed: the relation has to be re-checked
    disable_matrix_free_apply();   // ... and the gauge copy no longer describes the blocks
  }

  // stencil_2d.h:339-376 (including the reference's reset of built_rbjacobi where built_rbj_dagger is meant)
  void clear_stencils()
  {
    const long cm = lat->get_size_cm(), hp = lat->get_size_hopping(), cr = lat->get_size_corner();
    gamma5_hermitian = false;
    disable_matrix_free_apply();
    if (clover != 0) zero_vector(clover, cm);
    if (hopping != 0) zero_vector(hopping, hp);
    if (twolink != 0) zero_vector(twolink, hp);
    if (corner != 0) zero_vector(corner, cr);
    if (built_dagger)
    {
      if (dagger_clover != 0) zero_vector(dagger_clover, cm);
      if (dagger_hopping != 0) zero_vector(dagger_hopping, hp);
      built_dagger = false;
    }
    if (built_rbjacobi)
    {
      if (rbjacobi_clover != 0) zero_vector(rbjacobi_clover, cm);
      if (rbjacobi_hopping != 0) zero_vector(rbjacobi_hopping, hp);
      if (rbjacobi_cinv != 0) zero_vector(rbjacobi_cinv, cm);
      built_rbjacobi = false;
    }
    if (built_rbj_dagger)
    {
      if (rbj_dagger_clover != 0) zero_vector(rbj_dagger_clover, cm);
      if (rbj_dagger_hopping != 0) zero_vector(rbj_dagger_hopping, hp);
      built_rbjacobi = false;
    }
    generated = false;
  }

  void prune_stencils(int pieces)
  {
    if ((pieces & QMG_PIECE_CLOVER) && clover != 0) deallocate_vector(&clover);
    if ((pieces & QMG_PIECE_HOPPING) && hopping != 0) deallocate_vector(&hopping);
    if ((pieces & QMG_PIECE_TWOLINK) && twolink != 0) deallocate_vector(&twolink);
    if ((pieces & QMG_PIECE_CORNER) && corner != 0) deallocate_vector(&corner);
    if (clover == 0 && hopping == 0 && twolink == 0 && corner == 0) generated = false;
  }
  void try_prune_stencils(int pieces, double tol)
  {
    int drop = 0;
    if ((pieces & QMG_PIECE_CLOVER) && clover != 0 && norminf(clover, lat->get_size_cm()) < tol) drop |= QMG_PIECE_CLOVER;
    if ((pieces & QMG_PIECE_HOPPING) && hopping != 0 && norminf(hopping, lat->get_size_hopping()) < tol) drop |= QMG_PIECE_HOPPING;
    if ((pieces & QMG_PIECE_TWOLINK) && twolink != 0 && norminf(twolink, lat->get_size_hopping()) < tol) drop |= QMG_PIECE_TWOLINK;
    if ((pieces & QMG_PIECE_CORNER) && corner != 0 && norminf(corner, lat->get_size_corner()) < tol) drop |= QMG_PIECE_CORNER;
    prune_stencils(drop);
  }

  Lattice2D* get_lattice() { return lat; }
  complex<double>* expose_internal_cvector() { return scratch_extra(); }

  // host-side inspection of the blocks at one site (downloads nc*nc values per block)
  void print_stencil_site(int x, int y, string prefix = "")
  {
    const int nc = lat->get_nc();
    std::vector<complex<double> > blk((size_t)nc * nc);
    struct { const char* name; complex<double>* base; int mu; } rows[] = {
      { "Clover", clover, 0 }, { "Hopping +x", hopping, 0 }, { "Hopping +y", hopping, 1 }, { "Hopping -x", hopping, 2 }, { "Hopping -y", hopping, 3 } };
    std::cout << prefix << "Shift " << shift << "\n" << prefix << "Eo_shift " << eo_shift << "\n" << prefix << "Dof_shift " << dof_shift << "\n";
    for (int r = 0; r < 5; r++)
    {
      if (rows[r].base == 0) continue;
      qmg_host::download(blk.data(), rows[r].base + lat->hopping_coord_to_index(x, y, 0, 0, rows[r].mu), (long)nc * nc);
      std::cout << prefix << rows[r].name << "\n";
      for (int i = 0; i < nc; i++) { std::cout << prefix; for (int j = 0; j < nc; j++) std::cout << blk[i * nc + j] << " "; std::cout << "\n"; }
    }
  }

  void update_shifts(complex<double> s, complex<double> eo, complex<double> dof) { shift = shift_backup = s; eo_shift = eo_shift_backup = eo; dof_shift = dof_shift_backup = dof; }
  void update_shift(complex<double> s) { shift = shift_backup = s; }
  void update_eo_shift(complex<double> s) { eo_shift = eo_shift_backup = s; }
  void update_dof_shift(complex<double> s) { dof_shift = dof_shift_backup = s; }
  complex<double> get_shift() { return shift; }
  complex<double> get_shift_eo() { return eo_shift; }
  complex<double> get_shift_dof() { return dof_shift; }

  // ---- the accumulate-into-lhs pieces (stencil_2d.h:666-936): lhs += piece * rhs
  // ee / oo: clover block plus `shift` only (not eo/dof shift), and only when a clover exists (:666-692)
  void apply_M_ee(complex<double>* lhs, complex<double>* rhs)
  {
    if (clover == 0) return;
    qmg_stencil_desc d = describe(clover, 0, shift, 0.0, 0.0);
    QMG_CHK(qmg_stencil_apply(&d, QMG_APPLY_CLOVER | QMG_APPLY_SHIFT | QMG_APPLY_ACCUMULATE | QMG_APPLY_EVEN_ROWS_ONLY, 15, qmg_host::P(lhs), qmg_host::P(rhs)));
  }
  void apply_M_oo(complex<double>* lhs, complex<double>* rhs)
  {
    if (clover == 0) return;
    qmg_stencil_desc d = describe(clover, 0, shift, 0.0, 0.0);
    QMG_CHK(qmg_stencil_apply(&d, QMG_APPLY_CLOVER | QMG_APPLY_SHIFT | QMG_APPLY_ACCUMULATE | QMG_APPLY_ODD_ROWS_ONLY, 15, qmg_host::P(lhs), qmg_host::P(rhs)));
  }
  void apply_M_clover(complex<double>* lhs, complex<double>* rhs) { if (clover != 0) launch(QMG_APPLY_CLOVER | QMG_APPLY_ACCUMULATE, 15, lhs, rhs); }
  // eo: hopping into the EVEN rows (reads the odd half of rhs); in place is allowed, as in staggered.h:218
  void apply_M_eo(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(hopping != 0, "'apply_M_eo'", "there is no hopping term")) return;
    launch(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_ACCUMULATE | QMG_APPLY_EVEN_ROWS_ONLY, 15, lhs, rhs);
  }
  void apply_M_eo(complex<double>* lhs, complex<double>* rhs, stencil_dir_index dir)
  {
    if (!need(hopping != 0, "'apply_M_eo'", "there is no hopping term")) return;
    launch(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_ACCUMULATE | QMG_APPLY_EVEN_ROWS_ONLY, 1 << (int)dir, lhs, rhs);
  }
  void apply_M_oe(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(hopping != 0, "'apply_M_oe'", "there is no hopping term")) return;
    launch(QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ACCUMULATE | QMG_APPLY_ODD_ROWS_ONLY, 15, lhs, rhs);
  }
  void apply_M_oe(complex<double>* lhs, complex<double>* rhs, stencil_dir_index dir)
  {
    if (!need(hopping != 0, "'apply_M_oe'", "there is no hopping term")) return;
    launch(QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ACCUMULATE | QMG_APPLY_ODD_ROWS_ONLY, 1 << (int)dir, lhs, rhs);
  }
  void apply_M_hopping(complex<double>* lhs, complex<double>* rhs)
  { if (hopping != 0) launch(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ACCUMULATE, 15, lhs, rhs); }
  void apply_M_hopping(complex<double>* lhs, complex<double>* rhs, stencil_dir_index dir)
  { if (hopping != 0) launch(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ACCUMULATE, 1 << (int)dir, lhs, rhs); }
  void apply_M_shift(complex<double>* lhs, complex<double>* rhs) { launch(QMG_APPLY_SHIFT | QMG_APPLY_ACCUMULATE, 15, lhs, rhs); }
  // the whole operator, accumulating (stencil_2d.h:912)
  void apply_M(complex<double>* lhs, complex<double>* rhs)
  {
    unsupported_pieces(twolink, corner);
    launch(QMG_APPLY_ALL | QMG_APPLY_ACCUMULATE, 15, lhs, rhs);
  }
  // lhs = M rhs in one pass, nothing read from lhs (what the apply_stencil_2D_* wrappers need)
  void apply_M_overwrite(complex<double>* lhs, complex<double>* rhs) { launch(QMG_APPLY_ALL, 15, lhs, rhs); }

  // B200 extension used by the K-cycle: lhs = b - M_type rhs in ONE launch for the single-launch operator flavours
  // (original, right block Jacobi); returns false -- nothing done -- for the others, whose callers apply and subtract.
  bool apply_residual(complex<double>* lhs, complex<double>* b, complex<double>* rhs, QMGStencilType type)
  {
    if (type == QMG_MATVEC_ORIGINAL)
    {
      qmg_stencil_desc d = describe();
      int pieces = QMG_APPLY_ALL;
      if (clover == 0) pieces &= ~QMG_APPLY_CLOVER;
      if (hopping == 0) pieces &= ~(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD);
      QMG_CHK(qmg_stencil_apply_residual(&d, pieces, 15, qmg_host::P(lhs), qmg_host::P(rhs), qmg_host::P(b)));
      return true;
    }
    if (type == QMG_MATVEC_RIGHT_JACOBI && built_rbjacobi)
    {
      qmg_stencil_desc d = describe(0, rbjacobi_hopping, 0.0, 0.0, 0.0);
      int pieces = QMG_APPLY_IDENTITY_CLOVER;
      if (rbjacobi_hopping != 0) pieces |= QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD;
      QMG_CHK(qmg_stencil_apply_residual(&d, pieces, 15, qmg_host::P(lhs), qmg_host::P(rhs), qmg_host::P(b)));
      return true;
    }
    return false;
  }

  // ---- chirality (overridden by the operators)
  static int get_dof(int i = 0) { (void)i; return -1; }
  static chirality_state has_chirality() { return QMG_CHIRAL_UNKNOWN; }
  virtual void gamma5(complex<double>* vec) { (void)vec; }
  virtual void gamma5(complex<double>* g5_vec, complex<double>* vec) { copy_vector(g5_vec, vec, lat->get_size_cv()); }
  virtual void chiral_projection(complex<double>* vector, bool is_up) = 0;
  virtual void chiral_projection_copy(complex<double>* orig, complex<double>* dest, bool is_up) = 0;
  virtual void chiral_projection_both(complex<double>* orig_to_up, complex<double>* down) = 0;
  virtual void sigma1(complex<double>* vec) { (void)vec; }
  virtual void sigma1(complex<double>* s1_vec, complex<double>* vec) { copy_vector(s1_vec, vec, lat->get_size_cv()); }
  virtual QMGDefaultChirality get_default_chirality() = 0;

  // stencil_2d.h:1015-1073 (the reference's L/R error strings are swapped; messages here name the missing piece)
  void apply_sigma(complex<double>* output, complex<double>* input, QMGSigmaType type = QMG_SIGMA_DEFAULT)
  {
    const long n = lat->get_size_cv();
    switch (type)
    {
      case QMG_SIGMA_NONE: copy_vector(output, input, n); break;
      case QMG_SIGMA_DEFAULT:
        switch (get_default_chirality())
        {
          case QMG_CHIRALITY_SIGMA_1: sigma1(output, input); break;
          case QMG_CHIRALITY_GAMMA_5: gamma5(output, input); break;
          default: copy_vector(output, input, n); break;
        }
        break;
      case QMG_GAMMA_5: gamma5(output, input); break;
      case QMG_SIGMA_1: sigma1(output, input); break;
      case QMG_GAMMA_5_R_RBJ:
        if (!built_rbjacobi)
        {
          std::cout << "[QMG-ERROR]: In apply_sigma, cannot apply QMG_GAMMA_5_R_RBJ without rbjacobi stencil.\n";
          copy_vector(output, input, n);
        }
        else
        {
          complex<double>* tmp = scratch_extra();
          gamma5(tmp, input);
          qmg_stencil_desc d = describe(clover, 0, shift, 0.0, 0.0);   // clover + mass (:1054-1056)
          QMG_CHK(qmg_stencil_apply(&d, (clover ? QMG_APPLY_CLOVER : 0) | QMG_APPLY_SHIFT, 15, qmg_host::P(output), qmg_host::P(tmp)));
        }
        break;
      case QMG_GAMMA_5_L_RBJ:
        if (!built_rbj_dagger)
        {
          std::cout << "[QMG-ERROR]: In apply_sigma, cannot apply QMG_GAMMA_5_L_RBJ without rbj dagger stencil.\n";
          copy_vector(output, input, n);
        }
        else
        {
          complex<double>* tmp = scratch_extra();
          gamma5(tmp, input);
          cMATxy(rbj_dagger_cinv, tmp, output, lat->get_volume(), lat->get_nc(), lat->get_nc());
        }
        break;
    }
  }

  // ---- dagger link set: H^dag_mu(x) = [H_{-mu}(x + mu)]^dag (stencil_2d.h:1080-1139)
  void build_dagger_stencil()
  {
    if (built_dagger) { std::cout << "[QMG-WARNING]: Tried to call build_dagger_stencil, but it's already been called once.\n"; return; }
    if (clover != 0) dagger_clover = allocate_vector<complex<double> >(lat->get_size_cm());
    if (hopping != 0) dagger_hopping = allocate_vector<complex<double> >(lat->get_size_hopping());
    build_conjugate_set(clover, hopping, dagger_clover, dagger_hopping);
    unsupported_pieces(twolink, corner);
    built_dagger = true;
  }
  bool perform_swap_dagger()
  {
    if (!need(built_dagger, "perform_swap_dagger", "the dagger stencil has not been allocated")) return false;
    std::swap(clover, dagger_clover); std::swap(hopping, dagger_hopping); std::swap(twolink, dagger_twolink); std::swap(corner, dagger_corner);
    if (!swap_dagger) { shift = std::conj(shift); eo_shift = std::conj(eo_shift); dof_shift = std::conj(dof_shift); }
    else { shift = shift_backup; eo_shift = eo_shift_backup; dof_shift = dof_shift_backup; }
    swap_dagger = !swap_dagger;
    return swap_dagger;
  }
  void print_stencil_dagger_site(int x, int y, string prefix = "")
  {
    if (!need(built_dagger, "print_stencil_dagger_site", "the dagger stencil has not been allocated")) return;
    perform_swap_dagger(); print_stencil_site(x, y, prefix); perform_swap_dagger();
  }

#define QMG_FAMILY_PIECE(FN, BUILT, SWAP, WHAT, PTR, CALL)                                       \
  void FN(complex<double>* lhs, complex<double>* rhs)                                             \
  {                                                                                               \
    if (!need(BUILT, #FN, WHAT " stencil has not been allocated")) return;                        \
    if (!need(PTR != 0, #FN, WHAT " term does not exist")) return;                                \
    SWAP(); CALL(lhs, rhs); SWAP();                                                               \
  }
#define QMG_FAMILY_PIECE_DIR(FN, BUILT, SWAP, WHAT, PTR, CALL)                                   \
  void FN(complex<double>* lhs, complex<double>* rhs, stencil_dir_index dir)                      \
  {                                                                                               \
    if (!need(BUILT, #FN, WHAT " stencil has not been allocated")) return;                        \
    if (!need(PTR != 0, #FN, WHAT " term does not exist")) return;                                \
    SWAP(); CALL(lhs, rhs, dir); SWAP();                                                          \
  }
  QMG_FAMILY_PIECE(apply_M_dagger_clover, built_dagger, perform_swap_dagger, "the dagger", dagger_clover, apply_M_clover)
  QMG_FAMILY_PIECE(apply_M_dagger_ee, built_dagger, perform_swap_dagger, "the dagger", dagger_clover, apply_M_ee)
  QMG_FAMILY_PIECE(apply_M_dagger_oo, built_dagger, perform_swap_dagger, "the dagger", dagger_clover, apply_M_oo)
  QMG_FAMILY_PIECE(apply_M_dagger_eo, built_dagger, perform_swap_dagger, "the dagger", dagger_hopping, apply_M_eo)
  QMG_FAMILY_PIECE_DIR(apply_M_dagger_eo, built_dagger, perform_swap_dagger, "the dagger", dagger_hopping, apply_M_eo)
  QMG_FAMILY_PIECE(apply_M_dagger_oe, built_dagger, perform_swap_dagger, "the dagger", dagger_hopping, apply_M_oe)
  QMG_FAMILY_PIECE_DIR(apply_M_dagger_oe, built_dagger, perform_swap_dagger, "the dagger", dagger_hopping, apply_M_oe)
  QMG_FAMILY_PIECE(apply_M_dagger_hopping, built_dagger, perform_swap_dagger, "the dagger", dagger_hopping, apply_M_hopping)
  // the reference's (dir) overloads of the family hopping applies ignore dir (stencil_2d.h:1363,1798,2245); kept
  void apply_M_dagger_hopping(complex<double>* lhs, complex<double>* rhs, stencil_dir_index) { apply_M_dagger_hopping(lhs, rhs); }
  void apply_M_dagger_shift(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_dagger, "apply_M_dagger_shift", "the dagger stencil has not been allocated")) return;
    perform_swap_dagger(); apply_M_shift(lhs, rhs); perform_swap_dagger();
  }
  void apply_M_dagger(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_dagger, "apply_M_dagger", "the dagger stencil has not been allocated")) return;
    perform_swap_dagger(); apply_M(lhs, rhs); perform_swap_dagger();
  }
  // lhs += M^dag M rhs through the exposed scratch vector (stencil_2d.h:1400-1446)
  void apply_M_dagger_M(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_dagger, "apply_M_dagger_M", "the dagger stencil has not been built")) return;
    complex<double>* tmp = scratch_extra();
    apply_M_overwrite(tmp, rhs);
    apply_M_dagger(lhs, tmp);
  }
  void prepare_M_dagger_M(complex<double>* Mdagger_b, complex<double>* b)
  { if (need(built_dagger, "prepare_M_dagger_M", "the dagger stencil has not been built")) apply_M_dagger(Mdagger_b, b); }
  void apply_M_M_dagger(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_dagger, "apply_M_M_dagger", "the dagger stencil has not been built")) return;
    complex<double>* tmp = scratch_extra();
    perform_swap_dagger(); apply_M_overwrite(tmp, rhs); perform_swap_dagger();
    apply_M(lhs, tmp);
  }
  void reconstruct_M_M_dagger(complex<double>* x, complex<double>* y)
  { if (need(built_dagger, "reconstruct_M_M_dagger", "the dagger stencil has not been built")) apply_M_dagger(x, y); }

  // ---- right block Jacobi link set: cinv = B^-1, identity clover, H'_mu(x) = H_mu(x) B^-1(x + mu) (stencil_2d.h:1452-1601)
  void build_rbjacobi_stencil()
  {
    if (built_rbjacobi) { std::cout << "[QMG-WARNING]: Tried to call build_rbjacobi_stencil, but it's already been called once.\n"; return; }
    if (clover == 0 && shift == 0.0 && eo_shift == 0.0 && dof_shift == 0.0)
    {
      std::cout << "[QMG-ERROR]: Tried to call build_rbjacobi_stencil, but there is no clover term or shift.\n";
      return;
    }
    const long cm = lat->get_size_cm();
    rbjacobi_cinv = allocate_vector<complex<double> >(cm);
    rbjacobi_clover = allocate_vector<complex<double> >(cm);
    if (hopping != 0) rbjacobi_hopping = allocate_vector<complex<double> >(lat->get_size_hopping());
    if (lat->get_volume() == 1)
    {
      // one site: B = clover + (shift + eo_shift +- dof_shift); there is no hopping to rescale (:1530-1533)
      const int nc = lat->get_nc();
      std::vector<complex<double> > pat((size_t)nc * nc, 0.0), one((size_t)nc * nc, 0.0);
      for (int c = 0; c < nc; c++)
      {
        pat[c * nc + c] = shift + eo_shift + ((nc % 2 == 0) ? ((c < nc / 2) ? dof_shift : -dof_shift) : 0.0);
        one[c * nc + c] = 1.0;
      }
      if (clover != 0) copy_vector(rbjacobi_clover, clover, cm); else zero_vector(rbjacobi_clover, cm);
      capx_pattern(pat.data(), nc * nc, rbjacobi_clover, 1);
      cMATinverse_square(rbjacobi_clover, rbjacobi_cinv, 1, nc);
      qmg_host::upload(rbjacobi_clover, one.data(), (long)nc * nc);
      if (hopping != 0) zero_vector(rbjacobi_hopping, lat->get_size_hopping());
    }
    else
    {
      qmg_stencil_desc d = describe();
      QMG_CHK(qmg_build_rbjacobi(&d, qmg_host::P(rbjacobi_cinv), qmg_host::P(rbjacobi_clover), qmg_host::P(rbjacobi_hopping)));
    }
    unsupported_pieces(twolink, corner);
    built_rbjacobi = true;
  }
  bool perform_swap_rbjacobi()
  {
    if (!need(built_rbjacobi, "perform_swap_rbjacobi", "the rbjacobi stencil has not been allocated")) return false;
    std::swap(clover, rbjacobi_clover); std::swap(hopping, rbjacobi_hopping); std::swap(twolink, rbjacobi_twolink); std::swap(corner, rbjacobi_corner);
    if (!swap_rbjacobi) { shift = 0.0; eo_shift = 0.0; dof_shift = 0.0; }
    else { shift = shift_backup; eo_shift = eo_shift_backup; dof_shift = dof_shift_backup; }
    swap_rbjacobi = !swap_rbjacobi;
    return swap_rbjacobi;
  }
  void print_stencil_rbjacobi_site(int x, int y, string prefix = "")
  {
    if (!need(built_rbjacobi, "print_stencil_rbjacobi_site", "the rbjacobi stencil has not been allocated")) return;
    perform_swap_rbjacobi(); print_stencil_site(x, y, prefix); perform_swap_rbjacobi();
    const int nc = lat->get_nc();
    std::vector<complex<double> > blk((size_t)nc * nc);
    qmg_host::download(blk.data(), rbjacobi_cinv + lat->cm_coord_to_index(x, y, 0, 0), (long)nc * nc);
    std::cout << prefix << "Right Block Jacobi Inv Clover\n";
    for (int i = 0; i < nc; i++) { std::cout << prefix; for (int j = 0; j < nc; j++) std::cout << blk[i * nc + j] << " "; std::cout << "\n"; }
  }
  // the rbjacobi clover is the identity and is never read (stencil_2d.h:1685: cxpy)
  void apply_M_rbjacobi_clover(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_rbjacobi, "apply_M_rbjacobi_clover", "the rbjacobi stencil has not been allocated")) return;
    if (!need(rbjacobi_clover != 0, "apply_M_rbjacobi_clover", "the rbjacobi clover does not exist")) return;
    cxpy(rhs, lhs, lat->get_size_cv());
  }
  QMG_FAMILY_PIECE(apply_M_rbjacobi_eo, built_rbjacobi, perform_swap_rbjacobi, "the rbjacobi", rbjacobi_hopping, apply_M_eo)
  QMG_FAMILY_PIECE_DIR(apply_M_rbjacobi_eo, built_rbjacobi, perform_swap_rbjacobi, "the rbjacobi", rbjacobi_hopping, apply_M_eo)
  QMG_FAMILY_PIECE(apply_M_rbjacobi_oe, built_rbjacobi, perform_swap_rbjacobi, "the rbjacobi", rbjacobi_hopping, apply_M_oe)
  QMG_FAMILY_PIECE_DIR(apply_M_rbjacobi_oe, built_rbjacobi, perform_swap_rbjacobi, "the rbjacobi", rbjacobi_hopping, apply_M_oe)
  QMG_FAMILY_PIECE(apply_M_rbjacobi_hopping, built_rbjacobi, perform_swap_rbjacobi, "the rbjacobi", rbjacobi_hopping, apply_M_hopping)
  void apply_M_rbjacobi_hopping(complex<double>* lhs, complex<double>* rhs, stencil_dir_index) { apply_M_rbjacobi_hopping(lhs, rhs); }
  void apply_M_rbjacobi_shift(complex<double>*, complex<double>*) { }
  // lhs += (1 + H') rhs in one launch
  void apply_M_rbjacobi(complex<double>* lhs, complex<double>* rhs) { apply_M_rbjacobi_fused(lhs, rhs, true); }
  void apply_M_rbjacobi_fused(complex<double>* lhs, complex<double>* rhs, bool accumulate)
  {
    if (!need(built_rbjacobi, "apply_M_rbjacobi", "the rbjacobi stencil has not been allocated")) return;
    qmg_stencil_desc d = describe(0, rbjacobi_hopping, 0.0, 0.0, 0.0);
    int pieces = QMG_APPLY_IDENTITY_CLOVER | (accumulate ? QMG_APPLY_ACCUMULATE : 0);
    if (rbjacobi_hopping != 0) pieces |= QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD;
    QMG_CHK(qmg_stencil_apply(&d, pieces, 15, qmg_host::P(lhs), qmg_host::P(rhs)));
  }
  // lhs += B^-1 rhs
  void apply_M_rbjacobi_cinv(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_rbjacobi, "apply_M_rbjacobi_cinv", "the rbjacobi stencil has not been allocated")) return;
    if (!need(rbjacobi_cinv != 0, "apply_M_rbjacobi_cinv", "the rbjacobi cinv does not exist")) return;
    cMATxpy(rbjacobi_cinv, rhs, lhs, lat->get_volume(), lat->get_nc(), lat->get_nc());
  }
  void reconstruct_M_rbjacobi(complex<double>* x, complex<double>* y)
  { if (need(built_rbjacobi, "reconstruct_M_rbjacobi", "the rbjacobi stencil has not been allocated")) apply_M_rbjacobi_cinv(x, y); }

  // ---- even-odd Schur complement of the rbjacobi system (stencil_2d.h:1886-1983); vectors are half length
  // lhs_e = rhs_e - H'_eo H'_oe rhs_e
  void apply_M_rbjacobi_schur(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_rbjacobi, "apply_M_rbjacobi_schur", "the rbjacobi stencil has not been allocated")) return;
    complex<double>* tmp = scratch_eo();
    const long half = lat->get_size_cv() / 2;
    qmg_stencil_desc d = describe(0, rbjacobi_hopping, 0.0, 0.0, 0.0);
    QMG_CHK(qmg_stencil_apply(&d, QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ODD_ROWS_ONLY, 15, qmg_host::P(tmp), qmg_host::P(rhs)));
    QMG_CHK(qmg_stencil_apply(&d, QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_EVEN_ROWS_ONLY, 15, qmg_host::P(tmp), qmg_host::P(tmp)));
    caxpbyz(1.0, rhs, -1.0, tmp, lhs, half);
  }
  // b_r(even) = b_e - H'_eo b_o ; b_r(odd) = 0.  Accumulates the hop into b_r like the reference (:1922).
  void prepare_M_rbjacobi_schur(complex<double>* b_r, complex<double>* b)
  {
    if (!need(built_rbjacobi, "prepare_M_rbjacobi_schur", "the rbjacobi stencil has not been allocated")) return;
    const long half = lat->get_size_cv() / 2;
    apply_M_rbjacobi_eo(b_r, b);
    cxpay(b, -1.0, b_r, half);
    zero_vector(b_r + half, half);
  }
  // x = B^-1 (y_e, b_o - H'_oe y_e), accumulated into x like the reference (:1956)
  void reconstruct_M_rbjacobi_schur(complex<double>* x, complex<double>* y_e, complex<double>* b)
  {
    if (!need(built_rbjacobi, "reconstruct_M_rbjacobi_schur", "the rbjacobi stencil has not been allocated")) return;
    complex<double>* tmp = scratch_eo();
    const long half = lat->get_size_cv() / 2;
    qmg_stencil_desc d = describe(0, rbjacobi_hopping, 0.0, 0.0, 0.0);
    QMG_CHK(qmg_stencil_apply(&d, QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ODD_ROWS_ONLY, 15, qmg_host::P(tmp), qmg_host::P(y_e)));
    cxpay(b + half, -1.0, tmp + half, half);
    copy_vector(tmp, y_e, half);
    apply_M_rbjacobi_cinv(x, tmp);
  }
  void reconstruct_M_rbjacobi_schur_to_rbjacobi(complex<double>* x, complex<double>* y_e, complex<double>* b)
  {
    if (!need(built_rbjacobi, "reconstruct_M_rbjacobi_schur_to_rbjacobi", "the rbjacobi stencil has not been allocated")) return;
    complex<double>* tmp = scratch_eo();
    const long half = lat->get_size_cv() / 2;
    qmg_stencil_desc d = describe(0, rbjacobi_hopping, 0.0, 0.0, 0.0);
    QMG_CHK(qmg_stencil_apply(&d, QMG_APPLY_HOP_TO_ODD | QMG_APPLY_ODD_ROWS_ONLY, 15, qmg_host::P(tmp), qmg_host::P(y_e)));
    caxpbyz(1.0, b + half, -1.0, tmp + half, x + half, half);
    copy_vector(x, y_e, half);
  }

  // ---- (M B^-1)^dagger link set (stencil_2d.h:1989-2060)
  void build_rbj_dagger_stencil()
  {
    if (built_rbj_dagger) { std::cout << "[QMG-WARNING]: Tried to call build_rbj_dagger_stencil, but it's already been called once.\n"; return; }
    if (!need(built_rbjacobi, "build_rbj_dagger_stencil", "the right jacobi stencil has not been built yet")) return;
    const long cm = lat->get_size_cm();
    if (rbjacobi_clover != 0) rbj_dagger_clover = allocate_vector<complex<double> >(cm);
    if (rbjacobi_hopping != 0) rbj_dagger_hopping = allocate_vector<complex<double> >(lat->get_size_hopping());
    build_conjugate_set(rbjacobi_clover, rbjacobi_hopping, rbj_dagger_clover, rbj_dagger_hopping);
    if (rbjacobi_cinv != 0)
    {
      rbj_dagger_cinv = allocate_vector<complex<double> >(cm);
      cMATcopy_conjtrans_square(rbjacobi_cinv, rbj_dagger_cinv, lat->get_volume(), lat->get_nc());
    }
    unsupported_pieces(twolink, corner);
    built_rbj_dagger = true;
  }
  bool perform_swap_rbj_dagger()
  {
    if (!need(built_rbj_dagger, "perform_swap_rbj_dagger", "the right jacobi dagger stencil has not been allocated")) return false;
    std::swap(clover, rbj_dagger_clover); std::swap(hopping, rbj_dagger_hopping); std::swap(twolink, rbj_dagger_twolink); std::swap(corner, rbj_dagger_corner);
    if (!swap_rbj_dagger) { shift = 0.0; eo_shift = 0.0; dof_shift = 0.0; }
    else { shift = shift_backup; eo_shift = eo_shift_backup; dof_shift = dof_shift_backup; }
    swap_rbj_dagger = !swap_rbj_dagger;
    return swap_rbj_dagger;
  }
  void print_stencil_rbj_dagger_site(int x, int y, string prefix = "")
  {
    if (!need(built_rbj_dagger, "print_stencil_rbj_dagger_site", "the right jacobi dagger stencil has not been allocated")) return;
    perform_swap_rbj_dagger(); print_stencil_site(x, y, prefix); perform_swap_rbj_dagger();
  }
  QMG_FAMILY_PIECE(apply_M_rbj_dagger_clover, built_rbj_dagger, perform_swap_rbj_dagger, "the right jacobi dagger", rbj_dagger_clover, apply_M_clover)
  QMG_FAMILY_PIECE(apply_M_rbj_dagger_eo, built_rbj_dagger, perform_swap_rbj_dagger, "the right jacobi dagger", rbj_dagger_hopping, apply_M_eo)
  QMG_FAMILY_PIECE_DIR(apply_M_rbj_dagger_eo, built_rbj_dagger, perform_swap_rbj_dagger, "the right jacobi dagger", rbj_dagger_hopping, apply_M_eo)
  QMG_FAMILY_PIECE(apply_M_rbj_dagger_oe, built_rbj_dagger, perform_swap_rbj_dagger, "the right jacobi dagger", rbj_dagger_hopping, apply_M_oe)
  QMG_FAMILY_PIECE_DIR(apply_M_rbj_dagger_oe, built_rbj_dagger, perform_swap_rbj_dagger, "the right jacobi dagger", rbj_dagger_hopping, apply_M_oe)
  QMG_FAMILY_PIECE(apply_M_rbj_dagger_hopping, built_rbj_dagger, perform_swap_rbj_dagger, "the right jacobi dagger", rbj_dagger_hopping, apply_M_hopping)
  void apply_M_rbj_dagger_hopping(complex<double>* lhs, complex<double>* rhs, stencil_dir_index) { apply_M_rbj_dagger_hopping(lhs, rhs); }
  void apply_M_rbj_dagger_shift(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_rbj_dagger, "apply_M_rbj_dagger_shift", "the right jacobi dagger stencil has not been allocated")) return;
    perform_swap_rbj_dagger(); apply_M_shift(lhs, rhs); perform_swap_rbj_dagger();
  }
  void apply_M_rbj_dagger(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need(built_rbj_dagger, "apply_M_rbj_dagger", "the right jacobi dagger stencil has not been allocated")) return;
    perform_swap_rbj_dagger(); apply_M(lhs, rhs); perform_swap_rbj_dagger();
  }
#undef QMG_FAMILY_PIECE
#undef QMG_FAMILY_PIECE_DIR

  // ---- normal equations of the rbjacobi system (stencil_2d.h:2282-2411)
  bool need_both(const char* fn)
  {
    return need(built_rbjacobi, fn, "the rbjacobi stencil has not been built") && need(built_rbj_dagger, fn, "the rbjacobi dagger stencil has not been built");
  }
  void apply_M_rbjacobi_MDM(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need_both("apply_M_rbjacobi_MDM")) return;
    complex<double>* tmp = scratch_extra();
    apply_M_rbjacobi_fused(tmp, rhs, false);
    apply_M_rbj_dagger(lhs, tmp);
  }
  void prepare_M_rbjacobi_MDM(complex<double>* Mdagger_b, complex<double>* b) { if (need_both("prepare_M_rbjacobi_MDM")) apply_M_rbj_dagger(Mdagger_b, b); }
  void reconstruct_M_rbjacobi_MDM(complex<double>* x, complex<double>* y) { if (need_both("reconstruct_M_rbjacobi_MDM")) apply_M_rbjacobi_cinv(x, y); }
  // the reference copies x into y here (stencil_2d.h:2351); kept
  void reconstruct_M_rbjacobi_MDM_to_rbjacobi(complex<double>* x, complex<double>* y) { if (need_both("reconstruct_M_rbjacobi_MDM_to_rbjacobi")) copy_vector(y, x, lat->get_size_cv()); }
  void apply_M_rbjacobi_MMD(complex<double>* lhs, complex<double>* rhs)
  {
    if (!need_both("apply_M_rbjacobi_MMD")) return;
    complex<double>* tmp = scratch_extra();
    perform_swap_rbj_dagger(); apply_M_overwrite(tmp, rhs); perform_swap_rbj_dagger();
    apply_M_rbjacobi(lhs, tmp);
  }
  void reconstruct_M_rbjacobi_MMD(complex<double>* x, complex<double>* y)
  {
    if (!need_both("reconstruct_M_rbjacobi_MMD")) return;
    complex<double>* tmp = scratch_extra();
    perform_swap_rbj_dagger(); apply_M_overwrite(x, y); perform_swap_rbj_dagger();
    cMATxy(rbjacobi_cinv, x, tmp, lat->get_volume(), lat->get_nc(), lat->get_nc());
    copy_vector(x, tmp, lat->get_size_cv());
  }
  void reconstruct_M_rbjacobi_MMD_to_rbjacobi(complex<double>* x, complex<double>* y)
  {
    if (!need_both("reconstruct_M_rbjacobi_MMD_to_rbjacobi")) return;
    perform_swap_rbj_dagger(); apply_M_overwrite(x, y); perform_swap_rbj_dagger();
  }

  // ---- dispatch on QMGStencilType (stencil_2d.h:2418-2566); all accumulate like the member functions they name
  void apply_M(complex<double>* lhs, complex<double>* rhs, QMGStencilType stencil)
  {
    switch (stencil)
    {
      case QMG_MATVEC_ORIGINAL: apply_M(lhs, rhs); break;
      case QMG_MATVEC_DAGGER: apply_M_dagger(lhs, rhs); break;
      case QMG_MATVEC_RIGHT_JACOBI: apply_M_rbjacobi(lhs, rhs); break;
      case QMG_MATVEC_RIGHT_SCHUR: apply_M_rbjacobi_schur(lhs, rhs); break;
      case QMG_MATVEC_M_MDAGGER: apply_M_M_dagger(lhs, rhs); break;
      case QMG_MATVEC_MDAGGER_M: apply_M_dagger_M(lhs, rhs); break;
      case QMG_MATVEC_RBJ_DAGGER: apply_M_rbj_dagger(lhs, rhs); break;
      case QMG_MATVEC_RBJ_M_MDAGGER: apply_M_rbjacobi_MMD(lhs, rhs); break;
      case QMG_MATVEC_RBJ_MDAGGER_M: apply_M_rbjacobi_MDM(lhs, rhs); break;
      default: break;
    }
  }
  void prepare_M(complex<double>* b_prep, complex<double>* b, QMGStencilType stencil)
  {
    switch (stencil)
    {
      case QMG_MATVEC_RIGHT_SCHUR: prepare_M_rbjacobi_schur(b_prep, b); break;
      case QMG_MATVEC_MDAGGER_M: prepare_M_dagger_M(b_prep, b); break;
      case QMG_MATVEC_RBJ_MDAGGER_M: prepare_M_rbjacobi_MDM(b_prep, b); break;
      case QMG_MATVEC_ORIGINAL: case QMG_MATVEC_DAGGER: case QMG_MATVEC_RIGHT_JACOBI: case QMG_MATVEC_M_MDAGGER:
      case QMG_MATVEC_RBJ_DAGGER: case QMG_MATVEC_RBJ_M_MDAGGER:
        copy_vector(b_prep, b, lat->get_size_cv()); break;
      default: break;
    }
  }
  void reconstruct_M(complex<double>* x, complex<double>* y, complex<double>* b, QMGStencilType stencil)
  {
    switch (stencil)
    {
      case QMG_MATVEC_RIGHT_JACOBI: reconstruct_M_rbjacobi(x, y); break;
      case QMG_MATVEC_RIGHT_SCHUR: reconstruct_M_rbjacobi_schur(x, y, b); break;
      case QMG_MATVEC_M_MDAGGER: reconstruct_M_M_dagger(x, y); break;
      case QMG_MATVEC_RBJ_M_MDAGGER: reconstruct_M_rbjacobi_MMD(x, y); break;
      case QMG_MATVEC_RBJ_MDAGGER_M: reconstruct_M_rbjacobi_MDM(x, y); break;
      case QMG_MATVEC_ORIGINAL: case QMG_MATVEC_DAGGER: case QMG_MATVEC_MDAGGER_M: case QMG_MATVEC_RBJ_DAGGER:
        copy_vector(x, y, lat->get_size_cv()); break;
      default: break;
    }
  }
  static matrix_op_cplx get_apply_function(QMGStencilType stencil)
  {
    switch (stencil)
    {
      case QMG_MATVEC_ORIGINAL: return apply_stencil_2D_M;
      case QMG_MATVEC_DAGGER: return apply_stencil_2D_M_dagger;
      case QMG_MATVEC_RIGHT_JACOBI: return apply_stencil_2D_M_rbjacobi;
      case QMG_MATVEC_RIGHT_SCHUR: return apply_stencil_2D_M_rbjacobi_schur;
      case QMG_MATVEC_M_MDAGGER: return apply_stencil_2D_M_M_dagger;
      case QMG_MATVEC_MDAGGER_M: return apply_stencil_2D_M_dagger_M;
      case QMG_MATVEC_RBJ_DAGGER: return apply_stencil_2D_M_rbj_dagger;
      case QMG_MATVEC_RBJ_M_MDAGGER: return apply_stencil_2D_M_rbjacobi_MMD;
      case QMG_MATVEC_RBJ_MDAGGER_M: return apply_stencil_2D_M_rbjacobi_MDM;
      default: return 0;
    }
  }

  // overwrite flavours of the family applies used by the wrappers (one launch each, nothing read from lhs)
  void overwrite_dagger(complex<double>* lhs, complex<double>* rhs) { perform_swap_dagger(); apply_M_overwrite(lhs, rhs); perform_swap_dagger(); }
  void overwrite_rbj_dagger(complex<double>* lhs, complex<double>* rhs) { perform_swap_rbj_dagger(); apply_M_overwrite(lhs, rhs); perform_swap_rbj_dagger(); }

protected:
  // out_clover = in_clover^dag ; out_hopping_mu(x) = [in_hopping_{-mu}(x + mu)]^dag
  void build_conjugate_set(complex<double>* in_cl, complex<double>* in_hp, complex<double>* out_cl, complex<double>* out_hp)
  {
    if (lat->get_volume() == 1)
    {
      if (in_cl != 0) cMATcopy_conjtrans_square(in_cl, out_cl, 1, lat->get_nc());
      if (in_hp != 0) zero_vector(out_hp, lat->get_size_hopping());
      return;
    }
    QMG_CHK(qmg_build_dagger(lat->get_dim_mu(0), lat->get_dim_mu(1), lat->get_nc(), qmg_host::P(in_cl), qmg_host::P(in_hp), qmg_host::P(out_cl), qmg_host::P(out_hp)));
  }
};

// ---------------------------------------------------------------- wrappers --
// lhs = M rhs.  The reference zeroes lhs and accumulates (stencil_2d.h:2571-2576); here the kernel writes lhs directly.
inline void apply_stencil_2D_M(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{ ((Stencil2D*)extra_data)->apply_M_overwrite(lhs, rhs); }
inline void apply_stencil_2D_M_piece_clover(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (st->clover == 0) { zero_vector(lhs, st->lat->get_size_cv()); return; }
  st->launch(QMG_APPLY_CLOVER, 15, lhs, rhs);
}
inline void apply_stencil_2D_M_piece_hopping(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (st->hopping == 0) { zero_vector(lhs, st->lat->get_size_cv()); return; }
  st->launch(QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD, 15, lhs, rhs);
}
inline void apply_stencil_2D_M_dagger(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->built_dagger)
  {
    zero_vector(lhs, st->lat->get_size_cv());   // stencil_2d.h:2595-2602: zeroed, then the warning
    std::cout << "[QMG-WARNING]: Tried to call apply_stencil_2D_M_dagger, but the dagger stencil has not been built.\n";
    return;
  }
  st->overwrite_dagger(lhs, rhs);
}
inline void apply_stencil_2D_M_dagger_M(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->built_dagger) { std::cout << "[QMG-WARNING]: Tried to call apply_stencil_2D_M_dagger_M, but the dagger stencil has not been built.\n"; return; }
  complex<double>* tmp = st->expose_internal_cvector();
  st->apply_M_overwrite(tmp, rhs);
  st->overwrite_dagger(lhs, tmp);
}
inline void apply_stencil_2D_M_M_dagger(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->built_dagger) { std::cout << "[QMG-WARNING]: Tried to call apply_stencil_2D_M_M_dagger, but the dagger stencil has not been built.\n"; return; }
  complex<double>* tmp = st->expose_internal_cvector();
  st->overwrite_dagger(tmp, rhs);
  st->apply_M_overwrite(lhs, tmp);
}
inline void apply_stencil_2D_M_rbjacobi(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->built_rbjacobi) { std::cout << "[QMG-WARNING]: Tried to call apply_stencil_2D_M_rbjacobi, but the rbjacobi stencil has not been built.\n"; return; }
  st->apply_M_rbjacobi_fused(lhs, rhs, false);
}
inline void apply_stencil_2D_M_rbjacobi_cinv(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->built_rbjacobi) { std::cout << "[QMG-WARNING]: Tried to call apply_stencil_2D_M_rbjacobi_cinv, but the rbjacobi stencil has not been built.\n"; return; }
  cMATxy(st->rbjacobi_cinv, rhs, lhs, st->lat->get_volume(), st->lat->get_nc(), st->lat->get_nc());
}
inline void apply_stencil_2D_M_rbjacobi_schur(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->built_rbjacobi) { std::cout << "[QMG-WARNING]: Tried to call apply_stencil_2D_M_rbjacobi_schur, but the rbjacobi stencil has not been built.\n"; return; }
  st->apply_M_rbjacobi_schur(lhs, rhs);   // writes the even half of lhs outright
}
inline void apply_stencil_2D_M_rbj_dagger(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->built_rbj_dagger) { std::cout << "[QMG-WARNING]: Tried to call apply_stencil_2D_M_rbj_dagger, but the rbjacobi dagger stencil has not been built.\n"; return; }
  st->overwrite_rbj_dagger(lhs, rhs);
}
inline void apply_stencil_2D_M_rbjacobi_MMD(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->need_both("apply_stencil_2D_M_rbjacobi_MMD")) return;
  complex<double>* tmp = st->expose_internal_cvector();
  st->overwrite_rbj_dagger(tmp, rhs);
  st->apply_M_rbjacobi_fused(lhs, tmp, false);
}
inline void apply_stencil_2D_M_rbjacobi_MDM(complex<double>* lhs, complex<double>* rhs, void* extra_data)
{
  Stencil2D* st = (Stencil2D*)extra_data;
  if (!st->need_both("apply_stencil_2D_M_rbjacobi_MDM")) return;
  complex<double>* tmp = st->expose_internal_cvector();
  st->apply_M_rbjacobi_fused(tmp, rhs, false);
  st->overwrite_rbj_dagger(lhs, tmp);
}

#endif
