// Flat C driver API written ONLY against the reference's public class API
// (Lattice2D, Stencil2D and its operators, TransferMG, CoarseOperator2D,
// StatefulMultigridMG, the apply_stencil_2D_* wrappers and the quantum-linalg
// solver entry points).  The same text is compiled twice:
//
//   * quantum-mg_b200/host/qmg_host_capi.cpp  -> libqmg_host.so
//       against include/qmg/ (the B200 host classes; vectors live in HBM)
//   * oracle/ref_capi.cpp                     -> oracle/_ref/libqmg_ref.so
//       against the UNMODIFIED headers under /root/reference plus the
//       clean-room quantum-linalg shim (plain host memory; test oracle only)
//
// so a parity test is literally "same driver, two back ends".  The including
// file defines CAPI(name) (symbol prefix) and the staging helpers
//   capi_alloc(n) / capi_free(p)         storage usable by the class API
//   capi_put(dst, host_src, n)           host array  -> class-API storage
//   capi_get(host_dst, src, n)           class-API storage -> host array
// which are memcpy for the reference build and H2D/D2H copies for the GPU build.
//
// All array arguments of this API are HOST pointers to interleaved complex<double>.

#include <chrono>
#include <random>
#include <vector>

typedef std::complex<double> capi_cd;

namespace capi {

// RAII staging of one host array into class-API storage.
struct Stage
{
  capi_cd* p; capi_cd* host; long n; bool write_back;
  Stage(const capi_cd* h, long n_, bool in, bool out) : p(0), host((capi_cd*)h), n(n_), write_back(out)
  {
    p = capi_alloc(n);
    if (in && h != 0) capi_put(p, h, n);
  }
  ~Stage() { if (write_back && host != 0) capi_get(host, p, n); capi_free(p); }
  operator capi_cd*() { return p; }
private:
  Stage(const Stage&); Stage& operator=(const Stage&);
};

struct LatticeH { Lattice2D* lat; };
struct StencilH { Stencil2D* op; bool owned; };
struct TransferH { TransferMG* t; Lattice2D* fine; Lattice2D* coarse; };
struct MgH
{
  StatefulMultigridMG* mg;
  StatefulMultigridMG::CoarsestSolveMG* coarsest;
  std::vector<StatefulMultigridMG::LevelSolveMG*> levels;
};

inline capi_cd* stencil_array(Stencil2D* s, int which, long& n)
{
  const long cm = s->lat->get_size_cm();
  switch (which)
  {
    case 0: n = cm; return s->clover;
    case 1: n = 4 * cm; return s->hopping;
    case 2: n = cm; return s->dagger_clover;
    case 3: n = 4 * cm; return s->dagger_hopping;
    case 4: n = cm; return s->rbjacobi_clover;
    case 5: n = 4 * cm; return s->rbjacobi_hopping;
    case 6: n = cm; return s->rbjacobi_cinv;
    case 7: n = cm; return s->rbj_dagger_clover;
    case 8: n = 4 * cm; return s->rbj_dagger_hopping;
    case 9: n = cm; return s->rbj_dagger_cinv;
  }
  n = 0; return 0;
}

} // namespace capi

extern "C" {

// ------------------------------------------------------------------ lattice --
void* CAPI(lattice_new)(int X, int Y, int nc) { capi::LatticeH* h = new capi::LatticeH; h->lat = new Lattice2D(X, Y, nc); return h; }
void CAPI(lattice_free)(void* h_) { capi::LatticeH* h = (capi::LatticeH*)h_; delete h->lat; delete h; }
int CAPI(lattice_coord_to_index)(void* h, int x, int y) { return ((capi::LatticeH*)h)->lat->coord_to_index(x, y); }
void CAPI(lattice_index_to_coord)(void* h, int i, int* xy) { ((capi::LatticeH*)h)->lat->index_to_coord(i, xy[0], xy[1]); }
// sizes: volume, size_cv, size_cm, size_gauge, size_hopping, size_corner
void CAPI(lattice_sizes)(void* h_, int* out)
{
  Lattice2D* l = ((capi::LatticeH*)h_)->lat;
  out[0] = l->get_volume(); out[1] = l->get_size_cv(); out[2] = l->get_size_cm();
  out[3] = l->get_size_gauge(); out[4] = l->get_size_hopping(); out[5] = l->get_size_corner();
}
int CAPI(lattice_cm_index)(void* h, int x, int y, int c1, int c2) { return ((capi::LatticeH*)h)->lat->cm_coord_to_index(x, y, c1, c2); }
int CAPI(lattice_hopping_index)(void* h, int x, int y, int c1, int c2, int mu) { return ((capi::LatticeH*)h)->lat->hopping_coord_to_index(x, y, c1, c2, mu); }
int CAPI(lattice_gauge_index)(void* h, int x, int y, int c1, int c2, int mu) { return ((capi::LatticeH*)h)->lat->gauge_coord_to_index(x, y, c1, c2, mu); }
void CAPI(lattice_cv_index_to_coord)(void* h, int i, int* xyc) { ((capi::LatticeH*)h)->lat->cv_index_to_coord(i, xyc[0], xyc[1], xyc[2]); }

// cshift<complex<double>> (cshift/cshift_2d.h:225); lhs is in/out (only half is written per source parity)
void CAPI(cshift)(capi_cd* lhs, const capi_cd* rhs, int cdir, int eo, int dof, void* lat_)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  const long n = (long)lat->get_volume() * dof;
  capi::Stage dl(lhs, n, true, true), dr(rhs, n, true, false);
  cshift((capi_cd*)dl, (capi_cd*)dr, (qmg_cshift_dir)cdir, (qmg_eo)eo, dof, lat);
}

// ------------------------------------------------------------------ stencils --
// gauge: host array of size_gauge complex links on the nc=1 lattice (u1/u1_utils.h layout)
void* CAPI(wilson_new)(void* lat_, double mass_re, double mass_im, const capi_cd* gauge, double wilson_coeff)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage g(gauge, 2L * lat->get_volume(), true, false);
  capi::StencilH* h = new capi::StencilH;
  h->op = new Wilson2D(lat, capi_cd(mass_re, mass_im), (capi_cd*)g, wilson_coeff); h->owned = true;
  return h;
}
void* CAPI(staggered_new)(void* lat_, double mass_re, double mass_im, const capi_cd* gauge)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage g(gauge, 2L * lat->get_volume(), true, false);
  capi::StencilH* h = new capi::StencilH;
  h->op = new Staggered2D(lat, capi_cd(mass_re, mass_im), (capi_cd*)g); h->owned = true;
  return h;
}
void* CAPI(laplace_new)(void* lat_, double msq_re, double msq_im, const capi_cd* gauge)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage g(gauge, 2L * lat->get_volume(), true, false);
  capi::StencilH* h = new capi::StencilH;
  h->op = new GaugedLaplace2D(lat, capi_cd(msq_re, msq_im), (capi_cd*)g); h->owned = true;
  return h;
}
void* CAPI(dwf_new)(void* lat_, double mass_re, double mass_im, const capi_cd* gauge, int Ls, double M5)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage g(gauge, 2L * lat->get_volume(), true, false);
  Stencil2D* op = createDwfLs(lat, capi_cd(mass_re, mass_im), (capi_cd*)g, Ls, M5);
  if (op == 0) return 0;
  capi::StencilH* h = new capi::StencilH; h->op = op; h->owned = true;
  return h;
}
// A bare stencil with caller-supplied blocks (CoarseOperator2D's first ctor, operators/coarse.h:76).
void* CAPI(generic_new)(void* lat_, int is_chiral, int def_chirality, const double* shifts6, const capi_cd* clover, const capi_cd* hopping)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  int pieces = 0;
  if (clover != 0) pieces |= QMG_PIECE_CLOVER;
  if (hopping != 0) pieces |= QMG_PIECE_HOPPING;
  CoarseOperator2D* op = new CoarseOperator2D(lat, pieces, is_chiral != 0, (QMGDefaultChirality)def_chirality,
                                              capi_cd(shifts6[0], shifts6[1]), capi_cd(shifts6[2], shifts6[3]), capi_cd(shifts6[4], shifts6[5]));
  if (clover != 0) capi_put(op->clover, clover, lat->get_size_cm());
  if (hopping != 0) capi_put(op->hopping, hopping, lat->get_size_hopping());
  op->generated = true;
  capi::StencilH* h = new capi::StencilH; h->op = op; h->owned = true;
  return h;
}
void CAPI(stencil_free)(void* h_) { capi::StencilH* h = (capi::StencilH*)h_; if (h->owned) delete h->op; delete h; }

// which: 0 clover 1 hopping 2 dagger_clover 3 dagger_hopping 4 rbjacobi_clover 5 rbjacobi_hopping
//        6 rbjacobi_cinv 7 rbj_dagger_clover 8 rbj_dagger_hopping 9 rbj_dagger_cinv ; returns element count (0 if absent)
long CAPI(stencil_get)(void* h_, int which, capi_cd* out)
{
  long n; capi_cd* p = capi::stencil_array(((capi::StencilH*)h_)->op, which, n);
  if (p == 0) return 0;
  if (out != 0) capi_get(out, p, n);
  return n;
}
// the n18 mutation: cxpy(noise, stencil->clover, cm_size)
// (tests/n18_rbjacobi_stencil_test/rbjacobi_stencil_test.cpp:137)
void CAPI(stencil_add_to)(void* h_, int which, const capi_cd* noise)
{
  long n; capi_cd* p = capi::stencil_array(((capi::StencilH*)h_)->op, which, n);
  if (p == 0) return;
  capi::Stage dn(noise, n, true, false);
  cxpy((capi_cd*)dn, p, (int)n);
}
// shifts: out[0..5] = shift, eo_shift, dof_shift
void CAPI(stencil_get_shifts)(void* h_, double* out)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  out[0] = real(s->get_shift()); out[1] = imag(s->get_shift());
  out[2] = real(s->get_shift_eo()); out[3] = imag(s->get_shift_eo());
  out[4] = real(s->get_shift_dof()); out[5] = imag(s->get_shift_dof());
}
void CAPI(stencil_update_shifts)(void* h_, const double* s6)
{
  ((capi::StencilH*)h_)->op->update_shifts(capi_cd(s6[0], s6[1]), capi_cd(s6[2], s6[3]), capi_cd(s6[4], s6[5]));
}
// which: 1 dagger, 2 rbjacobi, 4 rbj_dagger (bitmask, built in that order)
void CAPI(stencil_build)(void* h_, int which)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  if (which & 1) s->build_dagger_stencil();
  if (which & 2) s->build_rbjacobi_stencil();
  if (which & 4) s->build_rbj_dagger_stencil();
}
int CAPI(stencil_built)(void* h_)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  return (s->built_dagger ? 1 : 0) | (s->built_rbjacobi ? 2 : 0) | (s->built_rbj_dagger ? 4 : 0);
}

// B200 extension: switch the link-compressed (gamma5-hermitian) apply on / off for this operator; returns 1 when active.
// The reference build has no such mode: it answers 0 and changes nothing.
int CAPI(stencil_gamma5_hermitian)(void* h_, int on)
{
#ifdef QMG_B200_HOST
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  if (!on) { s->disable_gamma5_hermitian_apply(); return 0; }
  return s->enable_gamma5_hermitian_apply() ? 1 : 0;
#else
  (void)h_; (void)on; return 0;
#endif
}

// B200 extension: matrix-free apply of a Wilson2D operator from the gauge field `gauge` (host, 2 V links, the array the
// operator was built from): on = 1 switch on (after the exact check of the stored blocks; returns 1 when active), 0 switch
// off, 2 / 3 pause / resume without dropping the gauge copy.  Reference build and other operators: 0, nothing changes.
int CAPI(stencil_matrix_free)(void* h_, const capi_cd* gauge, int on)
{
#ifdef QMG_B200_HOST
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  Wilson2D* w = dynamic_cast<Wilson2D*>(s);
  if (w == 0) return 0;
  if (on == 0) { w->disable_matrix_free_apply(); return 0; }
  if (on == 2) { w->pause_matrix_free_apply(true); return 0; }
  if (on == 3) { w->pause_matrix_free_apply(false); return w->uses_matrix_free_apply() ? 1 : 0; }
  if (gauge == 0) return 0;
  capi::Stage g(gauge, 2L * s->lat->get_volume(), true, false);
  return w->enable_matrix_free_apply((capi_cd*)g) ? 1 : 0;
#else
  (void)h_; (void)gauge; (void)on; return 0;
#endif
}

// lhs = M_type rhs through the function-pointer wrappers apply_stencil_2D_* (stencil_2d.h:2571-2716)
void CAPI(stencil_apply)(void* h_, int type, capi_cd* lhs, const capi_cd* rhs)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const long n = s->lat->get_size_cv();
  capi::Stage dl(lhs, n, true, true), dr(rhs, n, true, false);
  matrix_op_cplx fn = Stencil2D::get_apply_function((QMGStencilType)type);
  if (fn != 0) fn((capi_cd*)dl, (capi_cd*)dr, (void*)s);
}

// The accumulate-into-lhs member functions (stencil_2d.h:666-936, 1848).
// piece: 0 apply_M  1 clover  2 eo  3 oe  4 hopping  5 hopping(dir)  6 shift  7 eo(dir)  8 oe(dir)
//        9 ee  10 oo  11 rbjacobi_cinv  12 apply_M(type=dir)
void CAPI(stencil_apply_piece)(void* h_, int piece, int dir, capi_cd* lhs, const capi_cd* rhs)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const long n = s->lat->get_size_cv();
  capi::Stage dl(lhs, n, true, true), dr(rhs, n, true, false);
  capi_cd* l = dl; capi_cd* r = dr;
  switch (piece)
  {
    case 0: s->apply_M(l, r); break;
    case 1: s->apply_M_clover(l, r); break;
    case 2: s->apply_M_eo(l, r); break;
    case 3: s->apply_M_oe(l, r); break;
    case 4: s->apply_M_hopping(l, r); break;
    case 5: s->apply_M_hopping(l, r, (stencil_dir_index)dir); break;
    case 6: s->apply_M_shift(l, r); break;
    case 7: s->apply_M_eo(l, r, (stencil_dir_index)dir); break;
    case 8: s->apply_M_oe(l, r, (stencil_dir_index)dir); break;
    case 9: s->apply_M_ee(l, r); break;
    case 10: s->apply_M_oo(l, r); break;
    case 11: s->apply_M_rbjacobi_cinv(l, r); break;
    case 12: s->apply_M(l, r, (QMGStencilType)dir); break;
  }
}
void CAPI(stencil_prepare)(void* h_, int type, capi_cd* b_prep, const capi_cd* b)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const long n = s->lat->get_size_cv();
  capi::Stage dp(b_prep, n, true, true), db(b, n, true, false);
  s->prepare_M((capi_cd*)dp, (capi_cd*)db, (QMGStencilType)type);
}
void CAPI(stencil_reconstruct)(void* h_, int type, capi_cd* x, const capi_cd* y, const capi_cd* b)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const long n = s->lat->get_size_cv();
  capi::Stage dx(x, n, true, true), dy(y, n, true, false), db(b, n, true, false);
  s->reconstruct_M((capi_cd*)dx, (capi_cd*)dy, (capi_cd*)db, (QMGStencilType)type);
}
// op: 0 gamma5(a) in place  1 gamma5(b <- a)  2 sigma1(a) in place  3 sigma1(b <- a)
//     4 chiral_projection(a, up)  5 chiral_projection(a, down)  6 projection_copy(a -> b, up)  7 (down)
//     8 chiral_projection_both(a -> up in place, b <- down)   9.. apply_sigma(b <- a, type = op - 9)
void CAPI(stencil_chiral)(void* h_, int op, capi_cd* a, capi_cd* b)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const long n = s->lat->get_size_cv();
  capi::Stage da(a, n, true, true), db(b, n, b != 0, b != 0);
  switch (op)
  {
    case 0: s->gamma5((capi_cd*)da); break;
    case 1: s->gamma5((capi_cd*)db, (capi_cd*)da); break;
    case 2: s->sigma1((capi_cd*)da); break;
    case 3: s->sigma1((capi_cd*)db, (capi_cd*)da); break;
    case 4: s->chiral_projection((capi_cd*)da, true); break;
    case 5: s->chiral_projection((capi_cd*)da, false); break;
    case 6: s->chiral_projection_copy((capi_cd*)da, (capi_cd*)db, true); break;
    case 7: s->chiral_projection_copy((capi_cd*)da, (capi_cd*)db, false); break;
    case 8: s->chiral_projection_both((capi_cd*)da, (capi_cd*)db); break;
    default: s->apply_sigma((capi_cd*)db, (capi_cd*)da, (QMGSigmaType)(op - 9)); break;
  }
}

// Wall-clock seconds for `reps` wrapper applies of type `type` (after `warm` untimed ones).
double CAPI(stencil_time_apply)(void* h_, int type, int warm, int reps, const capi_cd* rhs)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const long n = s->lat->get_size_cv();
  capi::Stage dl(0, n, false, false), dr(rhs, n, true, false);
  matrix_op_cplx fn = Stencil2D::get_apply_function((QMGStencilType)type);
  for (int i = 0; i < warm; i++) fn((capi_cd*)dl, (capi_cd*)dr, (void*)s);
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; i++) fn((capi_cd*)dl, (capi_cd*)dr, (void*)s);
  capi_barrier();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ------------------------------------------------------------------- solvers --
// solver: 0 CG  1 CG(restart=iparam)  2 GCR  3 GCR(restart)  4 MR(omega=dparam)  5 BiCGstab-L(L=iparam)
//         6 Richardson(omega=dparam, check_freq=iparam)  7 BiCGstab  8 TFQMR (the oracle's shim only declares it)
// n = number of complex unknowns handed to the solver (size_cv, or size_cv/2 for Schur systems)
// info: resSq, iter, success, ops_count
void CAPI(solve)(void* h_, int solver, int type, capi_cd* x, const capi_cd* b, int n, int max_iter, double tol, int iparam, double dparam, int verbosity, double* info)
{
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const long full = s->lat->get_size_cv();
  capi::Stage dx(x, full, true, true), db(b, full, true, false);
  matrix_op_cplx fn = Stencil2D::get_apply_function((QMGStencilType)type);
  inversion_verbose_struct verb((inversion_verbose_level)verbosity, "[CAPI-SOLVE]: ");
  inversion_info inv;
  switch (solver)
  {
    case 0: inv = minv_vector_cg((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, fn, (void*)s, &verb); break;
    case 1: inv = minv_vector_cg_restart((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, iparam, fn, (void*)s, &verb); break;
    case 2: inv = minv_vector_gcr((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, fn, (void*)s, &verb); break;
    case 3: inv = minv_vector_gcr_restart((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, iparam, fn, (void*)s, &verb); break;
    case 4: inv = minv_vector_minres((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, dparam, fn, (void*)s, &verb); break;
    case 5: inv = minv_vector_bicgstab_l((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, iparam, fn, (void*)s, &verb); break;
    case 6: inv = minv_vector_richardson((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, dparam, iparam, fn, (void*)s, &verb); break;
    case 7: inv = minv_vector_bicgstab((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, fn, (void*)s, &verb); break;
    case 8: inv = minv_vector_tfqmr((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, fn, (void*)s, &verb); break;
  }
  info[0] = inv.resSq; info[1] = inv.iter; info[2] = inv.success ? 1.0 : 0.0; info[3] = inv.ops_count;
}

// ------------------------------------------------------------------ transfer --
// nullvecs: host, contiguous [nvec = coarse nc][fine size_cv]
void* CAPI(transfer_new)(void* fine_, void* coarse_, const capi_cd* nullvecs, int do_block_ortho, int save_decomp, int doubling)
{
  Lattice2D* fine = ((capi::LatticeH*)fine_)->lat; Lattice2D* coarse = ((capi::LatticeH*)coarse_)->lat;
  const int nvec = coarse->get_nc(); const long n = fine->get_size_cv();
  std::vector<capi_cd*> v(nvec);
  for (int i = 0; i < nvec; i++) { v[i] = capi_alloc(n); capi_put(v[i], nullvecs + (long)i * n, n); }
  capi::TransferH* h = new capi::TransferH;
  h->t = new TransferMG(fine, coarse, &v[0], do_block_ortho != 0, save_decomp != 0, (QMGDoublingType)doubling);
  h->fine = fine; h->coarse = coarse;
  for (int i = 0; i < nvec; i++) capi_free(v[i]);
  return h;
}
void* CAPI(transfer_new_asym)(void* fine_, void* coarse_, const capi_cd* prolong_vecs, const capi_cd* restrict_vecs, int do_block_bi_ortho, int save_decomp, int doubling)
{
  Lattice2D* fine = ((capi::LatticeH*)fine_)->lat; Lattice2D* coarse = ((capi::LatticeH*)coarse_)->lat;
  const int nvec = coarse->get_nc(); const long n = fine->get_size_cv();
  std::vector<capi_cd*> p(nvec), r(nvec);
  for (int i = 0; i < nvec; i++)
  {
    p[i] = capi_alloc(n); capi_put(p[i], prolong_vecs + (long)i * n, n);
    r[i] = capi_alloc(n); capi_put(r[i], restrict_vecs + (long)i * n, n);
  }
  capi::TransferH* h = new capi::TransferH;
  h->t = new TransferMG(fine, coarse, &p[0], &r[0], do_block_bi_ortho != 0, save_decomp != 0, (QMGDoublingType)doubling);
  h->fine = fine; h->coarse = coarse;
  for (int i = 0; i < nvec; i++) { capi_free(p[i]); capi_free(r[i]); }
  return h;
}
void CAPI(transfer_free)(void* h_) { capi::TransferH* h = (capi::TransferH*)h_; delete h->t; delete h; }
// which: 0 prolong (null_vectors), 1 restrict (restrict_null_vectors); out: [nvec][fine size_cv]
int CAPI(transfer_get_nullvecs)(void* h_, int which, capi_cd* out)
{
  capi::TransferH* h = (capi::TransferH*)h_;
  capi_cd** src = which == 0 ? h->t->null_vectors : h->t->restrict_null_vectors;
  if (src == 0) return 0;
  const int nvec = h->coarse->get_nc(); const long n = h->fine->get_size_cv();
  for (int i = 0; i < nvec; i++) capi_get(out + (long)i * n, src[i], n);
  return nvec;
}
// fine += P coarse (accumulates, transfer.h:455)
void CAPI(transfer_prolong)(void* h_, const capi_cd* coarse, capi_cd* fine)
{
  capi::TransferH* h = (capi::TransferH*)h_;
  capi::Stage dc(coarse, h->coarse->get_size_cv(), true, false), df(fine, h->fine->get_size_cv(), true, true);
  h->t->prolong_c2f((capi_cd*)dc, (capi_cd*)df);
}
// coarse += R fine (accumulates, transfer.h:487)
void CAPI(transfer_restrict)(void* h_, const capi_cd* fine, capi_cd* coarse)
{
  capi::TransferH* h = (capi::TransferH*)h_;
  capi::Stage df(fine, h->fine->get_size_cv(), true, false), dc(coarse, h->coarse->get_size_cv(), true, true);
  h->t->restrict_f2c((capi_cd*)df, (capi_cd*)dc);
}
int CAPI(transfer_props)(void* h_)
{
  TransferMG* t = ((capi::TransferH*)h_)->t;
  return (t->is_symmetric() ? 1 : 0) | (t->has_decompositions() ? 2 : 0) | (t->is_initialized() ? 4 : 0) | ((int)t->get_doubling() << 4);
}
void CAPI(transfer_get_cholesky)(void* h_, capi_cd* out)
{
  capi::TransferH* h = (capi::TransferH*)h_;
  capi::Stage d(out, h->coarse->get_size_cm(), false, true);
  h->t->copy_cholesky((capi_cd*)d);
}
void CAPI(transfer_get_LU)(void* h_, capi_cd* outL, capi_cd* outU)
{
  capi::TransferH* h = (capi::TransferH*)h_;
  capi::Stage dL(outL, h->coarse->get_size_cm(), false, true), dU(outU, h->coarse->get_size_cm(), false, true);
  h->t->copy_LU((capi_cd*)dL, (capi_cd*)dU);
}

// Galerkin coarse operator (operators/coarse.h:90): returns a stencil handle on the coarse lattice.
void* CAPI(coarse_new)(void* coarse_lat_, void* fine_stencil_, void* fine_lat_, void* transfer_, int is_chiral, int use_rbjacobi, int build_extra)
{
  Lattice2D* clat = ((capi::LatticeH*)coarse_lat_)->lat; Lattice2D* flat = ((capi::LatticeH*)fine_lat_)->lat;
  capi::StencilH* h = new capi::StencilH;
  h->op = new CoarseOperator2D(clat, ((capi::StencilH*)fine_stencil_)->op, flat, ((capi::TransferH*)transfer_)->t,
                               is_chiral != 0, use_rbjacobi != 0, (CoarseOperator2D::QMGCoarseBuildStencil)build_extra);
  h->owned = true;
  return h;
}

// CoarseOperator2D::apply_sigma(out, in, QMGSigmaTypeCoarse) (operators/coarse.h:661); type 6..9.  Returns 0 when the handle
// is not a coarse operator.
int CAPI(coarse_apply_sigma)(void* h_, int type, capi_cd* out, const capi_cd* in)
{
  CoarseOperator2D* op = dynamic_cast<CoarseOperator2D*>(((capi::StencilH*)h_)->op);
  if (op == 0) return 0;
  const long n = op->lat->get_size_cv();
  capi::Stage dout(out, n, true, true), din(in, n, true, false);
  op->apply_sigma((capi_cd*)dout, (capi_cd*)din, (QMGSigmaTypeCoarse)type);
  return 1;
}

// ----------------------------------------------------- time-slice reductions --
// op: 0 norm2sq_cv_timeslice, 1 redot_cv_timeslice, 2 dot_cv_timeslice (reductions/reductions.h); out: Y doubles (op 2: 2 Y)
void CAPI(timeslice)(void* lat_, int op, const capi_cd* a, const capi_cd* b, double* out)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  const long n = lat->get_size_cv();
  capi::Stage da(a, n, true, false), db(b, n, b != 0, false);
  if (op == 0) norm2sq_cv_timeslice(out, (capi_cd*)da, lat);
  else if (op == 1) redot_cv_timeslice(out, (capi_cd*)da, (capi_cd*)db, lat);
  else dot_cv_timeslice((capi_cd*)out, (capi_cd*)da, (capi_cd*)db, lat);
}
void CAPI(wall_source)(void* lat_, int timeslice, int color, unsigned seed, double deviation, double mean, capi_cd* out)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage d(out, lat->get_size_cv(), true, true);
  std::mt19937 gen(seed);
  gaussian_wall_source((capi_cd*)d, timeslice, color, lat, gen, deviation, mean);
}

// ------------------------------------------------------------ U(1) gauge side --
// u1/u1_utils.h through the class-API storage; gauge: host 2 V complex, phases: host 2 V doubles (nc = 1 lattice).
namespace capi {
struct StageReal
{
  double* p; double* host; long n; bool write_back;
  StageReal(const double* h, long n_, bool in, bool out) : p(0), host((double*)h), n(n_), write_back(out)
  { p = allocate_vector<double>(n); if (in && h != 0) capi_put_real(p, h, n); }
  ~StageReal() { if (write_back && host != 0) capi_get_real(host, p, n); deallocate_vector(&p); }
  operator double*() { return p; }
private:
  StageReal(const StageReal&); StageReal& operator=(const StageReal&);
};
}
// out: Re plaq, Im plaq, topological charge
void CAPI(u1_observables)(void* lat_, const capi_cd* gauge, double* out)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage g(gauge, lat->get_size_gauge(), true, false);
  const capi_cd pl = get_plaquette_u1((capi_cd*)g, lat);
  out[0] = real(pl); out[1] = imag(pl);
  out[2] = get_topo_u1((capi_cd*)g, lat);
}
double CAPI(u1_action)(void* lat_, const double* phases, double beta)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::StageReal p(phases, lat->get_size_gauge(), true, false);
  return get_noncompact_action_u1((double*)p, beta, lat);
}
void CAPI(u1_polar)(void* lat_, const double* phases, capi_cd* gauge)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::StageReal p(phases, lat->get_size_gauge(), true, false);
  capi::Stage g(gauge, lat->get_size_gauge(), false, true);
  polar_vector((double*)p, (capi_cd*)g, lat->get_size_gauge());
}
void CAPI(u1_gauge_trans)(void* lat_, capi_cd* gauge, const capi_cd* trans)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage g(gauge, lat->get_size_gauge(), true, true), t(trans, lat->get_size_cm(), true, false);
  apply_gauge_trans_u1((capi_cd*)g, (capi_cd*)t, lat);
}
void CAPI(u1_ape_smear)(void* lat_, capi_cd* smeared, const capi_cd* gauge, double alpha, int n_iter)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage s(smeared, lat->get_size_gauge(), false, true), g(gauge, lat->get_size_gauge(), true, false);
  apply_ape_smear_u1((capi_cd*)s, (capi_cd*)g, lat, alpha, n_iter);
}
void CAPI(u1_instanton)(void* lat_, capi_cd* gauge, double Q, int x0, int y0)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::Stage g(gauge, lat->get_size_gauge(), true, true);
  create_instanton_u1((capi_cd*)g, lat, Q, x0, y0);
}
void CAPI(u1_noncompact_instanton)(void* lat_, double* phases, double Q)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::StageReal p(phases, lat->get_size_gauge(), true, true);
  create_noncompact_instanton_u1((double*)p, lat, Q);
}
// n_update heatbath updates from std::mt19937(seed); phases in/out
void CAPI(u1_heatbath)(void* lat_, double* phases, double beta, int n_update, unsigned seed)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  capi::StageReal p(phases, lat->get_size_gauge(), true, true);
  std::mt19937 gen(seed);
  heatbath_noncompact_update((double*)p, lat, beta, n_update, gen);
}
// kind: 0 read_gauge_u1 -> gauge, 1 write_gauge_u1(gauge), 2 read_phase_u1 -> phases, 3 write_gauge_u1(phases)
void CAPI(u1_file)(void* lat_, int kind, const char* path, capi_cd* gauge, double* phases)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  const long n = lat->get_size_gauge();
  if (kind == 0) { capi::Stage g(gauge, n, false, true); read_gauge_u1((capi_cd*)g, lat, path); }
  else if (kind == 1) { capi::Stage g(gauge, n, true, false); write_gauge_u1((capi_cd*)g, lat, path); }
  else if (kind == 2) { capi::StageReal p(phases, n, false, true); read_phase_u1((double*)p, lat, path); }
  else if (kind == 3) { capi::StageReal p(phases, n, true, false); write_gauge_u1((double*)p, lat, path); }
}
// kind: 0 unit, 1 rand, 2 gauss(beta); 3 rand_trans (V elements).  Host draws from std::mt19937(seed) on both back ends.
void CAPI(u1_create)(void* lat_, int kind, double beta, unsigned seed, capi_cd* out)
{
  Lattice2D* lat = ((capi::LatticeH*)lat_)->lat;
  std::mt19937 gen(seed);
  const long n = kind == 3 ? lat->get_size_cm() : lat->get_size_gauge();
  capi::Stage g(out, n, false, true);
  if (kind == 0) unit_gauge_u1((capi_cd*)g, lat);
  else if (kind == 1) rand_gauge_u1((capi_cd*)g, lat, gen);
  else if (kind == 2) gauss_gauge_u1((capi_cd*)g, lat, gen, beta);
  else rand_trans_u1((capi_cd*)g, lat, gen);
}

// ----------------------------------------------------------------- multigrid --
void* CAPI(mg_new)(void* lat0_, void* stencil0_, int coarsest_type, double coarsest_tol, int coarsest_iters, int coarsest_restart)
{
  capi::MgH* h = new capi::MgH;
  h->coarsest = new StatefulMultigridMG::CoarsestSolveMG;
  h->coarsest->coarsest_stencil_app = (QMGStencilType)coarsest_type;
  h->coarsest->coarsest_tol = coarsest_tol;
  h->coarsest->coarsest_iters = coarsest_iters;
  h->coarsest->coarsest_restart_freq = coarsest_restart;
  h->mg = new StatefulMultigridMG(((capi::LatticeH*)lat0_)->lat, ((capi::StencilH*)stencil0_)->op, h->coarsest);
  return h;
}
void CAPI(mg_free)(void* h_)
{
  capi::MgH* h = (capi::MgH*)h_;
  delete h->mg;
  for (size_t i = 0; i < h->levels.size(); i++) delete h->levels[i];
  delete h->coarsest;
  delete h;
}
// iparams: fine_stencil_app, intermediate_iters, intermediate_restart, pre_iters, post_iters, pre_cgne, post_cgne
// dparams: intermediate_tol, pre_tol, post_tol
// nvecs (optional): host [coarse nc][fine size_cv] raw null vectors kept by the MG object
void CAPI(mg_push_level)(void* h_, void* new_lat_, void* transfer_, const int* iparams, const double* dparams,
                         int build_stencil, int is_chiral, int build_from, int build_extra, const capi_cd* nvecs)
{
  capi::MgH* h = (capi::MgH*)h_;
  Lattice2D* nl = ((capi::LatticeH*)new_lat_)->lat;
  capi::TransferH* th = (capi::TransferH*)transfer_;
  StatefulMultigridMG::LevelSolveMG* ls = new StatefulMultigridMG::LevelSolveMG;
  ls->fine_stencil_app = (QMGStencilType)iparams[0];
  ls->intermediate_iters = iparams[1];
  ls->intermediate_restart_freq = iparams[2];
  ls->pre_iters = iparams[3];
  ls->post_iters = iparams[4];
  ls->pre_cgne = iparams[5] != 0;
  ls->post_cgne = iparams[6] != 0;
  ls->intermediate_tol = dparams[0];
  ls->pre_tol = dparams[1];
  ls->post_tol = dparams[2];
  h->levels.push_back(ls);
  std::vector<capi_cd*> v;
  if (nvecs != 0)
  {
    const int nvec = nl->get_nc(); const long n = th->fine->get_size_cv();
    v.resize(nvec);
    for (int i = 0; i < nvec; i++) { v[i] = capi_alloc(n); capi_put(v[i], nvecs + (long)i * n, n); }
  }
  h->mg->push_level(nl, th->t, ls, build_stencil != 0, is_chiral != 0, (MultigridMG::QMGMultigridPrecondStencil)build_from,
                    (CoarseOperator2D::QMGCoarseBuildStencil)build_extra, v.empty() ? 0 : &v[0]);
  for (size_t i = 0; i < v.size(); i++) capi_free(v[i]);
}
int CAPI(mg_num_levels)(void* h_) { return ((capi::MgH*)h_)->mg->get_num_levels(); }
// borrowed stencil handle of level i (free with stencil_free; the operator itself stays owned by the MG object)
void* CAPI(mg_get_stencil)(void* h_, int level)
{
  Stencil2D* s = ((capi::MgH*)h_)->mg->get_stencil(level);
  if (s == 0) return 0;
  capi::StencilH* sh = new capi::StencilH; sh->op = s; sh->owned = false;
  return sh;
}
// one application of the K-cycle preconditioner at level 0 (stateful_multigrid.h:734)
void CAPI(mg_precond)(void* h_, capi_cd* lhs, const capi_cd* rhs, int verbosity)
{
  capi::MgH* h = (capi::MgH*)h_;
  const long n = h->mg->get_lattice(0)->get_size_cv();
  capi::Stage dl(lhs, n, true, true), dr(rhs, n, true, false);
  inversion_verbose_struct verb((inversion_verbose_level)verbosity, "[CAPI-MG]: ");
  h->mg->set_multigrid_level(0);
  StatefulMultigridMG::mg_preconditioner((capi_cd*)dl, (capi_cd*)dr, (int)n, (void*)h->mg, &verb);
}
// outer solve: VPGCR(restart) on level 0 preconditioned by the K-cycle
// (tests/n13_wilson_kcycle/wilson_kcycle.cpp:459).  restart = -1: unrestarted.
// info: resSq, iter, success, ops_count, seconds
void CAPI(mg_solve)(void* h_, int outer_type, capi_cd* x, const capi_cd* b, int max_iter, double tol, int restart, int verbosity, double* info)
{
  capi::MgH* h = (capi::MgH*)h_;
  Stencil2D* s0 = h->mg->get_stencil(0);
  const long full = h->mg->get_lattice(0)->get_size_cv();
  const int n = (int)((QMGStencilType)outer_type == QMG_MATVEC_RIGHT_SCHUR ? full / 2 : full);
  capi::Stage dx(x, full, true, true), db(b, full, true, false);
  inversion_verbose_struct verb((inversion_verbose_level)verbosity, "[CAPI-MG-OUTER]: ");
  verb.precond_verbosity = (inversion_verbose_level)verbosity;
  verb.precond_verb_prefix = "[CAPI-MG-PREC]: ";
  matrix_op_cplx fn = Stencil2D::get_apply_function((QMGStencilType)outer_type);
  h->mg->set_multigrid_level(0);
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  inversion_info inv;
  if (restart == -1)
    inv = minv_vector_gcr_var_precond((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, fn, (void*)s0,
                                      StatefulMultigridMG::mg_preconditioner, (void*)h->mg, &verb);
  else
    inv = minv_vector_gcr_var_precond_restart((capi_cd*)dx, (capi_cd*)db, n, max_iter, tol, restart, fn, (void*)s0,
                                              StatefulMultigridMG::mg_preconditioner, (void*)h->mg, &verb);
  capi_barrier();
  info[4] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  info[0] = inv.resSq; info[1] = inv.iter; info[2] = inv.success ? 1.0 : 0.0; info[3] = inv.ops_count;
}
// out: nullvec, krylov, presmooth, postsmooth op counts, total ops, krylov iterations
void CAPI(mg_tracker)(void* h_, int level, int* out)
{
  StatefulMultigridMG* mg = ((capi::MgH*)h_)->mg;
  out[0] = mg->get_tracker_count(QMG_DSLASH_TYPE_NULLVEC, level);
  out[1] = mg->get_tracker_count(QMG_DSLASH_TYPE_KRYLOV, level);
  out[2] = mg->get_tracker_count(QMG_DSLASH_TYPE_PRESMOOTH, level);
  out[3] = mg->get_tracker_count(QMG_DSLASH_TYPE_POSTSMOOTH, level);
  out[4] = mg->get_total_count(level);
  out[5] = mg->get_iterations_count(level);
}
void CAPI(mg_reset_tracker)(void* h_) { ((capi::MgH*)h_)->mg->reset_tracker(); }
// B200 extension: operator applications actually LAUNCHED at `level` (the trackers above keep the reference's counts, which
// include the A.0 and unread true-residual applies the fused K-cycle skips).  Reference build: the reference's total.
long CAPI(mg_executed)(void* h_, int level)
{
#ifdef QMG_B200_HOST
  return ((capi::MgH*)h_)->mg->get_executed_count(level);
#else
  return ((capi::MgH*)h_)->mg->get_total_count(level);
#endif
}
// B200 extension: 1 = fused K-cycle with the pre-smoother's residual handed over (default), 2 = fused K-cycle with the explicit
// residual (bit-identical to 0), 0 = the reference's sequence of separate sweeps.  Returns the previous setting.
int CAPI(mg_set_fused)(void* h_, int on)
{
#ifdef QMG_B200_HOST
  StatefulMultigridMG* mg = ((capi::MgH*)h_)->mg;
  const int was = mg->get_fused_cycle() ? (mg->get_residual_handover() ? 1 : 2) : 0;
  mg->set_fused_cycle(on != 0);
  if (on != 0) mg->set_residual_handover(on == 1);
  return was;
#else
  (void)h_; (void)on; return 0;
#endif
}
// emulated/explicit level operator (multigrid.h:465), prolong / restrict through the MG object
void CAPI(mg_apply_stencil)(void* h_, int level, int type, capi_cd* lhs, const capi_cd* rhs)
{
  capi::MgH* h = (capi::MgH*)h_;
  const long n = h->mg->get_lattice(level)->get_size_cv();
  capi::Stage dl(lhs, n, true, true), dr(rhs, n, true, false);
  h->mg->apply_stencil((capi_cd*)dl, (capi_cd*)dr, level, (QMGStencilType)type);
}
int CAPI(mg_storage_counts)(void* h_, int level, int* out)
{
  StatefulMultigridMG* mg = ((capi::MgH*)h_)->mg;
  out[0] = mg->get_storage_number_allocated(level);
  out[1] = mg->get_storage_number_checked(level);
  return 0;
}

// ------------------------------------------------------- whole K-cycle driver --
// The setup and solve of tests/n13_wilson_kcycle/wilson_kcycle.cpp (:226-471) as one call each, so that large
// lattices never round-trip vectors through the caller: Wilson operator from the given gauge field; per level
// coarse_dof/2 null vectors from BiCGstab-L(6) solves of A e = -A eta (eta gaussian from std::mt19937(seed), 500
// iterations, tol 5e-5), orthogonalised, chirally doubled and normalised; TransferMG with block
// orthonormalisation and QMG_DOUBLE_PROJECTION; Galerkin coarse operators; MR(pre, post) smoothing.
// iparams: n_refine, x_block, y_block, coarse_dof, pre_iters, post_iters, inner_iters, inner_restart,
//          coarsest_iters, coarsest_restart, null_max_iter, null_L, fine_stencil_app (all levels), coarsest_stencil_app,
//          fine operator (0 Wilson2D, 1 Staggered2D: nc = 1, null vectors doubled by the even / odd projection, staggered.h:176-181)
// dparams: inner_tol, coarsest_tol, null_tol, pre_tol, post_tol
namespace capi {
struct KCycleH
{
  std::vector<Lattice2D*> lats;
  Stencil2D* op;
  std::vector<TransferMG*> transfers;
  std::vector<StatefulMultigridMG::LevelSolveMG*> level_solves;
  StatefulMultigridMG::CoarsestSolveMG* coarsest;
  StatefulMultigridMG* mg;
  std::mt19937 generator;
  double setup_seconds;
  int null_ops;
  int ip[15]; double dp[5]; int verbosity;
};

// B200 extension (see kcycle_setup_link_compressed below): switch every operator to its link-compressed apply as soon as it
// exists, so that the null-vector solves of the set-up already run on clover / +x / +y blocks
inline int& setup_link_compressed() { static int on = 0; return on; }
// B200 extension (see kcycle_setup_matrix_free below): a Wilson fine operator applies matrix-free from the moment it is built
inline int& setup_matrix_free() { static int on = 0; return on; }
inline void maybe_link_compress(Stencil2D* s)
{
#ifdef QMG_B200_HOST
  // only where the shared-memory tile kernel exists (nc >= 4): at nc = 2 the link-compressed apply saves DRAM traffic but no time
  if (setup_link_compressed() && s != 0 && s->get_lattice()->get_nc() >= 4) s->enable_gamma5_hermitian_apply();
#else
  (void)s;
#endif
}

// null vectors -> TransferMG -> Galerkin coarse operator, level by level, on top of h->op (n13 :250-416, n16 :318-440)
inline void kcycle_build_hierarchy(KCycleH* h)
{
  const int* ip = h->ip; const double* dp = h->dp;
  const int n_refine = ip[0], xb = ip[1], yb = ip[2], coarse_dof = ip[3];
  const QMGStencilType level_app = (QMGStencilType)ip[12];
  const QMGStencilType coarsest_app = (QMGStencilType)ip[13];
  // a coarsest NORMAL-equation solve on the original operator needs M^dag on the coarse levels, nothing else
  const bool coarsest_normal_plain = (coarsest_app == QMG_MATVEC_M_MDAGGER || coarsest_app == QMG_MATVEC_MDAGGER_M);
  const bool need_rbj = (level_app != QMG_MATVEC_ORIGINAL) || (coarsest_app != QMG_MATVEC_ORIGINAL && !coarsest_normal_plain);
  if (need_rbj) h->op->build_rbjacobi_stencil();
  h->mg = new StatefulMultigridMG(h->lats[0], h->op, h->coarsest);
  inversion_verbose_struct verb((inversion_verbose_level)h->verbosity, "[CAPI-NULLVEC]: ");
  int cx = h->lats[0]->get_dim_mu(0), cy = h->lats[0]->get_dim_mu(1);
  for (int i = 1; i <= n_refine; i++)
  {
    cx /= xb; cy /= yb;
    h->lats.push_back(new Lattice2D(cx, cy, coarse_dof));
    Lattice2D* fl = h->lats[i - 1];
    const long nf = fl->get_size_cv();
    Stencil2D* fop = h->mg->get_stencil(i - 1);
    maybe_link_compress(fop);
    std::vector<capi_cd*> nv(coarse_dof);
    for (int j = 0; j < coarse_dof; j++) { nv[j] = capi_alloc(nf); zero_vector(nv[j], nf); }
    for (int j = 0; j < coarse_dof / 2; j++)
    {
      capi_cd* eta = h->mg->get_storage(i - 1)->check_out();
      gaussian(eta, nf, h->generator);
      for (int k = 0; k < j; k++) orthogonal(eta, nv[k], nf);
      capi_cd* Aeta = h->mg->get_storage(i - 1)->check_out();
      zero_vector(Aeta, nf);
      fop->apply_M(Aeta, eta);
      cax(-1.0, Aeta, nf);
      inversion_info inv = minv_vector_bicgstab_l(nv[j], Aeta, nf, ip[10], dp[2], ip[11], apply_stencil_2D_M, (void*)fop, &verb);
      h->null_ops += inv.ops_count + 1;
      cxpy(eta, nv[j], nf);
      h->mg->get_storage(i - 1)->check_in(eta);
      h->mg->get_storage(i - 1)->check_in(Aeta);
      for (int k = 0; k < j; k++) orthogonal(nv[j], nv[k], nf);
    }
    for (int j = 0; j < coarse_dof / 2; j++)
    {
      fop->chiral_projection_both(nv[j], nv[j + coarse_dof / 2]);
      normalize(nv[j], nf);
      normalize(nv[j + coarse_dof / 2], nf);
    }
    h->transfers.push_back(new TransferMG(fl, h->lats[i], &nv[0], true, false, QMG_DOUBLE_PROJECTION));
    StatefulMultigridMG::LevelSolveMG* ls = new StatefulMultigridMG::LevelSolveMG;
    ls->fine_stencil_app = level_app;
    ls->intermediate_tol = dp[0]; ls->intermediate_iters = ip[6]; ls->intermediate_restart_freq = ip[7];
    ls->pre_tol = dp[3]; ls->pre_iters = ip[4];
    ls->post_tol = dp[4]; ls->post_iters = ip[5];
    h->level_solves.push_back(ls);
    h->mg->push_level(h->lats[i], h->transfers[i - 1], ls, true, true,
                      level_app == QMG_MATVEC_ORIGINAL ? MultigridMG::QMG_MULTIGRID_PRECOND_ORIGINAL : MultigridMG::QMG_MULTIGRID_PRECOND_RIGHT_BLOCK_JACOBI,
                      need_rbj ? CoarseOperator2D::QMG_COARSE_BUILD_RBJACOBI
                               : (coarsest_normal_plain ? CoarseOperator2D::QMG_COARSE_BUILD_DAGGER : CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL), (capi_cd**)0);
    for (int j = 0; j < coarse_dof; j++) capi_free(nv[j]);
  }
}
// everything above level 0 (the fine operator and its lattice stay)
inline void kcycle_drop_hierarchy(KCycleH* h)
{
  delete h->mg; h->mg = 0;
  for (size_t i = 0; i < h->transfers.size(); i++) delete h->transfers[i];
  for (size_t i = 0; i < h->level_solves.size(); i++) delete h->level_solves[i];
  for (size_t i = 1; i < h->lats.size(); i++) delete h->lats[i];
  h->transfers.clear(); h->level_solves.clear(); h->lats.resize(1);
}
}

void* CAPI(kcycle_new)(int X, int Y, double mass, const capi_cd* gauge, const int* ip, const double* dp, unsigned seed, int verbosity)
{
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  capi::KCycleH* h = new capi::KCycleH;
  h->generator.seed(seed);
  h->null_ops = 0;
  for (int i = 0; i < 15; i++) h->ip[i] = ip[i];
  for (int i = 0; i < 5; i++) h->dp[i] = dp[i];
  h->verbosity = verbosity;
  const bool staggered = (ip[14] == 1);
  h->lats.push_back(new Lattice2D(X, Y, staggered ? 1 : 2));
  {
    capi::Stage g(gauge, 2L * X * Y, true, false);
    if (staggered) h->op = new Staggered2D(h->lats[0], capi_cd(mass, 0.0), (capi_cd*)g);
    else
    {
      Wilson2D* w = new Wilson2D(h->lats[0], capi_cd(mass, 0.0), (capi_cd*)g);
      h->op = w;
#ifdef QMG_B200_HOST
      if (capi::setup_matrix_free()) w->enable_matrix_free_apply((capi_cd*)g);
#endif
    }
  }
  h->coarsest = new StatefulMultigridMG::CoarsestSolveMG;
  h->coarsest->coarsest_stencil_app = (QMGStencilType)ip[13];
  h->coarsest->coarsest_tol = dp[1];
  h->coarsest->coarsest_iters = ip[8];
  h->coarsest->coarsest_restart_freq = ip[9];
  capi::kcycle_build_hierarchy(h);
  capi_barrier();
  h->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
#ifdef QMG_B200_HOST
  // the set-up's work vectors (14 per BiCGstab-6 solve, the raw null vectors) are parked in the allocator's cache: hand them
  // back to the driver so that the solve's own vectors -- other sizes, other counts -- have the room
  QMG_CHK(qmg_trim());
#endif
  return h;
}
// ---- adaptive set-up (tests/n22_wilson_kcycle_adaptive/wilson_kcycle.cpp:226-440): test vectors relaxed with
// Richardson(10, omega = 0.33) build a first hierarchy, then n_setup rounds re-solve every test vector with 10 iterations
// of flexible GCR preconditioned by the CURRENT K-cycle (on coarser levels: starting from the restriction of the
// vector one level up), rebuild that level's transfer / coarse operator (update_level) and re-relax everything below.
namespace capi {
typedef std::vector<std::vector<capi_cd*> > TestVectors;

// relax fresh gaussian vectors on level `fine` and build (or update) level fine + 1 from them (n22 :620-705)
inline TransferMG* adaptive_build_below(KCycleH* h, TestVectors& test, int fine, StatefulMultigridMG::LevelSolveMG* ls, bool fresh, inversion_verbose_struct* verb)
{
  Lattice2D* fl = h->mg->get_lattice(fine); Lattice2D* cl = h->lats[fine + 1];
  const int coarse_dof = cl->get_nc(); const long nf = fl->get_size_cv();
  std::vector<capi_cd*> nv(coarse_dof);
  for (int j = 0; j < coarse_dof / 2; j++)
  {
    nv[j] = capi_alloc(nf); nv[j + coarse_dof / 2] = capi_alloc(nf);
    zero_vector(nv[j], nf); zero_vector(nv[j + coarse_dof / 2], nf);
    capi_cd* rnd = h->mg->get_storage(fine)->check_out();
    gaussian(rnd, nf, h->generator);
    inversion_info inv = minv_vector_richardson(test[fine][j], rnd, nf, 10, 1e-10, 0.33, 250, apply_stencil_2D_M, (void*)h->mg->get_stencil(fine), verb);
    h->mg->add_tracker_count(QMG_DSLASH_TYPE_NULLVEC, inv.ops_count, fine);
    h->null_ops += inv.ops_count;
    h->mg->get_storage(fine)->check_in(rnd);
    for (int k = 0; k < j; k++) orthogonal(test[fine][j], test[fine][k], nf);
    normalize(test[fine][j], nf);
    copy_vector(nv[j], test[fine][j], nf);
    h->mg->get_stencil(fine)->chiral_projection_both(nv[j], nv[j + coarse_dof / 2]);
  }
  // level 0 states the doubling (n22 :308); the levels below use the 4-argument constructor (n22 :680)
  TransferMG* tr = (fine == 0) ? new TransferMG(fl, cl, &nv[0], true, false, QMG_DOUBLE_PROJECTION) : new TransferMG(fl, cl, &nv[0], true);
  if (fresh) h->mg->push_level(cl, tr, ls, true, true, MultigridMG::QMG_MULTIGRID_PRECOND_ORIGINAL, &nv[0]);
  else h->mg->update_level(fine + 1, cl, tr, ls, true, true, MultigridMG::QMG_MULTIGRID_PRECOND_ORIGINAL, &nv[0]);
  for (int j = 0; j < coarse_dof; j++) capi_free(nv[j]);
  return tr;
}
}

// iparams / dparams as kcycle_new (null_max_iter, null_L, null_tol unused: the relaxation is Richardson); Wilson only.
void* CAPI(kcycle_new_adaptive)(int X, int Y, double mass, const capi_cd* gauge, const int* ip, const double* dp, int n_setup, unsigned seed, int verbosity)
{
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  capi::KCycleH* h = new capi::KCycleH;
  h->generator.seed(seed);
  h->null_ops = 0;
  for (int i = 0; i < 15; i++) h->ip[i] = ip[i];
  for (int i = 0; i < 5; i++) h->dp[i] = dp[i];
  h->verbosity = verbosity;
  const int n_refine = ip[0], xb = ip[1], yb = ip[2], coarse_dof = ip[3];
  h->lats.push_back(new Lattice2D(X, Y, 2));
  {
    capi::Stage g(gauge, 2L * X * Y, true, false);
    h->op = new Wilson2D(h->lats[0], capi_cd(mass, 0.0), (capi_cd*)g);
  }
  h->coarsest = new StatefulMultigridMG::CoarsestSolveMG;
  h->coarsest->coarsest_stencil_app = QMG_MATVEC_ORIGINAL;
  h->coarsest->coarsest_tol = dp[1];
  h->coarsest->coarsest_iters = ip[8];
  h->coarsest->coarsest_restart_freq = ip[9];
  h->mg = new StatefulMultigridMG(h->lats[0], h->op, h->coarsest);
  inversion_verbose_struct verb((inversion_verbose_level)verbosity, "[CAPI-ADAPTIVE]: ");
  int cx = X, cy = Y;
  capi::TestVectors test(n_refine);
  for (int i = 0; i < n_refine; i++)
  {
    cx /= xb; cy /= yb;
    h->lats.push_back(new Lattice2D(cx, cy, coarse_dof));
    const long nf = h->lats[i]->get_size_cv();
    test[i].resize(coarse_dof / 2);
    for (int j = 0; j < coarse_dof / 2; j++) { test[i][j] = capi_alloc(nf); zero_vector(test[i][j], nf); }
    // while setting up, every level below the top runs 8 unrestarted flexible-GCR iterations (n22 :248-256)
    StatefulMultigridMG::LevelSolveMG* ls = new StatefulMultigridMG::LevelSolveMG;
    ls->fine_stencil_app = QMG_MATVEC_ORIGINAL;
    ls->intermediate_tol = 1e-10; ls->intermediate_iters = 8; ls->intermediate_restart_freq = 1024;
    ls->pre_tol = dp[3]; ls->pre_iters = ip[4];
    ls->post_tol = dp[4]; ls->post_iters = ip[5];
    h->level_solves.push_back(ls);
  }
  h->transfers.resize(n_refine, (TransferMG*)0);
  for (int i = 0; i < n_refine; i++) h->transfers[i] = capi::adaptive_build_below(h, test, i, h->level_solves[i], true, &verb);

  for (int m = 0; m < n_setup; m++)
  {
    for (int i = 0; i < n_refine; i++)
    {
      Lattice2D* fl = h->lats[i]; Lattice2D* cl = h->lats[i + 1];
      const long nf = fl->get_size_cv();
      std::vector<capi_cd*> nv(coarse_dof);
      for (int j = 0; j < coarse_dof / 2; j++)
      {
        nv[j] = capi_alloc(nf); nv[j + coarse_dof / 2] = capi_alloc(nf);
        capi_cd* rhs = h->mg->get_storage(i)->check_out();
        if (i == 0) copy_vector(rhs, test[0][j], nf);
        else { zero_vector(rhs, nf); h->mg->get_transfer(i - 1)->restrict_f2c(test[i - 1][j], rhs); }
        zero_vector(test[i][j], nf);
        inversion_info inv = minv_vector_gcr_var_precond(test[i][j], rhs, nf, 10, 1e-10, apply_stencil_2D_M, (void*)h->mg->get_stencil(i),
                                                         StatefulMultigridMG::mg_preconditioner, (void*)h->mg, &verb);
        h->mg->get_storage(i)->check_in(rhs);
        h->mg->add_tracker_count(QMG_DSLASH_TYPE_NULLVEC, inv.ops_count + 1, i);
        h->null_ops += inv.ops_count + 1;
        for (int k = 0; k < j; k++) orthogonal(test[i][j], test[i][k], nf);
        normalize(test[i][j], nf);
        zero_vector(nv[j + coarse_dof / 2], nf);
        copy_vector(nv[j], test[i][j], nf);
        h->mg->get_stencil(i)->chiral_projection_both(nv[j], nv[j + coarse_dof / 2]);
      }
      delete h->transfers[i];
      h->transfers[i] = new TransferMG(fl, cl, &nv[0], true, false, QMG_DOUBLE_PROJECTION);
      h->mg->update_level(i + 1, cl, h->transfers[i], h->level_solves[i], true, true, MultigridMG::QMG_MULTIGRID_PRECOND_ORIGINAL, &nv[0]);
      for (int j = i + 1; j < n_refine; j++)
      {
        delete h->transfers[j];
        h->transfers[j] = capi::adaptive_build_below(h, test, j, h->level_solves[j], false, &verb);
      }
      for (int j = 0; j < coarse_dof; j++) capi_free(nv[j]);
      if (i < n_refine - 1) h->mg->go_coarser();
    }
    for (int i = 0; i < n_refine - 1; i++) h->mg->go_finer();
  }
  for (int i = 0; i <= n_refine; i++) h->mg->shift_all_to_nullvec(i);
  for (int i = 0; i < n_refine; i++)
  {
    // the solve uses the n13 parameters again (n22 :420-433)
    StatefulMultigridMG::LevelSolveMG* ls = h->level_solves[i];
    ls->intermediate_tol = dp[0]; ls->intermediate_iters = ip[6]; ls->intermediate_restart_freq = ip[7];
  }
  for (int i = 0; i < n_refine; i++)
    for (size_t j = 0; j < test[i].size(); j++) capi_free(test[i][j]);
  capi_barrier();
  h->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return h;
}

// The measurement-loop step of tests/n16_wilson_kcycle_heatbath/wilson_kcycle_heatbath.cpp:300-440: new links into the
// SAME Wilson operator (Wilson2D::update_links drops its derived link sets, wilson.h:211-225), then a fresh hierarchy.
void CAPI(kcycle_update_links)(void* h_, const capi_cd* gauge)
{
  capi::KCycleH* h = (capi::KCycleH*)h_;
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  capi::kcycle_drop_hierarchy(h);
  {
    capi::Stage g(gauge, 2L * h->lats[0]->get_volume(), true, false);
    if (h->ip[14] == 1) static_cast<Staggered2D*>(h->op)->update_links((capi_cd*)g);
    else static_cast<Wilson2D*>(h->op)->update_links((capi_cd*)g);
  }
  capi::kcycle_build_hierarchy(h);
  capi_barrier();
  h->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// The would-be pion correlator from a point source at (x0, y0): one K-cycle-preconditioned solve per spin component,
// |prop|^2 summed over each time slice and folded about the source (n16 :452-506).  pion: Y doubles.
// info: total outer iterations, all converged (1/0), seconds
void CAPI(kcycle_pion)(void* h_, int x0, int y0, int max_iter, double tol, int restart, int verbosity, double* pion, double* info)
{
  capi::KCycleH* h = (capi::KCycleH*)h_;
  Lattice2D* l0 = h->lats[0];
  const long n = l0->get_size_cv();
  const int Y = l0->get_dim_mu(1);
  capi_cd* src = h->mg->check_out(0); capi_cd* prop = h->mg->check_out(0);
  std::vector<capi_cd> hsrc((size_t)n);
  std::vector<double> part((size_t)Y);
  inversion_verbose_struct verb((inversion_verbose_level)verbosity, "[CAPI-PION]: ");
  for (int t = 0; t < Y; t++) pion[t] = 0.0;
  info[0] = 0.0; info[1] = 1.0;
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  for (int spin = 0; spin < l0->get_nc(); spin++)
  {
    std::fill(hsrc.begin(), hsrc.end(), capi_cd(0.0, 0.0));
    hsrc[l0->cv_coord_to_index(x0, y0, spin)] = 1.0;
    capi_put(src, &hsrc[0], n);
    zero_vector(prop, n);
    h->mg->set_multigrid_level(0);
    inversion_info inv = minv_vector_gcr_var_precond_restart(prop, src, (int)n, max_iter, tol, restart, apply_stencil_2D_M, (void*)h->op,
                                                             StatefulMultigridMG::mg_preconditioner, (void*)h->mg, &verb);
    info[0] += inv.iter; if (!inv.success) info[1] = 0.0;
    norm2sq_cv_timeslice(&part[0], prop, l0);
    for (int t = 0; t < Y; t++)
    {
      const int mirror = (2 * y0 - t + 2 * Y) % Y;     // the slice at the same distance on the other side of the source
      pion[t] += 0.5 * (part[t] + part[mirror]);
    }
  }
  capi_barrier();
  info[2] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  h->mg->check_in(src, 0); h->mg->check_in(prop, 0);
}
void CAPI(kcycle_free)(void* h_)
{
  capi::KCycleH* h = (capi::KCycleH*)h_;
  delete h->mg;
  for (size_t i = 0; i < h->transfers.size(); i++) delete h->transfers[i];
  for (size_t i = 0; i < h->level_solves.size(); i++) delete h->level_solves[i];
  delete h->coarsest;
  delete h->op;
  for (size_t i = 0; i < h->lats.size(); i++) delete h->lats[i];
  delete h;
}
// Outer solve A x = b with restarted flexible GCR preconditioned by the K-cycle (n13 :459-471).
// b: host array, or NULL for a gaussian right-hand side drawn from the handle's generator (as n13 :419 does).
// x_out: host array for the solution or NULL.  outer_type selects the system the outer solver sees (n19: Schur).
// info: resSq, iter, success, ops_count, solve seconds, explicit |b - A x| / |b| on the ORIGINAL system, setup seconds, null-vector ops
void CAPI(kcycle_solve)(void* h_, const capi_cd* b, capi_cd* x_out, int outer_type, int max_iter, double tol, int restart, int verbosity, double* info)
{
  capi::KCycleH* h = (capi::KCycleH*)h_;
  Lattice2D* l0 = h->lats[0];
  const long n = l0->get_size_cv();
  const QMGStencilType type = (QMGStencilType)outer_type;
  capi_cd* bd = h->mg->check_out(0);
  if (b != 0) capi_put(bd, b, n); else gaussian(bd, n, h->generator);
  const double bnorm = sqrt(norm2sq(bd, n));
  capi_cd* x = h->mg->check_out(0);
  capi_cd* bprep = h->mg->check_out(0);
  capi_cd* xfull = h->mg->check_out(0);
  zero_vector(x, n); zero_vector(bprep, n); zero_vector(xfull, n);
  h->op->prepare_M(bprep, bd, type);
  inversion_verbose_struct verb((inversion_verbose_level)verbosity, "[CAPI-KCYCLE]: ");
  verb.precond_verbosity = (inversion_verbose_level)verbosity;
  verb.precond_verb_prefix = "[CAPI-KCYCLE-PREC]: ";
  h->mg->set_multigrid_level(0);
  h->mg->reset_tracker();
  const int nsolve = (int)(type == QMG_MATVEC_RIGHT_SCHUR ? n / 2 : n);
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  inversion_info inv = minv_vector_gcr_var_precond_restart(x, bprep, nsolve, max_iter, tol, restart, Stencil2D::get_apply_function(type), (void*)h->op,
                                                           StatefulMultigridMG::mg_preconditioner, (void*)h->mg, &verb);
  capi_barrier();
  info[4] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  h->op->reconstruct_M(xfull, x, bd, type);
  capi_cd* Ax = bprep;
  zero_vector(Ax, n);
  h->op->apply_M(Ax, xfull);
  info[5] = sqrt(diffnorm2sq(bd, Ax, n)) / bnorm;
  info[0] = inv.resSq; info[1] = inv.iter; info[2] = inv.success ? 1.0 : 0.0; info[3] = inv.ops_count;
  info[6] = h->setup_seconds; info[7] = h->null_ops;
  if (x_out != 0) capi_get(x_out, xfull, n);
  h->mg->check_in(bd, 0); h->mg->check_in(x, 0); h->mg->check_in(bprep, 0); h->mg->check_in(xfull, 0);
}
// B200 extension: 1 = hierarchies built from now on switch each level's operator to the link-compressed apply the moment it is
// built (checked per level, like kcycle_gamma5_hermitian), so the BiCGstab-L null-vector solves of the set-up use it too.
// Reference build: no-op.  Returns the previous setting.
int CAPI(kcycle_setup_link_compressed)(int on)
{
  const int was = capi::setup_link_compressed();
  capi::setup_link_compressed() = on ? 1 : 0;
  return was;
}
// B200 extension: 1 = K-cycle objects built from now on apply their Wilson fine operator matrix-free (links instead of stored
// blocks, same bits; Wilson2D::enable_matrix_free_apply) -- in the set-up's null-vector solves as well as in the solves.
// Reference build: no-op.  Returns the previous setting.
int CAPI(kcycle_setup_matrix_free)(int on)
{
  const int was = capi::setup_matrix_free();
  capi::setup_matrix_free() = on ? 1 : 0;
  return was;
}
// B200 extension: pause (on = 0) / resume (on = 1) the matrix-free apply of the fine operator of this K-cycle object; returns 1
// when it is active afterwards.  Reference build: 0.
int CAPI(kcycle_matrix_free)(void* h_, int on)
{
#ifdef QMG_B200_HOST
  capi::KCycleH* h = (capi::KCycleH*)h_;
  if (h->op == 0) return 0;
  h->op->pause_matrix_free_apply(on == 0);
  return h->op->uses_matrix_free_apply() ? 1 : 0;
#else
  (void)h_; (void)on; return 0;
#endif
}
// B200 extension: link-compressed applies on every level of the hierarchy whose stored blocks obey the gamma5-hermitian
// relation (checked per level); returns the number of levels switched.  Reference build: 0.
int CAPI(kcycle_gamma5_hermitian)(void* h_, int on)
{
#ifdef QMG_B200_HOST
  capi::KCycleH* h = (capi::KCycleH*)h_;
  int count = 0;
  for (int l = 0; l < h->mg->get_num_levels(); l++)
  {
    Stencil2D* s = h->mg->get_stencil(l);
    if (s == 0) continue;
    // on == 2: only where the shared-memory patch kernels exist (nc >= 4); at nc = 2 the link-compressed apply saves DRAM
    // traffic but is slower than the stored-block one (4.16 vs 3.64 ms on 8192^2)
    if (!on || (on == 2 && s->get_lattice()->get_nc() < 4)) s->disable_gamma5_hermitian_apply();
    else if (s->enable_gamma5_hermitian_apply()) count++;
  }
  return count;
#else
  (void)h_; (void)on; return 0;
#endif
}

// B200 build only (the reference needs ARPACK for this, absent here): StatefulMultigridMG::deflate_coarsest
// (multigrid/stateful_multigrid.h:613) -- num_low / num_high eigenpairs of the coarsest normal operator; evals_out
// (num_low + num_high doubles, may be NULL) receives the eigenvalues.  Returns the size of the deflation space.
int CAPI(kcycle_deflate_coarsest)(void* h_, int num_low, int num_high, double* evals_out)
{
#ifdef QMG_B200_HOST
  capi::KCycleH* h = (capi::KCycleH*)h_;
  h->mg->clear_deflation();
  h->mg->deflate_coarsest(num_low, num_high, false);
  const int n = (int)h->mg->get_coarsest_deflated();
  if (evals_out != 0) for (int i = 0; i < n; i++) evals_out[i] = real(h->mg->get_coarsest_evals()[i]);
  return n;
#else
  (void)h_; (void)num_low; (void)num_high; (void)evals_out; return 0;
#endif
}
// B200 build only: nev eigenpairs at the low (which = 0) or high (1) end of the spectrum of the HERMITIAN operator
// `type` of this stencil (e.g. QMG_MATVEC_MDAGGER_M) through arpack_dcn; evecs_out: nev * size_cv or NULL.
// Returns 1 on success.
int CAPI(stencil_eigs)(void* h_, int type, int nev, int ncv, int which, double tol, double* evals_out, capi_cd* evecs_out)
{
#ifdef QMG_B200_HOST
  Stencil2D* s = ((capi::StencilH*)h_)->op;
  const int n = s->lat->get_size_cv();
  arpack_dcn eig(n, 100000, tol, Stencil2D::get_apply_function((QMGStencilType)type), (void*)s, nev, ncv);
  const arpack_dcn::arpack_spectrum_piece piece = which ? arpack_dcn::ARPACK_LARGEST_REAL : arpack_dcn::ARPACK_SMALLEST_REAL;
  if (!eig.prepare_eigensystem(piece, nev, ncv)) return 0;
  std::vector<capi_cd> ev(nev);
  std::vector<capi_cd*> vec(nev);
  for (int i = 0; i < nev; i++) vec[i] = capi_alloc(n);
  const bool ok = eig.get_eigensystem(&ev[0], &vec[0], piece);
  for (int i = 0; i < nev; i++)
  {
    evals_out[i] = real(ev[i]);
    if (ok && evecs_out != 0) capi_get(evecs_out + (long)i * n, vec[i], n);
    capi_free(vec[i]);
  }
  return ok ? 1 : 0;
#else
  (void)h_; (void)type; (void)nev; (void)ncv; (void)which; (void)tol; (void)evals_out; (void)evecs_out; return 0;
#endif
}

void* CAPI(kcycle_mg)(void* h_)
{
  capi::KCycleH* h = (capi::KCycleH*)h_;
  capi::MgH* m = new capi::MgH; m->mg = h->mg; m->coarsest = 0;   // borrowed view for mg_tracker & co; never pass to mg_free
  return m;
}
// Wall-clock seconds of `reps` K-cycle applications (mg_preconditioner at level 0) on a gaussian vector.
double CAPI(kcycle_time_precond)(void* h_, int warm, int reps)
{
  capi::KCycleH* h = (capi::KCycleH*)h_;
  const long n = h->lats[0]->get_size_cv();
  capi_cd* r = h->mg->check_out(0); capi_cd* z = h->mg->check_out(0);
  gaussian(r, n, h->generator);
  inversion_verbose_struct verb;
  h->mg->set_multigrid_level(0);
  for (int i = 0; i < warm; i++) { zero_vector(z, n); StatefulMultigridMG::mg_preconditioner(z, r, (int)n, (void*)h->mg, &verb); }
  capi_barrier();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; i++) { zero_vector(z, n); StatefulMultigridMG::mg_preconditioner(z, r, (int)n, (void*)h->mg, &verb); }
  capi_barrier();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  h->mg->check_in(r, 0); h->mg->check_in(z, 0);
  return sec;
}

} // extern "C"
