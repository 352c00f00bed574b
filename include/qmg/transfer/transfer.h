// quantum-mg on B200 -- TransferMG: aggregation-based prolongator P and restrictor R = P^dagger (or a separate R)
// between a fine and a coarse Lattice2D (/root/reference/transfer/transfer.h:42-820).
//
// The reference stores a sorted fine-index list per coarse site (build_mapping, :410-448) and expresses the block
// Gram-Schmidt as nvec (nvec + 1) / 2 single-vector restrict / prolong sweeps over the lattice (:514-607).  Here the
// aggregation is arithmetic on the even-odd layout inside the kernels, prolong / restrict are one streaming pass
// over the null vectors each, and one warp per aggregate orthonormalises its (fine dof x nvec) panel in shared
// memory (quantum-mg_b200/csrc/qmg_transfer.cu).  null_vectors / restrict_null_vectors stay public device arrays.
#ifndef QMG_B200_TRANSFER
#define QMG_B200_TRANSFER

#include <cstdlib>
#include <iostream>
#include <complex>
#include <vector>
#include "blas/generic_vector.h"
#include "lattice/lattice.h"

enum QMGDoublingType
{
  QMG_DOUBLE_NONE = 0,
  QMG_DOUBLE_PROJECTION = 1,   // null vectors doubled with chiral projectors
  QMG_DOUBLE_OPERATOR = 2,     // doubled with gamma_5^{L/R}
};

class TransferMG
{
private:
  TransferMG(TransferMG const&);
  TransferMG& operator=(TransferMG const&);

  Lattice2D* fine_lat;
  Lattice2D* coarse_lat;
  int num_null_vec;
  int* blocksizes;
  int fine_sites_per_coarse;
  QMGDoublingType doubling;
  bool is_init;
  qmg_transfer_desc desc;
  // chirality-packed copy of the null vectors for prolong / restrict (see use_packed), built on first use
  complex<double>* packed;
  int packed_state;            // 0 not tried yet, 1 in use, -1 not applicable

public:
  complex<double>** null_vectors;            // [coarse nc][fine size_cv], device
  complex<double>** restrict_null_vectors;   // 0 when R = P^dagger
  complex<double>* block_cholesky;           // per coarse site nc x nc factor of the block Gram matrix, or 0
  complex<double>* block_L;
  complex<double>* block_U;

private:
  complex<double>** clone_vectors(complex<double>** in)
  {
    complex<double>** out = new complex<double>*[num_null_vec];
    for (int i = 0; i < num_null_vec; i++)
    {
      out[i] = allocate_vector<complex<double> >(fine_lat->get_size_cv());
      copy_vector(out[i], in[i], fine_lat->get_size_cv());
    }
    return out;
  }

  // generic flavours (public ones below use every vector): fine += sum_v nv_v coarse[.][v], coarse[.][v] += <nv_v | fine>_aggregate
  void prolong_c2f(complex<double>* coarse_cv, complex<double>* fine_cv, complex<double>** vecs, int nvec)
  { QMG_CHK(qmg_prolong(&desc, reinterpret_cast<const qmg_cplx* const*>(vecs), nvec, qmg_host::P(coarse_cv), qmg_host::P(fine_cv))); }
  void restrict_f2c(complex<double>* fine_cv, complex<double>* coarse_cv, complex<double>** vecs, int nvec)
  { QMG_CHK(qmg_restrict(&desc, reinterpret_cast<const qmg_cplx* const*>(vecs), nvec, qmg_host::P(fine_cv), qmg_host::P(coarse_cv))); }

  // B200 extension: with QMG_DOUBLE_PROJECTION null vector j carries the upper, j + nc/2 the lower chirality of one solve, so at
  // every fine element half of the vectors are zero.  The first whole-set prolong / restrict packs the non-zero halves into one
  // interleaved array (half the size of the set) after CHECKING that what it drops is exactly zero, and from then on the
  // whole-set operations read that copy: 96 / 80 instead of 160 / 144 bytes per fine dof.  Writing through the public
  // null_vectors pointers afterwards must be followed by drop_packed().  QMG_PACKED_TRANSFER=0 switches it off.
  bool use_packed()
  {
    if (packed_state != 0) return packed_state > 0;
    packed_state = -1;
    const char* e = getenv("QMG_PACKED_TRANSFER");
    if (e != 0 && e[0] == '0') return false;
    if (!is_init || doubling != QMG_DOUBLE_PROJECTION || restrict_null_vectors != 0 || !qmg_transfer_packed_supported(&desc)) return false;
    packed = allocate_vector<complex<double> >((long)fine_lat->get_size_cv() * (num_null_vec / 2));
    double dropped = 1.0;
    QMG_CHK(qmg_transfer_pack_chiral(&desc, reinterpret_cast<const qmg_cplx* const*>(null_vectors), num_null_vec, qmg_host::P(packed), &dropped));
    if (dropped != 0.0) { deallocate_vector(&packed); return false; }
    packed_state = 1;
    return true;
  }

  // one Gram-Schmidt pass over every aggregate (transfer.h:514-607); the factor is stored when block_cholesky != 0
  void block_orthonormalize()
  { QMG_CHK(qmg_block_orthonormalize(&desc, reinterpret_cast<qmg_cplx* const*>(null_vectors), num_null_vec, qmg_host::P(block_cholesky))); }

  // Bi-orthonormalisation of (P, R) with optional LU factors (transfer.h:610-769), written like the reference
  // as single-vector restrict / prolong sweeps: it is only reached through the asymmetric constructor.
  void block_bi_orthonormalize()
  {
    const long nf = fine_lat->get_size_cv(), ncv = coarse_lat->get_size_cv(), vol_c = coarse_lat->get_volume();
    const int nd = num_null_vec;
    complex<double>* fv = allocate_vector<complex<double> >(nf);
    complex<double>* cv = allocate_vector<complex<double> >(ncv);
    for (int i = 0; i < nd; i++)
    {
      for (int j = 0; j < i; j++)
      {
        // P_i -= <R_j|P_i> P_j
        zero_vector(fv, nf); zero_vector(cv, ncv);
        restrict_f2c(null_vectors[i], cv, &restrict_null_vectors[j], 1);
        if (block_U != 0) copy_vector_blas(block_U + j * nd + i, nd * nd, cv, nd, vol_c);
        prolong_c2f(cv, fv, &null_vectors[j], 1);
        caxpy(-1.0, fv, null_vectors[i], nf);
        // R_i -= <P_j|R_i> R_j
        zero_vector(fv, nf); zero_vector(cv, ncv);
        restrict_f2c(restrict_null_vectors[i], cv, &null_vectors[j], 1);
        if (block_L != 0) copy_vector_blas(block_L + i * nd + j, nd * nd, cv, nd, vol_c);
        prolong_c2f(cv, fv, &restrict_null_vectors[j], 1);
        caxpy(-1.0, fv, restrict_null_vectors[i], nf);
      }
      // split <R_i|P_i> = |d| e^{i phi}: R_i *= e^{i phi}/sqrt|d|, P_i /= sqrt|d|
      zero_vector(fv, nf); zero_vector(cv, ncv);
      restrict_f2c(null_vectors[i], cv, &restrict_null_vectors[i], 1);
      QMG_CHK(qmg_elementwise(3, qmg_host::P(cv), ncv));
      if (block_L != 0) { cinvx(cv, ncv); copy_vector_blas(block_L + i * (nd + 1), nd * nd, cv, nd, vol_c); cinvx(cv, ncv); }
      prolong_c2f(cv, fv, &restrict_null_vectors[i], 1);
      copy_vector(restrict_null_vectors[i], fv, nf);
      zero_vector(fv, nf);
      abs_vector(cv, ncv);
      if (block_U != 0) { cinvx(cv, ncv); copy_vector_blas(block_U + i * (nd + 1), nd * nd, cv, nd, vol_c); cinvx(cv, ncv); }
      prolong_c2f(cv, fv, &null_vectors[i], 1);
      copy_vector(null_vectors[i], fv, nf);
    }
    if (block_L != 0) conj_vector(block_L, coarse_lat->get_size_cm());
    deallocate_vector(&cv);
    deallocate_vector(&fv);
  }

  bool setup_geometry()
  {
    blocksizes = new int[2];
    fine_sites_per_coarse = fine_lat->get_nc();
    for (int mu = 0; mu < 2; mu++)
    {
      if (fine_lat->get_dim_mu(mu) % coarse_lat->get_dim_mu(mu) != 0)
      {
        std::cout << "[QMG-ERROR]: Fine lattice dimension " << mu << "isn't divided evenly by coarse dimension.\n";
        return false;
      }
      blocksizes[mu] = fine_lat->get_dim_mu(mu) / coarse_lat->get_dim_mu(mu);
      fine_sites_per_coarse *= blocksizes[mu];
    }
    desc.Xf = fine_lat->get_dim_mu(0); desc.Yf = fine_lat->get_dim_mu(1); desc.ncf = fine_lat->get_nc();
    desc.Xc = coarse_lat->get_dim_mu(0); desc.Yc = coarse_lat->get_dim_mu(1); desc.ncc = coarse_lat->get_nc();
    return true;
  }

public:
  // R = P^dagger.  Copies the caller's null vectors, then block-orthonormalises them twice (transfer.h:118-179).
  TransferMG(Lattice2D* in_fine_lat, Lattice2D* in_coarse_lat, complex<double>** in_null_vectors, bool do_block_ortho = true,
             bool save_decomp = false, QMGDoublingType in_doubling = QMG_DOUBLE_NONE)
    : fine_lat(in_fine_lat), coarse_lat(in_coarse_lat), num_null_vec(in_coarse_lat->get_nc()), blocksizes(0), fine_sites_per_coarse(0),
      doubling(in_doubling), is_init(false), packed(0), packed_state(0), null_vectors(0), restrict_null_vectors(0), block_cholesky(0), block_L(0), block_U(0)
  {
    if (!setup_geometry()) return;
    null_vectors = clone_vectors(in_null_vectors);
    if (save_decomp)
    {
      block_cholesky = allocate_vector<complex<double> >(coarse_lat->get_size_cm());
      zero_vector(block_cholesky, coarse_lat->get_size_cm());
    }
    if (do_block_ortho)
    {
      block_orthonormalize();
      // the stored factor is the one of the FIRST pass
      complex<double>* keep = block_cholesky;
      block_cholesky = 0;
      block_orthonormalize();
      block_cholesky = keep;
    }
    is_init = true;
  }

  // separate prolong and restrict vectors (transfer.h:185-225)
  TransferMG(Lattice2D* in_fine_lat, Lattice2D* in_coarse_lat, complex<double>** in_prolong_null_vectors, complex<double>** in_restrict_null_vectors,
             bool do_block_bi_ortho = true, bool save_decomp = false, QMGDoublingType in_doubling = QMG_DOUBLE_NONE)
    : fine_lat(in_fine_lat), coarse_lat(in_coarse_lat), num_null_vec(in_coarse_lat->get_nc()), blocksizes(0), fine_sites_per_coarse(0),
      doubling(in_doubling), is_init(false), packed(0), packed_state(0), null_vectors(0), restrict_null_vectors(0), block_cholesky(0), block_L(0), block_U(0)
  {
    if (!setup_geometry()) return;
    null_vectors = clone_vectors(in_prolong_null_vectors);
    restrict_null_vectors = clone_vectors(in_restrict_null_vectors);
    if (save_decomp)
    {
      block_L = allocate_vector<complex<double> >(coarse_lat->get_size_cm());
      block_U = allocate_vector<complex<double> >(coarse_lat->get_size_cm());
      zero_vector(block_L, coarse_lat->get_size_cm());
      zero_vector(block_U, coarse_lat->get_size_cm());
    }
    if (do_block_bi_ortho)
    {
      block_bi_orthonormalize();
      complex<double>* keepL = block_L; complex<double>* keepU = block_U;
      block_L = 0; block_U = 0;
      block_bi_orthonormalize();
      block_L = keepL; block_U = keepU;
    }
    is_init = true;
  }

  ~TransferMG()
  {
    if (blocksizes != 0) delete[] blocksizes;
    complex<double>*** sets[] = { &null_vectors, &restrict_null_vectors };
    for (int s = 0; s < 2; s++)
    {
      complex<double>** v = *sets[s];
      if (v == 0) continue;
      for (int i = 0; i < num_null_vec; i++) if (v[i] != 0) deallocate_vector(&v[i]);
      delete[] v;
    }
    if (packed != 0) deallocate_vector(&packed);
    if (block_cholesky != 0) deallocate_vector(&block_cholesky);
    if (block_L != 0) deallocate_vector(&block_L);
    if (block_U != 0) deallocate_vector(&block_U);
  }

  bool is_initialized() { return is_init; }
  // fine += P coarse (accumulates, transfer.h:455)
  void prolong_c2f(complex<double>* coarse_cv, complex<double>* fine_cv)
  {
    if (use_packed()) QMG_CHK(qmg_prolong_packed(&desc, qmg_host::P(packed), qmg_host::P(coarse_cv), 0, qmg_host::P(fine_cv), 0));
    else prolong_c2f(coarse_cv, fine_cv, null_vectors, num_null_vec);
  }
  // coarse += R fine (accumulates, transfer.h:487)
  void restrict_f2c(complex<double>* fine_cv, complex<double>* coarse_cv)
  {
    if (use_packed()) QMG_CHK(qmg_restrict_packed(&desc, qmg_host::P(packed), qmg_host::P(fine_cv), qmg_host::P(coarse_cv), 0));
    else restrict_f2c(fine_cv, coarse_cv, restrict_null_vectors == 0 ? null_vectors : restrict_null_vectors, num_null_vec);
  }
  // the packed copy is a snapshot of null_vectors: drop it after editing them through the public pointers
  void drop_packed() { if (packed != 0) deallocate_vector(&packed); packed_state = 0; }
  bool uses_packed_null_vectors() { return use_packed(); }
  // B200 extensions used by the K-cycle: the same sums in one pass each.
  // coarse = R fine, written outright (zero_vector + restrict_f2c)
  void restrict_f2c_overwrite(complex<double>* fine_cv, complex<double>* coarse_cv)
  {
    if (use_packed()) { QMG_CHK(qmg_restrict_packed(&desc, qmg_host::P(packed), qmg_host::P(fine_cv), qmg_host::P(coarse_cv), 1)); return; }
    complex<double>** vecs = restrict_null_vectors == 0 ? null_vectors : restrict_null_vectors;
    QMG_CHK(qmg_restrict_overwrite(&desc, reinterpret_cast<const qmg_cplx* const*>(vecs), num_null_vec, qmg_host::P(fine_cv), qmg_host::P(coarse_cv)));
  }
  // fine_out = base + P coarse (zero_vector + prolong_c2f + cxpyz); at most 8 null vectors, see can_fuse_prolong
  bool can_fuse_prolong() const { return num_null_vec <= 8; }
  void prolong_c2f_add(complex<double>* coarse_cv, complex<double>* base_cv, complex<double>* fine_out)
  {
    if (use_packed()) { QMG_CHK(qmg_prolong_packed(&desc, qmg_host::P(packed), qmg_host::P(coarse_cv), qmg_host::P(base_cv), qmg_host::P(fine_out), 1)); return; }
    QMG_CHK(qmg_prolong_add(&desc, reinterpret_cast<const qmg_cplx* const*>(null_vectors), num_null_vec, qmg_host::P(coarse_cv),
                            qmg_host::P(base_cv), qmg_host::P(fine_out)));
  }
  bool is_symmetric() { return restrict_null_vectors == 0; }
  bool has_decompositions() { return is_symmetric() ? (block_cholesky != 0) : (block_L != 0 && block_U != 0); }
  void copy_cholesky(complex<double>* save_cholesky)
  {
    if (block_cholesky == 0) std::cout << "[QMG-WARNING]: In expose_cholesky, block Cholesky has not been computed.\n";
    else copy_vector(save_cholesky, block_cholesky, coarse_lat->get_size_cm());
  }
  void copy_LU(complex<double>* save_L, complex<double>* save_U)
  {
    if (block_L == 0 || block_U == 0) std::cout << "[QMG-WARNING]: In expose_LU, block LU has not been computed.\n";
    else { copy_vector(save_L, block_L, coarse_lat->get_size_cm()); copy_vector(save_U, block_U, coarse_lat->get_size_cm()); }
  }
  QMGDoublingType get_doubling() { return doubling; }

  // kernel-level view, for CoarseOperator2D
  const qmg_transfer_desc* get_desc() const { return &desc; }
  Lattice2D* get_fine_lattice() { return fine_lat; }
  Lattice2D* get_coarse_lattice() { return coarse_lat; }
  int get_fine_sites_per_coarse() const { return fine_sites_per_coarse; }
};

#endif
