// quantum-mg on B200 -- MultigridMG: the level container (/root/reference/multigrid/multigrid.h:54-600):
// lattices, transfers, stencils, one device-vector pool per level and (optionally) the raw null vectors.
// Level 0's stencil, all lattices and all transfers belong to the caller; coarse stencils built here belong to the object.
#ifndef QMG_B200_MULTIGRID
#define QMG_B200_MULTIGRID

#include <iostream>
#include <vector>
#include "blas/generic_vector.h"
#include "lattice/lattice.h"
#include "stencil/stencil_2d.h"
#include "transfer/transfer.h"
#include "storage/array_storage.h"
#include "operators/coarse.h"

class MultigridMG
{
protected:
  MultigridMG(MultigridMG const&);
  MultigridMG& operator=(MultigridMG const&);

  struct Level
  {
    Lattice2D* lattice;
    Stencil2D* stencil;
    bool stencil_managed;
    ArrayStorageMG<complex<double> >* storage;
    TransferMG* transfer_up;            // transfer between level-1 and this level (0 on level 0)
    complex<double>** raw_null_vectors; // copy of the vectors that built transfer_up, or 0; lives on level-1's lattice
    int n_raw;
  };
  std::vector<Level> levels;
  int num_levels;   // kept as a member: StatefulMultigridMG reads it directly (stateful_multigrid.h:459)

  static bool in_range(bool ok, const char* what, int i)
  {
    if (!ok) std::cout << "[QMG-ERROR]: Out of range: " << what << " level " << i << " does not exist in MultigridMG object.\n";
    return ok;
  }
  void release(Level& l, long fine_size)
  {
    (void)fine_size;
    if (l.storage != 0) { delete l.storage; l.storage = 0; }
    if (l.stencil_managed && l.stencil != 0) { delete l.stencil; l.stencil = 0; }
    if (l.raw_null_vectors != 0)
    {
      for (int j = 0; j < l.n_raw; j++) if (l.raw_null_vectors[j] != 0) deallocate_vector(&l.raw_null_vectors[j]);
      delete[] l.raw_null_vectors; l.raw_null_vectors = 0;
    }
  }

public:
  enum QMGMultigridPrecondStencil
  {
    QMG_MULTIGRID_PRECOND_ORIGINAL = 0,
    QMG_MULTIGRID_PRECOND_RIGHT_BLOCK_JACOBI = 1,
  };

private:
  Level make_level(int below, Lattice2D* new_lat, TransferMG* new_transfer, bool build_stencil, bool is_chiral,
                   QMGMultigridPrecondStencil build_stencil_from, CoarseOperator2D::QMGCoarseBuildStencil build_extra, complex<double>** nvecs)
  {
    Level l;
    l.lattice = new_lat;
    l.transfer_up = new_transfer;
    l.storage = new ArrayStorageMG<complex<double> >(new_lat->get_size_cv(), 6);   // multigrid.h:110,269
    l.stencil = 0; l.stencil_managed = false;
    if (build_stencil)
    {
      l.stencil = new CoarseOperator2D(new_lat, levels[below].stencil, levels[below].lattice, new_transfer, is_chiral,
                                       build_stencil_from != QMG_MULTIGRID_PRECOND_ORIGINAL, build_extra);
      l.stencil_managed = true;
    }
    l.raw_null_vectors = 0; l.n_raw = new_lat->get_nc();
    if (nvecs != 0)
    {
      const long n = levels[below].lattice->get_size_cv();
      l.raw_null_vectors = new complex<double>*[l.n_raw];
      for (int j = 0; j < l.n_raw; j++)
      {
        l.raw_null_vectors[j] = 0;
        if (nvecs[j] != 0) { l.raw_null_vectors[j] = allocate_vector<complex<double> >(n); copy_vector(l.raw_null_vectors[j], nvecs[j], n); }
      }
    }
    return l;
  }

public:
  MultigridMG(Lattice2D* in_lat, Stencil2D* in_stencil) : num_levels(1)
  {
    Level l = { in_lat, in_stencil, false, new ArrayStorageMG<complex<double> >(in_lat->get_size_cv(), 6), 0, 0, 0 };
    levels.push_back(l);
  }
  virtual ~MultigridMG() { for (size_t i = 0; i < levels.size(); i++) release(levels[i], 0); }

  inline int get_num_levels() { return num_levels; }
  inline Lattice2D* get_lattice(int i) { return in_range(i >= 0 && i < num_levels, "Lattice", i) ? levels[i].lattice : 0; }
  inline TransferMG* get_transfer(int i) { return in_range(i >= 0 && i < num_levels - 1, "Transfer object", i) ? levels[i + 1].transfer_up : 0; }
  inline Stencil2D* get_stencil(int i) { return in_range(i >= 0 && i < num_levels, "Stencil object", i) ? levels[i].stencil : 0; }
  inline ArrayStorageMG<complex<double> >* get_storage(int i) { return in_range(i >= 0 && i < num_levels, "Array storage object", i) ? levels[i].storage : 0; }

  void get_global_null_vectors(int i, complex<double>** out_vectors)
  {
    if (!in_range(i >= 0 && i < num_levels - 1, "Null vectors", i)) return;
    Level& l = levels[i + 1];
    if (out_vectors == 0 || l.raw_null_vectors == 0) return;
    for (int j = 0; j < l.n_raw; j++)
      if (out_vectors[j] != 0 && l.raw_null_vectors[j] != 0) copy_vector(out_vectors[j], l.raw_null_vectors[j], levels[i].lattice->get_size_cv());
  }

  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, bool build_stencil = false, bool is_chiral = false,
                  QMGMultigridPrecondStencil build_stencil_from = QMG_MULTIGRID_PRECOND_ORIGINAL,
                  CoarseOperator2D::QMGCoarseBuildStencil build_extra = CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, complex<double>** nvecs = 0)
  {
    levels.push_back(make_level(num_levels - 1, new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, build_extra, nvecs));
    num_levels++;
  }
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, bool build_stencil, bool is_chiral,
                  QMGMultigridPrecondStencil build_stencil_from, complex<double>** nvecs)
  { push_level(new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, nvecs); }
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, complex<double>** nvecs)
  { push_level(new_lat, new_transfer, false, false, QMG_MULTIGRID_PRECOND_ORIGINAL, CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, nvecs); }

  void pop_level()
  {
    if (num_levels == 1) { std::cout << "[QMG-ERROR]: In MultigridMG::pop_level, cannot pop when there is only one level.\n"; return; }
    release(levels.back(), 0);
    levels.pop_back();
    num_levels--;
  }

  // replace level `level` (>= 1) and the transfer above it (multigrid.h:375-450)
  void update_level(int level, Lattice2D* new_lat, TransferMG* new_transfer, bool build_stencil = false, bool is_chiral = false,
                    QMGMultigridPrecondStencil build_stencil_from = QMG_MULTIGRID_PRECOND_ORIGINAL,
                    CoarseOperator2D::QMGCoarseBuildStencil build_extra = CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, complex<double>** nvecs = 0)
  {
    if (level < 1 || level >= num_levels)
    {
      std::cout << "[QMG-ERROR]: In MultigridMG::update_level, cannot update level " << level << " as it does not exist yet anyway.\n";
      return;
    }
    release(levels[level], 0);
    levels[level] = make_level(level - 1, new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, build_extra, nvecs);
  }
  void update_level(int level, Lattice2D* new_lat, TransferMG* new_transfer, bool build_stencil, bool is_chiral,
                    QMGMultigridPrecondStencil build_stencil_from, complex<double>** nvecs)
  { update_level(level, new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, nvecs); }

  // lhs += A_i rhs; a level without an explicit stencil is emulated as R A_{i-1} P (multigrid.h:465-505)
  void apply_stencil(complex<double>* lhs, complex<double>* rhs, int i, QMGStencilType app_type = QMG_MATVEC_ORIGINAL)
  {
    if (!(i >= 0 && i < num_levels)) { std::cout << "[QMG-ERROR]: Out of range: Cannot apply stencil at level " << i << "\n"; return; }
    if (levels[i].stencil != 0) { levels[i].stencil->apply_M(lhs, rhs, app_type); return; }
    if (app_type != QMG_MATVEC_ORIGINAL)
    {
      std::cout << "[QMG-ERROR]: In MultigridMG::apply_stencil, the emulated operator must be QMG_MATVEC_ORIGINAL.\n";
      return;
    }
    const long nf = levels[i - 1].lattice->get_size_cv();
    complex<double>* pro = levels[i - 1].storage->check_out();
    complex<double>* Apro = levels[i - 1].storage->check_out();
    zero_vector(pro, nf); zero_vector(Apro, nf);
    levels[i].transfer_up->prolong_c2f(rhs, pro);
    apply_stencil(Apro, pro, i - 1);
    levels[i].transfer_up->restrict_f2c(Apro, lhs);
    levels[i - 1].storage->check_in(pro);
    levels[i - 1].storage->check_in(Apro);
  }
  void prolong_c2f(complex<double>* coarse_cv, complex<double>* fine_cv, int i)
  {
    if (i >= 0 && i < num_levels - 1) levels[i + 1].transfer_up->prolong_c2f(coarse_cv, fine_cv);
    else std::cout << "[QMG-ERROR]: Out of range: Cannot apply prolong at level " << i << "\n";
  }
  void restrict_f2c(complex<double>* fine_cv, complex<double>* coarse_cv, int i)
  {
    if (i >= 0 && i < num_levels - 1) levels[i + 1].transfer_up->restrict_f2c(fine_cv, coarse_cv);
    else std::cout << "[QMG-ERROR]: Out of range: Cannot apply restrict at level " << i << "\n";
  }
  complex<double>* check_out(int i)
  {
    if (i >= 0 && i < num_levels) return levels[i].storage->check_out();
    std::cout << "[QMG-ERROR]: Out of range: Cannot check out vector at level " << i << ".\n";
    return 0;
  }
  void check_in(complex<double>* vec, int i)
  {
    if (i >= 0 && i < num_levels) levels[i].storage->check_in(vec);
    else std::cout << "[QMG-ERROR]: Out of range: Cannot check in vector at level " << i << ".\n";
  }
  int get_storage_number_allocated(int i)
  {
    if (i >= 0 && i < num_levels) return levels[i].storage->get_number_allocated();
    std::cout << "[QMG-ERROR]: Out of range: Cannot query number of allocated arrays at level " << i << ".\n";
    return -1;
  }
  int get_storage_number_checked(int i)
  {
    if (i >= 0 && i < num_levels) return levels[i].storage->get_number_checked();
    std::cout << "[QMG-ERROR]: Out of range: Cannot query number of checked out arrays at level " << i << ".\n";
    return -1;
  }
};

#endif
