// quantum-mg on B200 -- StatefulMultigridMG: the recursive K-cycle preconditioner
// (/root/reference/multigrid/stateful_multigrid.h:43-1060).  Control flow stays on the host, as in the reference;
// every vector it touches is a device vector from the per-level pools and every operation inside is a kernel of
// libqmg_b200.so: MR pre-/post-smoothing, residual, restrict, recursive flexible-GCR (or coarsest GCR / CG) solve,
// prolong and correction.  Operator applications are counted per level and phase exactly like the reference
// (DslashTrackerMG), which is the observable the parity tests compare.
#ifndef QMG_B200_STATEFUL_MULTIGRID
#define QMG_B200_STATEFUL_MULTIGRID

#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
#include "blas/generic_vector.h"
#include "inverters/generic_minres.h"
#include "inverters/generic_cg.h"
#include "inverters/generic_gcr.h"
#include "inverters/generic_gcr_var_precond.h"
#include "interfaces/arpack/generic_arpack.h"
#include "lattice/lattice.h"
#include "transfer/transfer.h"
#include "stencil/stencil_2d.h"
#include "multigrid/multigrid.h"

enum QMGDslashType
{
  QMG_DSLASH_TYPE_NULLVEC = 0,
  QMG_DSLASH_TYPE_KRYLOV = 1,
  QMG_DSLASH_TYPE_PRESMOOTH = 2,
  QMG_DSLASH_TYPE_POSTSMOOTH = 3,
};

class StatefulMultigridMG : public MultigridMG
{
private:
  StatefulMultigridMG(StatefulMultigridMG const&);
  StatefulMultigridMG& operator=(StatefulMultigridMG const&);
  int current_level;

public:
  // how the level below a transfer is solved and smoothed (stateful_multigrid.h:62-114)
  struct LevelSolveMG
  {
    QMGStencilType fine_stencil_app;   // original, right Jacobi or Schur
    double intermediate_tol; int intermediate_iters; int intermediate_restart_freq;
    double pre_tol; int pre_iters; bool pre_cgne;
    double post_tol; int post_iters; bool post_cgne;
    LevelSolveMG()
      : fine_stencil_app(QMG_MATVEC_ORIGINAL), intermediate_tol(1e-20), intermediate_iters(10000000), intermediate_restart_freq(32),
        pre_tol(1e-20), pre_iters(1000000), pre_cgne(false), post_tol(1e-20), post_iters(1000000), post_cgne(false) { }
  };

  // operator applications per phase and Krylov iterations of one level (stateful_multigrid.h:118-200)
  class DslashTrackerMG
  {
    DslashTrackerMG(DslashTrackerMG const&);
    DslashTrackerMG& operator=(DslashTrackerMG const&);
    int counts[4];
    int iterations;
    int total;
    long executed;     // B200 extension: operator applications actually launched (the K-cycle skips applies whose result is known or unread)
  public:
    DslashTrackerMG() { reset_tracker(); }
    void add_tracker_count(QMGDslashType type, int accum) { counts[type] += accum; total += accum; }
    void add_executed_count(int accum) { executed += accum; }
    long get_executed_count() { return executed; }
    void add_iterations_count(int accum) { iterations += accum; }
    void shift_all_to_nullvec()
    {
      counts[QMG_DSLASH_TYPE_NULLVEC] += counts[QMG_DSLASH_TYPE_KRYLOV] + counts[QMG_DSLASH_TYPE_PRESMOOTH] + counts[QMG_DSLASH_TYPE_POSTSMOOTH];
      counts[QMG_DSLASH_TYPE_KRYLOV] = counts[QMG_DSLASH_TYPE_PRESMOOTH] = counts[QMG_DSLASH_TYPE_POSTSMOOTH] = 0;
      iterations = 0;
    }
    int get_tracker_count(QMGDslashType type) { return counts[type]; }
    int get_total_count() { return total; }
    int get_iterations_count() { return iterations; }
    void reset_tracker() { counts[0] = counts[1] = counts[2] = counts[3] = 0; total = 0; iterations = 0; executed = 0; }
  };

  // the coarsest solve (stateful_multigrid.h:204-241)
  struct CoarsestSolveMG
  {
    QMGStencilType coarsest_stencil_app;
    double coarsest_tol; int coarsest_iters; int coarsest_restart_freq;
    bool deflate;                 // use the eigenpairs of deflate_coarsest, once computed, to start the coarsest solve (:225)
    double normal_shift;
    CoarsestSolveMG() : coarsest_stencil_app(QMG_MATVEC_ORIGINAL), coarsest_tol(1e-20), coarsest_iters(100000000), coarsest_restart_freq(32),
                        deflate(true), normal_shift(0.0) { }
  };

protected:
  std::vector<LevelSolveMG*> level_solve_list;
  std::vector<DslashTrackerMG*> dslash_tracker_list;
  CoarsestSolveMG* coarsest_solve;
  // deflation space of the coarsest normal operator (stateful_multigrid.h:256-259)
  int coarsest_deflated;
  complex<double>* coarsest_evals;
  complex<double>** coarsest_evecs;
  bool fused_cycle;
  bool residual_handover;
  bool two_step_mr;          // QMG_MR2=0: MR(2) smoothers step by step even when residual_handover allows the two-pass form

  static void check_level_solve(LevelSolveMG* s, const char* where)
  {
    if (s->fine_stencil_app != QMG_MATVEC_ORIGINAL && s->fine_stencil_app != QMG_MATVEC_RIGHT_JACOBI && s->fine_stencil_app != QMG_MATVEC_RIGHT_SCHUR)
      std::cout << "[QMG-ERROR]: In StatefulMultigridMG:;" << where << ", LevelSolveMG::fine_stencil_app should only be original, right jacobi, or schur.\n";
  }
  void pushed(LevelSolveMG* s) { level_solve_list.push_back(s); dslash_tracker_list.push_back(new DslashTrackerMG()); }
  bool tracker_ok(int i, const char* what)
  {
    if (i >= 0 && i < num_levels) return true;
    std::cout << "[QMG-ERROR]: Out of range: Cannot " << what << " at level " << i << ".\n";
    return false;
  }

public:
  StatefulMultigridMG(Lattice2D* in_lat, Stencil2D* in_stencil, CoarsestSolveMG* in_coarsest_solve)
    : MultigridMG(in_lat, in_stencil), current_level(0), coarsest_solve(in_coarsest_solve), coarsest_deflated(0), coarsest_evals(0), coarsest_evecs(0)
  {
    dslash_tracker_list.push_back(new DslashTrackerMG());
    const char* e = getenv("QMG_FUSED_CYCLE");
    fused_cycle = !(e != 0 && e[0] == '0');
    const char* e2 = getenv("QMG_RESIDUAL_HANDOVER");
    residual_handover = !(e2 != 0 && e2[0] == '0');
    const char* e3 = getenv("QMG_MR2");
    two_step_mr = !(e3 != 0 && e3[0] == '0');
    qmg_host::overwriting_precond() = mg_preconditioner;    // writes every element of its output: the solvers need not zero it
  }
  ~StatefulMultigridMG()
  {
    for (size_t i = 0; i < dslash_tracker_list.size(); i++) delete dslash_tracker_list[i];
    clear_deflation();
  }

  // ---- coarsest-level deflation (stateful_multigrid.h:613-711): num_low smallest and num_high largest eigenpairs of the
  // coarsest NORMAL operator, from the Lanczos restatement of arpack_dcn (interfaces/arpack/generic_arpack.h; ncv = 3 nev,
  // tol 1e-5 as the reference asks of ARPACK).  mg_preconditioner then starts every coarsest solve from the projection of
  // its right-hand side onto that space.
  void clear_deflation()
  {
    if (coarsest_evecs != 0)
    {
      for (int i = 0; i < coarsest_deflated; i++) if (coarsest_evecs[i] != 0) deallocate_vector(&coarsest_evecs[i]);
      delete[] coarsest_evecs;
    }
    if (coarsest_evals != 0) delete[] coarsest_evals;
    coarsest_evecs = 0; coarsest_evals = 0; coarsest_deflated = 0;
  }
  void deflate_coarsest(int num_low, int num_high, bool print_evals = false)
  {
    if (!coarsest_solve->deflate)
      std::cout << "[QMG-WARNING]: Coarsest level is not set to deflate. Skipping computing eigenvectors.\n";
    const QMGStencilType app = coarsest_solve->coarsest_stencil_app;
    if (app != QMG_MATVEC_M_MDAGGER && app != QMG_MATVEC_MDAGGER_M && app != QMG_MATVEC_RBJ_M_MDAGGER && app != QMG_MATVEC_RBJ_MDAGGER_M)
    {
      std::cout << "[QMG-ERROR]: Cannot deflate coarsest operator unless it's a normal op solve.\n";
      return;
    }
    if (coarsest_deflated != 0 || coarsest_evals != 0 || coarsest_evecs != 0)
    {
      std::cout << "[QMG-WARNING]: Coarsest operator space already deflated.\n";
      return;
    }
    if (num_low + num_high == 0) return;
    Stencil2D* coarsest = get_stencil(get_num_levels() - 1);
    const int n = coarsest->get_lattice()->get_size_cv();
    const int total = num_low + num_high;
    coarsest_evals = new complex<double>[total];
    coarsest_evecs = new complex<double>*[total];
    for (int i = 0; i < total; i++) coarsest_evecs[i] = allocate_vector<complex<double> >(n);
    coarsest_deflated = total;
    bool ok = true;
    if (num_low > 0)
    {
      arpack_dcn eig(n, 100000, 1e-5, Stencil2D::get_apply_function(app), (void*)coarsest, num_low, 3 * num_low);
      ok = eig.prepare_eigensystem(arpack_dcn::ARPACK_SMALLEST_REAL, num_low, 3 * num_low) &&
           eig.get_eigensystem(coarsest_evals, coarsest_evecs, arpack_dcn::ARPACK_SMALLEST_REAL);
    }
    if (ok && num_high > 0)
    {
      arpack_dcn eig(n, 100000, 1e-5, Stencil2D::get_apply_function(app), (void*)coarsest, num_high, 3 * num_high);
      ok = eig.prepare_eigensystem(arpack_dcn::ARPACK_LARGEST_REAL, num_high, 3 * num_high) &&
           eig.get_eigensystem(coarsest_evals + num_low, coarsest_evecs + num_low, arpack_dcn::ARPACK_LARGEST_REAL);
    }
    if (!ok) { clear_deflation(); return; }
    for (int i = 0; i < total; i++) normalize(coarsest_evecs[i], n);
    if (print_evals)
      for (int i = 0; i < total; i++) std::cout << "[QMG-COARSEST-EVALS]: " << i << " " << real(coarsest_evals[i]) << "\n";
  }
  unsigned int get_coarsest_deflated() { return coarsest_deflated; }
  complex<double>* get_coarsest_evals() { return coarsest_evals; }
  complex<double>** get_coarsest_evecs() { return coarsest_evecs; }

  // ---- the level cursor the preconditioner callback reads (the object is stateful and not re-entrant)
  void set_multigrid_level(int level)
  {
    if (level >= 0 && level < get_num_levels()) current_level = level;
    else std::cout << "[QMG-ERROR]: Out of range: StatefulMultigridMG->current_level " << level << " is outside of [0,max_level-1].\n";
  }
  void go_finer()
  {
    if (current_level > 0) current_level--;
    else std::cout << "[QMG-ERROR]: Out of range: Cannot go finer than the top level in StatefulMultigridMG.\n";
  }
  void go_coarser()
  {
    if (current_level < get_num_levels() - 2) current_level++;
    else std::cout << "[QMG-ERROR]: Out of range: Cannot go coarser than the second-coarsest level in StatefulMultigridMG.\n";
  }
  int get_multigrid_level() { return current_level; }
  LevelSolveMG* get_level_solve(int i)
  {
    if (i >= 0 && i < get_num_levels() - 1 && level_solve_list[i] != 0) return level_solve_list[i];
    std::cout << "[QMG-ERROR]: Out of range: LevelSolveMG level " << i << " does not exist in StatefulMultigridMG object.\n";
    return 0;
  }
  LevelSolveMG* get_level_solve() { return get_level_solve(current_level); }
  CoarsestSolveMG* get_coarsest_solve() { return coarsest_solve; }

  // ---- level management: the MultigridMG overloads, each with and without a LevelSolveMG (stateful_multigrid.h:374-496)
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, bool build_stencil = false, bool is_chiral = false,
                  QMGMultigridPrecondStencil build_stencil_from = QMG_MULTIGRID_PRECOND_ORIGINAL,
                  CoarseOperator2D::QMGCoarseBuildStencil build_extra = CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, complex<double>** nvecs = 0)
  { MultigridMG::push_level(new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, build_extra, nvecs); pushed(0); }
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, bool build_stencil, bool is_chiral, QMGMultigridPrecondStencil build_stencil_from, complex<double>** nvecs)
  { MultigridMG::push_level(new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, nvecs); pushed(0); }
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, complex<double>** nvecs)
  { MultigridMG::push_level(new_lat, new_transfer, nvecs); pushed(0); }
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, LevelSolveMG* in_solve, bool build_stencil = false, bool is_chiral = false,
                  QMGMultigridPrecondStencil build_stencil_from = QMG_MULTIGRID_PRECOND_ORIGINAL,
                  CoarseOperator2D::QMGCoarseBuildStencil build_extra = CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, complex<double>** nvecs = 0)
  {
    MultigridMG::push_level(new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, build_extra, nvecs);
    check_level_solve(in_solve, "push_level");
    pushed(in_solve);
  }
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, LevelSolveMG* in_solve, bool build_stencil, bool is_chiral,
                  QMGMultigridPrecondStencil build_stencil_from, complex<double>** nvecs)
  {
    MultigridMG::push_level(new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, nvecs);
    check_level_solve(in_solve, "push_level");
    pushed(in_solve);
  }
  void push_level(Lattice2D* new_lat, TransferMG* new_transfer, LevelSolveMG* in_solve, complex<double>** nvecs)
  {
    MultigridMG::push_level(new_lat, new_transfer, nvecs);
    check_level_solve(in_solve, "push_level");
    pushed(in_solve);
  }
  void pop_level()
  {
    if (num_levels == 1) { MultigridMG::pop_level(); return; }
    level_solve_list.pop_back();
    delete dslash_tracker_list.back();
    dslash_tracker_list.pop_back();
    MultigridMG::pop_level();
  }
  void update_level(int level, Lattice2D* new_lat, TransferMG* new_transfer, LevelSolveMG* in_solve, bool build_stencil = false, bool is_chiral = false,
                    QMGMultigridPrecondStencil build_stencil_from = QMG_MULTIGRID_PRECOND_ORIGINAL,
                    CoarseOperator2D::QMGCoarseBuildStencil build_extra = CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, complex<double>** nvecs = 0)
  {
    if (in_solve->fine_stencil_app != QMG_MATVEC_ORIGINAL && in_solve->fine_stencil_app != QMG_MATVEC_RIGHT_JACOBI && in_solve->fine_stencil_app != QMG_MATVEC_RIGHT_SCHUR)
    { check_level_solve(in_solve, "update_level"); return; }
    MultigridMG::update_level(level, new_lat, new_transfer, build_stencil, is_chiral, build_stencil_from, build_extra, nvecs);
    if (level >= 1 && level < num_levels) level_solve_list[level - 1] = in_solve;
  }
  void update_level(int level, Lattice2D* new_lat, TransferMG* new_transfer, LevelSolveMG* in_solve, bool build_stencil, bool is_chiral,
                    QMGMultigridPrecondStencil build_stencil_from, complex<double>** nvecs)
  { update_level(level, new_lat, new_transfer, in_solve, build_stencil, is_chiral, build_stencil_from, CoarseOperator2D::QMG_COARSE_BUILD_ORIGINAL, nvecs); }

  // ---- counters (stateful_multigrid.h:500-609)
  void add_tracker_count(QMGDslashType type, int accum, int i) { if (tracker_ok(i, "update tracker")) dslash_tracker_list[i]->add_tracker_count(type, accum); }
  void add_iterations_count(int accum, int i) { if (tracker_ok(i, "update tracker")) dslash_tracker_list[i]->add_iterations_count(accum); }
  // B200 extension: applications launched, beside the reference's counts (which include the skipped A.0 and unread residual applies)
  void add_executed_count(int accum, int i) { if (tracker_ok(i, "update tracker")) dslash_tracker_list[i]->add_executed_count(accum); }
  long get_executed_count(int i) { return tracker_ok(i, "query tracker") ? dslash_tracker_list[i]->get_executed_count() : -1; }
  // 1 (default): zero-start / unread-residual shortcuts and one-pass residual, restrict and prolong-correct steps;
  // 0: the reference's sequence of separate sweeps, launch for launch (what round 1 ran).  Results are bit-identical.
  void set_fused_cycle(bool on) { fused_cycle = on; }
  bool get_fused_cycle() { return fused_cycle; }
  // With the fused cycle: 1 (default) the pre-smoother hands its own residual r = rhs - A z1 (the MR recurrence r -= alpha A r,
  // which its last step then completes without taking the norm) to the restriction, so that the reference's explicit
  // "Atmp = A z1; r1 = rhs - Atmp" (:863-866) is not launched: one operator apply less per K-cycle application and level
  // (7 -> 6 at the BASELINE settings).  The two residuals are the same vector up to rounding (two MR steps from a zero start),
  // so this is the one shortcut that is NOT bit-identical: 0 keeps the explicit residual and with it the bits of the unfused cycle.
  // The same switch lets the post-smoother answer a flexible solver's request for A lhs (A lhs = rhs - r2', r2' its recurrence
  // residual; inverters/generic_gcr.h PrecondAzRequest): the Krylov apply after each K-cycle application goes too (6 -> 5), and
  // lets MR(2) smoothers run in their two-pass form (SOLVE_TWO_STEP_MR, inverters/generic_minres.h): 20 instead of 30 vector
  // passes of BLAS-1 per K-cycle application and one host wait per smoother instead of two.
  void set_residual_handover(bool on) { residual_handover = on; }
  bool get_residual_handover() { return residual_handover; }
  void shift_all_to_nullvec(int i) { if (tracker_ok(i, "shift to null vectors")) dslash_tracker_list[i]->shift_all_to_nullvec(); }
  int get_tracker_count(QMGDslashType type, int i) { return tracker_ok(i, "query tracker") ? dslash_tracker_list[i]->get_tracker_count(type) : -1; }
  int get_total_count(int i) { return tracker_ok(i, "query tracker") ? dslash_tracker_list[i]->get_total_count() : -1; }
  int get_iterations_count(int i) { return tracker_ok(i, "query tracker") ? dslash_tracker_list[i]->get_iterations_count() : -1; }
  std::vector<double> query_average_iterations()
  {
    std::vector<double> avg(num_levels);
    avg[0] = dslash_tracker_list[0]->get_iterations_count();
    for (int i = 1; i < num_levels; i++)
      avg[i] = ((double)dslash_tracker_list[i]->get_iterations_count()) / ((double)dslash_tracker_list[i - 1]->get_iterations_count());
    return avg;
  }
  void reset_tracker(int i = -1)
  {
    if (i == -1) for (int j = 0; j < num_levels; j++) dslash_tracker_list[j]->reset_tracker();
    else if (tracker_ok(i, "reset tracker")) dslash_tracker_list[i]->reset_tracker();
  }

protected:
  // A + sigma for the shifted coarsest normal-equation solve (stateful_multigrid.h:716-729)
  struct ShiftedFunctionStruct { matrix_op_cplx function; void* extra_data; complex<double> extra_shift; int length; };
  static void shift_function(complex<double>* out, complex<double>* in, void* data)
  {
    ShiftedFunctionStruct* s = (ShiftedFunctionStruct*)data;
    s->function(out, in, s->extra_data);
    caxpy(s->extra_shift, in, out, s->length);
  }

  // z = MR_smooth(A_type, b) from a zero start; with cgne the smoother runs on A A^dag and z = A^dag z'.  Returns the operator
  // applications the reference counts; `executed` receives those launched.  hints == 0: z must be zero on entry.
  static int smooth(Stencil2D* op, ArrayStorageMG<complex<double> >* pool, QMGStencilType type, bool cgne, int iters, double tol,
                    complex<double>* z, complex<double>* b, int n_solve, long n_full, qmg_host::SolveHints* hints, int& executed)
  {
    if (cgne && (type == QMG_MATVEC_ORIGINAL || type == QMG_MATVEC_RIGHT_JACOBI))
    {
      const bool orig = (type == QMG_MATVEC_ORIGINAL);
      complex<double>* zp = pool->check_out();
      qmg_host::SolveHints h2(hints != 0 ? (hints->flags & ~qmg_host::SOLVE_LAST_X_ONLY) : 0);
      if (hints == 0) zero_vector(zp, n_full);
      inversion_info inv = qmg_host::minres_core(zp, b, n_solve, iters, tol, 0.85,
                                                 Stencil2D::get_apply_function(orig ? QMG_MATVEC_M_MDAGGER : QMG_MATVEC_RBJ_M_MDAGGER), (void*)op, 0, hints != 0 ? &h2 : 0);
      if (hints != 0) zero_vector(z, n_full);      // apply_M accumulates
      op->apply_M(z, zp, orig ? QMG_MATVEC_DAGGER : QMG_MATVEC_RBJ_DAGGER);
      if (hints != 0 && hints->accumulate_into != 0) cxpy(z, hints->accumulate_into, n_solve);
      pool->check_in(zp);
      executed += (hints != 0) ? 2 * h2.executed + 1 : 2 * inv.ops_count + 1;
      return 2 * inv.ops_count + 1;
    }
    inversion_info inv = qmg_host::minres_core(z, b, n_solve, iters, tol, 0.85, Stencil2D::get_apply_function(type), (void*)op, 0, hints);
    executed += (hints != 0) ? hints->executed : inv.ops_count;
    return inv.ops_count;
  }

public:
  // One K-cycle application at the current level: lhs ~ A^-1 rhs.  Signature of a quantum-linalg preconditioner;
  // extra_data is the StatefulMultigridMG (stateful_multigrid.h:734).
  //
  // The steps and their order are the reference's (:840-1051).  With fused_cycle (default) the work whose result is
  // known or never read is not launched -- the smoothers and coarse solves start from a zero vector, so their initial
  // residual is the right-hand side and A.0 is not applied (:860, :915-990); the true-residual apply at the end of each
  // of them feeds only inversion_info::resSq, which this function never reads (:854,:861,:996 use ops_count and iter) --
  // and steps the reference spells as several sweeps run as one: r = b - A z (:863-866, :1023-1029), zero + restrict
  // (:876-878), zero + prolong + z1 + z2 (:1005-1019), lhs += z3 (:1050).  Every number that IS computed is formed by
  // the same kernel arithmetic in the same order: lhs is bit-identical with fused_cycle on and off -- with ONE exception that
  // has its own switch (set_residual_handover, on by default): the residual after the pre-smoother is the smoother's own
  // (recurrence) residual instead of a fresh rhs - A z1, equal to it up to rounding.  The trackers keep the reference's
  // operator counts; add_executed_count records what was launched.
  static void mg_preconditioner(complex<double>* lhs, complex<double>* rhs, int size, void* extra_data, inversion_verbose_struct* verb)
  {
    (void)size;
    StatefulMultigridMG* mg = (StatefulMultigridMG*)extra_data;
    const int level = mg->get_multigrid_level();
    const int nlev = mg->get_num_levels();
    const bool fuse = mg->fused_cycle;
    using qmg_host::SolveHints;
    using qmg_host::SOLVE_ZERO_START; using qmg_host::SOLVE_NO_FINAL_RESIDUAL; using qmg_host::SOLVE_LAST_X_ONLY;

    // a flexible solver above may have asked for A lhs along with lhs (inverters/generic_gcr.h); the request is taken off the
    // board here so that the solves one level down, which post their own, do not see it
    qmg_host::PrecondAzRequest* az_req = qmg_host::precond_az_request();
    qmg_host::precond_az_request() = 0;

    LevelSolveMG* ls = mg->get_level_solve();
    if (ls == 0) { std::cout << "[QMG-MG-SOLVE-ERROR]: Level solve for level " << level << " does not exist.\n"; return; }
    const QMGStencilType ftype = ls->fine_stencil_app;
    const long nf = mg->get_lattice(level)->get_size_cv();
    const int nf_solve = (int)(ftype == QMG_MATVEC_RIGHT_SCHUR ? nf / 2 : nf);
    if (nlev == 1) { copy_vector(lhs, rhs, nf_solve); return; }

    Stencil2D* fine = mg->get_stencil(level);
    Stencil2D* coarse = mg->get_stencil(level + 1);
    TransferMG* transfer = mg->get_transfer(level);
    ArrayStorageMG<complex<double> >* fpool = mg->get_storage(level);
    ArrayStorageMG<complex<double> >* cpool = mg->get_storage(level + 1);
    const long nc = mg->get_lattice(level + 1)->get_size_cv();
    matrix_op_cplx fine_op = Stencil2D::get_apply_function(ftype);

    // what the solvers below this level print with
    inversion_verbose_struct verb2(VERB_SUMMARY, std::string(" "));
    if (verb == 0 || verb->verbosity == VERB_NONE) { verb2.verbosity = VERB_NONE; verb2.precond_verbosity = VERB_NONE; }
    else verb2.precond_verbosity = VERB_SUMMARY;
    verb2.verb_prefix = std::string(2 * (level + 1), ' ') + "[QMG-MG-SOLVE-INFO]: Level " + std::to_string(level + 1) + " ";

    // the solve one level down: an intermediate level (recursive K-cycle) or the coarsest
    const bool coarsest = (level == nlev - 2);
    QMGStencilType ctype; int c_iters; double c_tol; int c_restart;
    if (!coarsest)
    {
      LevelSolveMG* next = mg->get_level_solve(level + 1);
      ctype = next->fine_stencil_app; c_iters = next->intermediate_iters; c_tol = next->intermediate_tol; c_restart = next->intermediate_restart_freq;
    }
    else
    {
      CoarsestSolveMG* cs = mg->get_coarsest_solve();
      ctype = cs->coarsest_stencil_app; c_iters = cs->coarsest_iters; c_tol = cs->coarsest_tol; c_restart = cs->coarsest_restart_freq;
    }
    matrix_op_cplx coarse_op = Stencil2D::get_apply_function(ctype);
    const int nc_solve = (int)(ctype == QMG_MATVEC_RIGHT_SCHUR ? nc / 2 : nc);
    // the smoother shortcuts: zero start, no unread residual; its last step may skip r only when the tolerance cannot stop it earlier anyway
    const int smooth_flags = SOLVE_ZERO_START | SOLVE_NO_FINAL_RESIDUAL | SOLVE_LAST_X_ONLY | ((mg->residual_handover && mg->two_step_mr) ? qmg_host::SOLVE_TWO_STEP_MR : 0);

    // 1. pre-smooth: z1 ~ A^-1 rhs, r1 = rhs - A z1
    complex<double>* z1 = fpool->check_out();
    complex<double>* r1 = fpool->check_out();
    if (!fuse) zero_vector(z1, nf);
    else if (nf_solve != nf) zero_vector(z1 + nf_solve, nf - nf_solve);     // Schur: the odd half is read as zero further down
    // Schur: the residual lives on the even sites, but the restriction below reads the whole vector.  The reference leaves
    // the odd half of r1 as whatever the pool vector last held (stateful_multigrid.h:843 "gets initialized in the next code
    // block" -- only its first fine_size_solve elements are); here it is zero, so that the K-cycle is a function of its input.
    if (nf_solve != nf) zero_vector(r1 + nf_solve, nf - nf_solve);
    if (ls->pre_iters > 0)
    {
      SolveHints hints(smooth_flags);
      if (fuse && mg->residual_handover) hints.residual_out = r1;      // the smoother's own residual, if it can give one
      int executed = 0;
      const int ops = smooth(fine, fpool, ftype, ls->pre_cgne, ls->pre_iters, ls->pre_tol, z1, rhs, nf_solve, nf, fuse ? &hints : 0, executed);
      mg->add_tracker_count(QMG_DSLASH_TYPE_PRESMOOTH, ops, level);
      const bool handed = fuse && hints.residual_valid;
      if (!handed && !(fuse && fine->apply_residual(r1, rhs, z1, ftype)))
      {
        complex<double>* Az = fpool->check_out();
        fine_op(Az, z1, (void*)fine);
        caxpbyz(1.0, rhs, -1.0, Az, r1, nf_solve);
        fpool->check_in(Az);
      }
      mg->add_tracker_count(QMG_DSLASH_TYPE_PRESMOOTH, 1, level);      // the reference's count: it applies the operator here
      mg->add_executed_count(executed + (handed ? 0 : 1), level);
    }
    else
    {
      copy_vector(r1, rhs, nf_solve);
      copy_vector(z1, rhs, nf_solve);
    }

    // 2. restrict the residual and bring it to the form the coarse system is solved in
    complex<double>* r_c = cpool->check_out();
    if (fuse) transfer->restrict_f2c_overwrite(r1, r_c);
    else { zero_vector(r_c, nc); transfer->restrict_f2c(r1, r_c); }
    fpool->check_in(r1);
    const double rsq_c = norm2sq(r_c, nc);
    const double rnorm = sqrt(rsq_c);
    // prepare_M is a copy for every operator flavour except Schur and the M^dag M normal equations (stencil_2d.h:2455-2489)
    const bool prep_is_copy = !(ctype == QMG_MATVEC_RIGHT_SCHUR || ctype == QMG_MATVEC_MDAGGER_M || ctype == QMG_MATVEC_RBJ_MDAGGER_M);
    complex<double>* r_c_prep = r_c;
    double rsq_prep = rsq_c;
    if (!(fuse && prep_is_copy))
    {
      r_c_prep = cpool->check_out();
      zero_vector(r_c_prep, nc);
      coarse->prepare_M(r_c_prep, r_c, ctype);
      rsq_prep = norm2sq(r_c_prep, nc);
    }
    const double rnorm_prep = sqrt(rsq_prep);
    const double tol = c_tol * rnorm / rnorm_prep;

    // 3. coarse solve from a zero start
    complex<double>* e_c = cpool->check_out();
    inversion_info inv;
    int coarse_executed = -1;
    const bool normal = coarsest && (ctype == QMG_MATVEC_M_MDAGGER || ctype == QMG_MATVEC_MDAGGER_M || ctype == QMG_MATVEC_RBJ_M_MDAGGER || ctype == QMG_MATVEC_RBJ_MDAGGER_M);
    // zero-start hints for the GCR family; |b|^2 is the norm just taken when it ran over the same elements
    SolveHints chints(SOLVE_ZERO_START | SOLVE_NO_FINAL_RESIDUAL, nc_solve == nc ? rsq_prep : -1.0);
    const bool hinted = fuse && !normal;
    if (!hinted) zero_vector(e_c, nc);
    else if (nc_solve != nc) zero_vector(e_c + nc_solve, nc - nc_solve);
    if (coarsest)
    {
      ShiftedFunctionStruct shifted;
      matrix_op_cplx op = coarse_op; void* op_data = (void*)coarse;
      if (normal && mg->get_coarsest_solve()->normal_shift != 0.0)
      {
        shifted.function = coarse_op; shifted.extra_data = (void*)coarse;
        shifted.extra_shift = mg->get_coarsest_solve()->normal_shift; shifted.length = nc_solve;
        op = shift_function; op_data = (void*)&shifted;
      }
      // start from the projection of the right-hand side onto the deflation space (stateful_multigrid.h:895-907)
      if (normal && mg->get_coarsest_solve()->deflate && mg->get_coarsest_deflated() > 0)
      {
        const int num_evecs = mg->get_coarsest_deflated();
        complex<double>* evals = mg->get_coarsest_evals();
        complex<double>** evecs = mg->get_coarsest_evecs();
        for (int i = 0; i < num_evecs; i++)
        {
          const complex<double> bra_n_ket_b = dot(evecs[i], r_c_prep, nc_solve);
          caxpy(bra_n_ket_b / evals[i], evecs[i], e_c, nc_solve);
        }
      }
      if (!normal)
      {
        if (hinted) { inv = qmg_host::gcr_solve(e_c, r_c_prep, nc_solve, c_iters, tol, c_restart, op, op_data, 0, 0, &verb2, &chints); coarse_executed = chints.executed; }
        else inv = (c_restart == -1) ? minv_vector_gcr(e_c, r_c_prep, nc_solve, c_iters, tol, op, op_data, &verb2)
                                     : minv_vector_gcr_restart(e_c, r_c_prep, nc_solve, c_iters, tol, c_restart, op, op_data, &verb2);
      }
      else
        inv = (c_restart == -1) ? minv_vector_cg(e_c, r_c_prep, nc_solve, c_iters, tol, op, op_data, &verb2)
                                : minv_vector_cg_restart(e_c, r_c_prep, nc_solve, c_iters, tol, c_restart, op, op_data, &verb2);
    }
    else
    {
      mg->go_coarser();
      if (hinted) { inv = qmg_host::gcr_solve(e_c, r_c_prep, nc_solve, c_iters, tol, c_restart, coarse_op, (void*)coarse, mg_preconditioner, (void*)mg, &verb2, &chints); coarse_executed = chints.executed; }
      else inv = (c_restart == -1) ? minv_vector_gcr_var_precond(e_c, r_c_prep, nc_solve, c_iters, tol, coarse_op, (void*)coarse, mg_preconditioner, (void*)mg, &verb2)
                                   : minv_vector_gcr_var_precond_restart(e_c, r_c_prep, nc_solve, c_iters, tol, c_restart, coarse_op, (void*)coarse, mg_preconditioner, (void*)mg, &verb2);
      mg->go_finer();
    }
    mg->add_tracker_count(QMG_DSLASH_TYPE_KRYLOV, inv.ops_count, level + 1);
    mg->add_executed_count(coarse_executed >= 0 ? coarse_executed : inv.ops_count, level + 1);
    mg->add_iterations_count(inv.iter, level + 1);
    if (r_c_prep != r_c) cpool->check_in(r_c_prep);

    // 4. undo the preparation, prolong, correct: lhs = z1 + P e
    // reconstruct_M is a copy for the flavours whose solution IS the unknown (stencil_2d.h:2492-2527)
    const bool recon_is_copy = (ctype == QMG_MATVEC_ORIGINAL || ctype == QMG_MATVEC_DAGGER || ctype == QMG_MATVEC_MDAGGER_M || ctype == QMG_MATVEC_RBJ_DAGGER);
    complex<double>* e_full = e_c;
    if (!(fuse && recon_is_copy))
    {
      e_full = cpool->check_out();
      zero_vector(e_full, nc);
      coarse->reconstruct_M(e_full, e_c, r_c, ctype);
    }
    cpool->check_in(r_c);
    if (e_full != e_c) cpool->check_in(e_c);
    if (fuse && ftype != QMG_MATVEC_RIGHT_SCHUR && transfer->can_fuse_prolong())
      transfer->prolong_c2f_add(e_full, z1, lhs);
    else
    {
      complex<double>* z2 = fpool->check_out();
      zero_vector(z2, nf);
      transfer->prolong_c2f(e_full, z2);
      if (ctype == QMG_MATVEC_RIGHT_SCHUR) zero_vector(z2 + nf / 2, nf / 2);   // stateful_multigrid.h:1012
      cxpyz(z1, z2, lhs, nf_solve);
      fpool->check_in(z2);
    }
    cpool->check_in(e_full);
    fpool->check_in(z1);

    // 5. post-smooth on the new residual
    if (ls->post_iters > 0)
    {
      complex<double>* r2 = fpool->check_out();
      if (!(fuse && fine->apply_residual(r2, rhs, lhs, ftype)))
      {
        complex<double>* Ax = fpool->check_out();
        fine_op(Ax, lhs, (void*)fine);
        caxpbyz(1.0, rhs, -1.0, Ax, r2, nf_solve);
        fpool->check_in(Ax);
      }
      complex<double>* z3 = fpool->check_out();
      SolveHints hints(smooth_flags);
      hints.accumulate_into = lhs;        // lhs += z3 rides on the smoother's last step
      // A lhs for the solver that asked: r2 = rhs - A (z1 + P e) is exact here, the smoother's recurrence turns it (in place)
      // into r2' = r2 - A z3, so A lhs = rhs - r2' -- two more vector passes instead of an operator apply
      const bool serve_az = fuse && mg->residual_handover && az_req != 0 && az_req->op == fine_op && az_req->op_data == (void*)fine && nf_solve == nf;
      if (serve_az) hints.residual_out = r2;
      int executed = 0;
      if (!fuse) zero_vector(z3, nf);
      const int ops = smooth(fine, fpool, ftype, ls->post_cgne, ls->post_iters, ls->post_tol, z3, r2, nf_solve, nf, fuse ? &hints : 0, executed);
      mg->add_tracker_count(QMG_DSLASH_TYPE_POSTSMOOTH, ops, level);
      mg->add_executed_count(executed + 1, level);
      if (!fuse) cxpy(z3, lhs, nf_solve);
      if (serve_az && hints.residual_valid) { caxpbyz(1.0, rhs, -1.0, r2, az_req->out, nf_solve); az_req->valid = true; }
      fpool->check_in(r2);
      fpool->check_in(z3);
    }
  }
};

#endif
