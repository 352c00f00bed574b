// quantum-mg on B200 -- result record and callback types of the Krylov solvers
// (fields fixed by /root/reference/multigrid/stateful_multigrid.h:854,997 and
//  tests/n13_wilson_kcycle/wilson_kcycle.cpp:464-466).
#ifndef QMG_B200_INVERTER_STRUCT
#define QMG_B200_INVERTER_STRUCT

#include <complex>
#include <string>
#include "../verbosity/verbosity.h"

struct inversion_info
{
  double resSq;     // |b - A x|^2 at exit, recomputed with the operator
  int iter;         // Krylov iterations performed
  bool success;     // tolerance reached within max_iter
  std::string name;
  int ops_count;    // operator applications, including the initial and the final residual
  inversion_info() : resSq(0.0), iter(0), success(false), ops_count(0) { }
};

// B200 extension (not in quantum-linalg): what a caller that knows more than the solver signature can say tells the
// device solvers, so that work whose result is known or never read is not launched.  Used by the K-cycle
// (multigrid/stateful_multigrid.h); the public minv_vector_* entry points pass none and behave as before.
namespace qmg_host {
enum
{
  SOLVE_ZERO_START = 1,          // x is zero by contract and its content is NOT read: r0 = b without applying A to zero
  SOLVE_NO_FINAL_RESIDUAL = 2,   // do not recompute |b - A x| at exit (resSq then holds the recursive value)
  SOLVE_LAST_X_ONLY = 4,         // smoother: the last permitted iteration updates x only (its residual is never read)
  SOLVE_TWO_STEP_MR = 8,         // MR with exactly two iterations from a zero start may run in its two-pass form (A applied to A r0
                                 // instead of to r1, both step lengths from one pass of dot products): same iterates up to rounding
};
// (SolveHints::residual_out, MR only: the solver's own residual vector at exit -- b - A x by the recurrence r -= alpha A r --
//  is left there for the caller, who then does not have to apply A to x to get it.  With SOLVE_LAST_X_ONLY the last step
//  still forms that vector, only its norm is not taken.)
struct SolveHints
{
  int flags;
  double bnorm2;                            // |b|^2 when the caller has just computed it, else < 0
  std::complex<double>* accumulate_into;    // MR: add the solution to this vector as well (folded into the last step)
  std::complex<double>* residual_out;       // MR: where to leave the recursive residual of the returned x (0: nobody wants it)
  bool residual_valid;                      // out: residual_out holds b - A x of the returned x
  int executed;                             // out: operator applications actually launched (ops_count keeps the reference's count)
  SolveHints(int f = 0, double b2 = -1.0) : flags(f), bnorm2(b2), accumulate_into(0), residual_out(0), residual_valid(false), executed(0) { }
};
}

#ifndef QLINALG_FCN_POINTER
#define QLINALG_FCN_POINTER
typedef void (*matrix_op_real)(double*, double*, void*);
// lhs = A rhs on DEVICE vectors; the callee overwrites lhs (stencil/stencil_2d.h:2571-2716)
typedef void (*matrix_op_cplx)(std::complex<double>*, std::complex<double>*, void*);
#endif
typedef void (*precond_op_cplx)(std::complex<double>*, std::complex<double>*, int, void*, inversion_verbose_struct*);

#endif
