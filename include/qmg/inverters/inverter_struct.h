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

#ifndef QLINALG_FCN_POINTER
#define QLINALG_FCN_POINTER
typedef void (*matrix_op_real)(double*, double*, void*);
// lhs = A rhs on DEVICE vectors; the callee overwrites lhs (stencil/stencil_2d.h:2571-2716)
typedef void (*matrix_op_cplx)(std::complex<double>*, std::complex<double>*, void*);
#endif
typedef void (*precond_op_cplx)(std::complex<double>*, std::complex<double>*, int, void*, inversion_verbose_struct*);

#endif
