// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "inverters/inverter_struct.h".
// Fields fixed by use: /root/reference/multigrid/stateful_multigrid.h:854 (ops_count),
// :997 (iter), tests/n13_wilson_kcycle/wilson_kcycle.cpp:464-466 (success, iter, resSq).
#ifndef QLINALG_SHIM_INVERTER_STRUCT
#define QLINALG_SHIM_INVERTER_STRUCT

#include <complex>
#include <string>
#include "../verbosity/verbosity.h"

struct inversion_info
{
  double resSq;     // |b - A x|^2 at exit (true residual, recomputed)
  int iter;         // Krylov iterations performed
  bool success;     // reached tolerance before max_iter
  std::string name;
  int ops_count;    // operator applications, including the initial and final residual
  inversion_info() : resSq(0.0), iter(0), success(false), name(""), ops_count(0) { }
};

#ifndef QLINALG_FCN_POINTER
#define QLINALG_FCN_POINTER
typedef void (*matrix_op_real)(double*, double*, void*);
typedef void (*matrix_op_cplx)(std::complex<double>*, std::complex<double>*, void*);
#endif

typedef void (*precond_op_cplx)(std::complex<double>*, std::complex<double>*, int, void*, inversion_verbose_struct*);

#endif
