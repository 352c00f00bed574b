// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// Builds oracle/_ref/libqmg_ref.so: the reference's own UNMODIFIED headers,
// compiled from where they lie under /root/reference, against the clean-room
// quantum-linalg shim (oracle/qlinalg_shim), behind the flat C driver API of
// quantum-mg_b200/host/qmg_capi_body.h with the symbol prefix ref_.
// No reference source is copied into this repository.
//
// Recipe: oracle/Makefile (g++ -O2 -std=c++11, the reference's own flags,
// /root/reference/tests/n11_wilson_test/Makefile:15).

#include <cstring>
#include <complex>

// quantum-linalg stand-ins
#include "blas/generic_vector.h"
#include "inverters/generic_cg.h"
#include "inverters/generic_gcr.h"
#include "inverters/generic_gcr_var_precond.h"
#include "inverters/generic_minres.h"
#include "inverters/generic_bicgstab_l.h"
#include "inverters/generic_richardson.h"
#include "inverters/generic_bicgstab.h"
#include "inverters/generic_tfqmr.h"

// the reference (resolved through -I/root/reference)
#include "lattice/lattice.h"
#include "cshift/cshift_2d.h"
#include "stencil/stencil_2d.h"
#include "operators/wilson.h"
#include "operators/staggered.h"
#include "operators/gaugedlaplace.h"
#include "operators/dwf.h"
#include "operators/coarse.h"
#include "transfer/transfer.h"
#include "multigrid/stateful_multigrid.h"
#include "u1/u1_utils.h"
#include "reductions/reductions.h"

#define CAPI(name) ref_##name
static inline std::complex<double>* capi_alloc(long n) { return allocate_vector<std::complex<double> >((int)n); }
static inline void capi_free(std::complex<double>* p) { deallocate_vector(&p); }
static inline void capi_put(std::complex<double>* dst, const std::complex<double>* src, long n) { std::memcpy(dst, src, sizeof(std::complex<double>) * n); }
static inline void capi_get(std::complex<double>* dst, const std::complex<double>* src, long n) { std::memcpy(dst, src, sizeof(std::complex<double>) * n); }
static inline void capi_put_real(double* dst, const double* src, long n) { std::memcpy(dst, src, sizeof(double) * n); }
static inline void capi_get_real(double* dst, const double* src, long n) { std::memcpy(dst, src, sizeof(double) * n); }
static inline void capi_barrier() { }

#include "../quantum-mg_b200/host/qmg_capi_body.h"

extern "C" const char* ref_backend(void) { return "reference headers (/root/reference) + qlinalg_shim, CPU, single thread"; }
