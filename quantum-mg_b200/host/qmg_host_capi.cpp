// libqmg_host.so: the flat driver API of qmg_capi_body.h compiled against the B200 host classes
// (include/qmg/: Lattice2D, Stencil2D and operators, TransferMG, CoarseOperator2D, StatefulMultigridMG and the
// device solvers).  Every array the classes see lives in HBM; the host arrays of the driver API are staged in and out.
// The same driver text compiled against the unmodified reference headers is the oracle (oracle/ref_capi.cpp).
#include <cstring>
#include <complex>

#include "blas/generic_vector.h"
#include "inverters/generic_cg.h"
#include "inverters/generic_gcr.h"
#include "inverters/generic_gcr_var_precond.h"
#include "inverters/generic_minres.h"
#include "inverters/generic_bicgstab_l.h"
#include "inverters/generic_richardson.h"
#include "inverters/generic_bicgstab.h"
#include "inverters/generic_tfqmr.h"

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

#define QMG_B200_HOST 1
#define CAPI(name) qmgh_##name
static inline std::complex<double>* capi_alloc(long n) { return allocate_vector<std::complex<double> >(n); }
static inline void capi_free(std::complex<double>* p) { deallocate_vector(&p); }
static inline void capi_put(std::complex<double>* dst, const std::complex<double>* src, long n) { qmg_host::upload(dst, src, n); }
static inline void capi_get(std::complex<double>* dst, const std::complex<double>* src, long n) { qmg_host::download(dst, src, n); }
static inline void capi_put_real(double* dst, const double* src, long n) { qmg_host::upload_real(dst, src, n); }
static inline void capi_get_real(double* dst, const double* src, long n) { qmg_host::download_real(dst, src, n); }
static inline void capi_barrier() { QMG_CHK(qmg_sync()); }

#include "qmg_capi_body.h"

extern "C" const char* qmgh_backend(void) { return "B200 host classes (include/qmg) over libqmg_b200.so, sm_100a kernels"; }
