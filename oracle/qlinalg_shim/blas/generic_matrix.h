// ORACLE / TEST INFRASTRUCTURE ONLY.
// quantum-linalg's "blas/generic_matrix.h" (global dense-matrix helpers) is
// included by /root/reference/stencil/stencil_2d.h:12 but nothing on the hot
// path calls into it; the batched per-site routines are in
// generic_local_matrix.h.
#ifndef QLINALG_SHIM_GENERIC_MATRIX
#define QLINALG_SHIM_GENERIC_MATRIX
#include "generic_vector.h"
#endif
