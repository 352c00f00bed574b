// quantum-linalg keeps the dense (non-batched) matrix helpers here; the hot path uses none of them.
#ifndef QMG_B200_GENERIC_MATRIX
#define QMG_B200_GENERIC_MATRIX
#include "generic_local_matrix.h"
#endif
