// quantum-mg on B200 -- minv_vector_bicgstab is referenced only by drivers outside the hot-path scope
// (tests/n11_wilson_test); it is not provided on the device.
#ifndef QMG_B200_BICGSTAB
#define QMG_B200_BICGSTAB
#include "inverter_struct.h"
#endif
