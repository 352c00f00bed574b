// quantum-mg on B200 -- flexible GCR with a variable preconditioner: the outer solver of every K-cycle
// test and the intermediate-level solver inside the K-cycle
// (/root/reference/tests/n13_wilson_kcycle/wilson_kcycle.cpp:459, multigrid/stateful_multigrid.h:976-990).
// Same recurrences as generic_gcr.h with the new direction taken from precond(r).
#ifndef QMG_B200_GCR_VAR_PRECOND
#define QMG_B200_GCR_VAR_PRECOND

#include "generic_gcr.h"

inline inversion_info minv_vector_gcr_var_precond(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps,
                                                  matrix_op_cplx matrix_vector, void* extra_info,
                                                  precond_op_cplx precond_matrix_vector, void* precond_info,
                                                  inversion_verbose_struct* verb = 0)
{
  return qmg_host::gcr_core("VPGCR", phi, phi0, size, max_iter, eps, matrix_vector, extra_info, precond_matrix_vector, precond_info, verb);
}

inline inversion_info minv_vector_gcr_var_precond_restart(complex<double>* phi, complex<double>* phi0, int size, int max_iter, double eps, int restart_freq,
                                                          matrix_op_cplx matrix_vector, void* extra_info,
                                                          precond_op_cplx precond_matrix_vector, void* precond_info,
                                                          inversion_verbose_struct* verb = 0)
{
  return qmg_host::restarted("Restarted VPGCR", phi0, size, max_iter, eps, restart_freq, verb,
    [&](int burst, inversion_verbose_struct* quiet, qmg_host::SolveHints* h) {
      return qmg_host::gcr_core("VPGCR", phi, phi0, size, burst, eps, matrix_vector, extra_info, precond_matrix_vector, precond_info, quiet, h); });
}

#endif
