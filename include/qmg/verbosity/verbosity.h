// quantum-mg on B200 -- solver verbosity descriptors with quantum-linalg's field names
// (fixed by /root/reference/multigrid/stateful_multigrid.h:761-776 and
//  tests/n13_wilson_kcycle/wilson_kcycle.cpp:126-131).
#ifndef QMG_B200_VERBOSITY
#define QMG_B200_VERBOSITY

#include <iostream>
#include <string>

enum inversion_verbose_level
{
  VERB_NONE = 0,
  VERB_SUMMARY = 1,
  VERB_RESTART_DETAIL = 2,
  VERB_DETAIL = 3,
};

struct inversion_verbose_struct
{
  inversion_verbose_level verbosity;
  std::string verb_prefix;
  inversion_verbose_level precond_verbosity;
  std::string precond_verb_prefix;
  inversion_verbose_struct() : verbosity(VERB_NONE), precond_verbosity(VERB_NONE) { }
  inversion_verbose_struct(inversion_verbose_level level, std::string prefix)
    : verbosity(level), verb_prefix(prefix), precond_verbosity(VERB_NONE) { }
};

namespace qmg_host {

// what a preconditioner called from inside a solver gets to print with
inline inversion_verbose_struct precond_view(const inversion_verbose_struct* outer)
{
  inversion_verbose_struct v;
  if (outer != 0)
  {
    v.verbosity = v.precond_verbosity = outer->precond_verbosity;
    v.verb_prefix = v.precond_verb_prefix = outer->precond_verb_prefix;
  }
  return v;
}
// restarted solvers silence the per-burst summaries of the solver they wrap
inline inversion_verbose_struct burst_view(const inversion_verbose_struct* outer)
{
  inversion_verbose_struct v;
  if (outer != 0) { v = *outer; if (v.verbosity == VERB_SUMMARY || v.verbosity == VERB_RESTART_DETAIL) v.verbosity = VERB_NONE; }
  return v;
}
inline void say(const inversion_verbose_struct* v, inversion_verbose_level at, const char* alg, const char* what, bool show_ok, bool ok, int iter, int ops, double relres)
{
  if (v == 0 || v->verbosity < at) return;
  std::cout << v->verb_prefix << alg << what;
  if (show_ok) std::cout << " Success " << (ok ? "Y" : "N");
  std::cout << " Iter " << iter << " Ops " << ops << " RelRes " << relres << "\n";
}

} // namespace qmg_host
#endif
