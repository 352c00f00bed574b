// ORACLE / TEST INFRASTRUCTURE ONLY.
// Clean-room stand-in for quantum-linalg "verbosity/verbosity.h".
// Field names are fixed by /root/reference/multigrid/stateful_multigrid.h:761-776
// and tests/n13_wilson_kcycle/wilson_kcycle.cpp:126-131.
#ifndef QLINALG_SHIM_VERBOSITY
#define QLINALG_SHIM_VERBOSITY

#include <iostream>
#include <string>

enum inversion_verbose_level
{
  VERB_NONE = 0,     // print nothing
  VERB_SUMMARY = 1,  // one line at the end of a solve
  VERB_RESTART_DETAIL = 2, // plus one line per restart
  VERB_DETAIL = 3,   // plus one line per iteration
};

struct inversion_verbose_struct
{
  inversion_verbose_level verbosity;
  std::string verb_prefix;
  inversion_verbose_level precond_verbosity;
  std::string precond_verb_prefix;

  inversion_verbose_struct()
    : verbosity(VERB_NONE), verb_prefix(""), precond_verbosity(VERB_NONE), precond_verb_prefix("") { }
  inversion_verbose_struct(inversion_verbose_level level, std::string prefix)
    : verbosity(level), verb_prefix(prefix), precond_verbosity(VERB_NONE), precond_verb_prefix("") { }
};

// Hand the preconditioner its own verbosity (prefix/level swapped in).
inline void shuffle_verbosity_precond(inversion_verbose_struct* out, inversion_verbose_struct* in)
{
  if (in == 0) { out->verbosity = VERB_NONE; out->precond_verbosity = VERB_NONE; return; }
  out->verbosity = in->precond_verbosity;
  out->verb_prefix = in->precond_verb_prefix;
  out->precond_verbosity = in->precond_verbosity;
  out->precond_verb_prefix = in->precond_verb_prefix;
}

inline void print_verbosity_resid(inversion_verbose_struct* verb, const std::string& alg, int iter, int ops, double relres)
{
  if (verb != 0 && verb->verbosity >= VERB_DETAIL)
    std::cout << verb->verb_prefix << alg << " Iter " << iter << " Ops " << ops << " RelRes " << relres << "\n";
}

inline void print_verbosity_restart(inversion_verbose_struct* verb, const std::string& alg, int iter, int ops, double relres)
{
  if (verb != 0 && verb->verbosity >= VERB_RESTART_DETAIL)
    std::cout << verb->verb_prefix << alg << " Restart Iter " << iter << " Ops " << ops << " RelRes " << relres << "\n";
}

inline void print_verbosity_summary(inversion_verbose_struct* verb, const std::string& alg, bool success, int iter, int ops, double relres)
{
  if (verb != 0 && verb->verbosity >= VERB_SUMMARY)
    std::cout << verb->verb_prefix << alg << " Success " << (success ? "Y" : "N") << " Iter " << iter << " Ops " << ops << " RelRes " << relres << "\n";
}

#endif
