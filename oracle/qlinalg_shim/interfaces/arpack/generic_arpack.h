// ORACLE / TEST INFRASTRUCTURE ONLY.
// quantum-linalg "interfaces/arpack/generic_arpack.h".  ARPACK/gfortran are not
// in this image and spectra are out of scope; the
// class is declared so drivers that mention it in dead branches
// (/root/reference/tests/n13_wilson_kcycle/wilson_kcycle.cpp:290) still compile.
#ifndef QLINALG_SHIM_ARPACK
#define QLINALG_SHIM_ARPACK
#include <complex>
#include <iostream>
#include "../../inverters/inverter_struct.h"
class arpack_dcn
{
public:
  enum arpack_spectrum_piece { ARPACK_NONE, ARPACK_LARGEST_MAGNITUDE, ARPACK_SMALLEST_MAGNITUDE, ARPACK_LARGEST_REAL, ARPACK_SMALLEST_REAL, ARPACK_LARGEST_IMAGINARY, ARPACK_SMALLEST_IMAGINARY };
  arpack_dcn(int, int, double, matrix_op_cplx, void*) { complain(); }
  arpack_dcn(int, int, double, matrix_op_cplx, void*, int, int) { complain(); }
  bool prepare_eigensystem(arpack_spectrum_piece, int, int) { return false; }
  bool get_eigensystem(std::complex<double>*, std::complex<double>**, arpack_spectrum_piece) { return false; }
  bool get_eigensystem(std::complex<double>*, arpack_spectrum_piece) { return false; }
  bool get_entire_eigensystem(std::complex<double>*, arpack_spectrum_piece) { return false; }
  bool get_entire_eigensystem(std::complex<double>*, std::complex<double>**, arpack_spectrum_piece) { return false; }
private:
  static void complain() { std::cout << "[QLINALG-SHIM]: ARPACK is not available in the oracle shim.\n"; }
};
#endif
