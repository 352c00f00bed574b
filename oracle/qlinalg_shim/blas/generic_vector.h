// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// Clean-room stand-in for quantum-linalg's "blas/generic_vector.h".
// quantum-linalg (github.com/weinbe2/quantum-linalg, no pinned version) is
// an un-vendored dependency of the reference and is NOT under /root/reference.
// Every function here is restated from the way the reference calls it; the
// call site that fixes the meaning is cited next to each one
// (paths relative to /root/reference).
//
// Parity status: BLAS semantics pinned by use (and by the reference's own
// identity tests n05/n06/n08); RNG draw order and solver internals UNPINNED.
//
// Plain single-threaded loops on host memory, like the reference.

#ifndef QLINALG_SHIM_GENERIC_VECTOR
#define QLINALG_SHIM_GENERIC_VECTOR

#include <cmath>
#include <complex>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

// The reference leans on names leaking out of quantum-linalg
// (stencil/stencil_2d.h:17 `complex`, :447 `string`, u1/u1_utils.h:38 `fstream`).
using namespace std;

// ---------------------------------------------------------------- alloc ----
// stencil/stencil_2d.h:220 (allocate), :299 (deallocate sets pointer to 0).
template <typename T> inline T* allocate_vector(int n) { return new T[n > 0 ? n : 1]; }
template <typename T> inline void deallocate_vector(T** v) { delete[] *v; *v = 0; }

// ------------------------------------------------------------ fill/copy ----
// stencil/stencil_2d.h:341
template <typename T> inline void zero_vector(T* v, int n) { for (int i = 0; i < n; i++) v[i] = T(0); }
// operators/wilson.h:100 zero_vector_blas(ptr, stride, count)
template <typename T> inline void zero_vector_blas(T* v, int stride, int n) { for (int i = 0; i < n; i++) v[i * stride] = T(0); }
// stencil/stencil_2d.h:986 copy_vector(dst, src, n)
template <typename T> inline void copy_vector(T* dst, const T* src, int n) { for (int i = 0; i < n; i++) dst[i] = src[i]; }
// operators/wilson.h:110 copy_vector_blas(dst, src, stride, n) (same stride both sides)
template <typename T> inline void copy_vector_blas(T* dst, const T* src, int stride, int n) { for (int i = 0; i < n; i++) dst[i * stride] = src[i * stride]; }
// transfer/transfer.h:560 copy_vector_blas(dst, dst_stride, src, src_stride, n)
template <typename T> inline void copy_vector_blas(T* dst, int dstride, const T* src, int sstride, int n) { for (int i = 0; i < n; i++) dst[i * dstride] = src[i * sstride]; }
// operators/gaugedlaplace.h:45 constant_vector(v, value, n)
template <typename T, typename U> inline void constant_vector(T* v, U val, int n) { for (int i = 0; i < n; i++) v[i] = T(val); }
// operators/wilson.h:167 constant_vector_blas(v, stride, value, n)
template <typename T, typename U> inline void constant_vector_blas(T* v, int stride, U val, int n) { for (int i = 0; i < n; i++) v[i * stride] = T(val); }

// ---------------------------------------------------------- elementwise ----
// operators/wilson.h:194
template <typename T> inline void conj_vector(complex<T>* v, int n) { for (int i = 0; i < n; i++) v[i] = conj(v[i]); }
// u1/u1_utils.h:407 is called on a real (double) phase field: conj of a real is the identity.
inline void conj_vector(double*, int) { }
// transfer/transfer.h:736 (abs of each element, stays complex)
template <typename T> inline void abs_vector(complex<T>* v, int n) { for (int i = 0; i < n; i++) v[i] = complex<T>(abs(v[i]), 0); }
inline void abs_vector(double* v, int n) { for (int i = 0; i < n; i++) v[i] = fabs(v[i]); }
// u1/u1_utils.h:371 arg_vector(complex* in, double* out, n) and in-place complex flavour
template <typename T> inline void arg_vector(complex<T>* in, T* out, int n) { for (int i = 0; i < n; i++) out[i] = arg(in[i]); }
template <typename T> inline void arg_vector(complex<T>* v, int n) { for (int i = 0; i < n; i++) v[i] = complex<T>(arg(v[i]), 0); }
// u1/u1_utils.h:194 polar(v, n): real part holds a phase -> e^{i phase}
template <typename T> inline void polar(complex<T>* v, int n) { for (int i = 0; i < n; i++) v[i] = std::polar(T(1), real(v[i])); }
// tests/n13_wilson_kcycle/wilson_kcycle.cpp:212 polar_vector(phases, links, n)
template <typename T> inline void polar_vector(T* phases, complex<T>* out, int n) { for (int i = 0; i < n; i++) out[i] = std::polar(T(1), phases[i]); }
// transfer/transfer.h:588 cinvx(x, n): x -> 1/x
template <typename T> inline void cinvx(T* v, int n) { for (int i = 0; i < n; i++) v[i] = T(1) / v[i]; }
// u1/u1_utils.h:255 cxty(x, y, n): y *= x
template <typename T> inline void cxty(const T* x, T* y, int n) { for (int i = 0; i < n; i++) y[i] *= x[i]; }
// transfer/transfer.h:583, operators/staggered.h:61
template <typename T> inline void arb_local_function_vector(T* v, void (*fcn)(int, T&, void*), void* extra, int n) { for (int i = 0; i < n; i++) fcn(i, v[i], extra); }

// ------------------------------------------------------------ axpy family --
// operators/wilson.h:79 cax_blas(a, x, stride, n): x *= a (strided); staggered.h:142 cax(a, x, n)
template <typename T, typename U> inline void cax(U a, T* x, int n) { for (int i = 0; i < n; i++) x[i] *= a; }
template <typename T, typename U> inline void cax_blas(U a, T* x, int stride, int n) { for (int i = 0; i < n; i++) x[i * stride] *= a; }
// operators/staggered.h:148 caxy(a, x, y, n): y = a x
template <typename T, typename U> inline void caxy(U a, const T* x, T* y, int n) { for (int i = 0; i < n; i++) y[i] = a * x[i]; }
// operators/wilson.h:181 caxy_blas(a, x, xstride, y, ystride, n)
template <typename T, typename U> inline void caxy_blas(U a, const T* x, int xs, T* y, int ys, int n) { for (int i = 0; i < n; i++) y[i * ys] = a * x[i * xs]; }
// stencil/stencil_2d.h:676 caxpy(a, x, y, n): y += a x
template <typename T, typename U> inline void caxpy(U a, const T* x, T* y, int n) { for (int i = 0; i < n; i++) y[i] += a * x[i]; }
// operators/dwf.h:190 caxpy_blas(a, x, xstride, y, ystride, n)
template <typename T, typename U> inline void caxpy_blas(U a, const T* x, int xs, T* y, int ys, int n) { for (int i = 0; i < n; i++) y[i * ys] += a * x[i * xs]; }
// stencil/stencil_2d.h:903 caxpy_stride(a, x, y, size, offset, stride)
template <typename T, typename U> inline void caxpy_stride(U a, const T* x, T* y, int size, int offset, int stride) { for (int i = offset; i < size; i += stride) y[i] += a * x[i]; }
// stencil/stencil_2d.h:1685 cxpy(x, y, n): y += x
template <typename T> inline void cxpy(const T* x, T* y, int n) { for (int i = 0; i < n; i++) y[i] += x[i]; }
// multigrid/stateful_multigrid.h:1019 cxpyz(x, y, z, n): z = x + y
template <typename T> inline void cxpyz(const T* x, const T* y, T* z, int n) { for (int i = 0; i < n; i++) z[i] = x[i] + y[i]; }
// stencil/stencil_2d.h:1924 cxpay(x, a, y, n): y = x + a y
template <typename T, typename U> inline void cxpay(const T* x, U a, T* y, int n) { for (int i = 0; i < n; i++) y[i] = x[i] + a * y[i]; }
// operators/staggered.h:201 caxpby(a, x, b, y, n): y = a x + b y
template <typename T, typename U, typename W> inline void caxpby(U a, const T* x, W b, T* y, int n) { for (int i = 0; i < n; i++) y[i] = a * x[i] + b * y[i]; }
// stencil/stencil_2d.h:1907 caxpbyz(a, x, b, y, z, n): z = a x + b y
template <typename T, typename U, typename W> inline void caxpbyz(U a, const T* x, W b, const T* y, T* z, int n) { for (int i = 0; i < n; i++) z[i] = a * x[i] + b * y[i]; }
// tests/n07_free_laplace_mg/free_laplace_mg.cpp:178 caxpbypz(a, x, b, y, z, n): z += a x + b y
template <typename T, typename U, typename W> inline void caxpbypz(U a, const T* x, W b, const T* y, T* z, int n) { for (int i = 0; i < n; i++) z[i] += a * x[i] + b * y[i]; }

// --------------------------------------------------------------- patterns --
// stencil/stencil_2d.h:1526 capx_pattern(pattern, len, v, nrepeat): v[r*len+k] += pattern[k]
template <typename T, typename P> inline void capx_pattern(const P* pattern, int len, T* v, int nrepeat)
{
  for (int r = 0; r < nrepeat; r++) for (int k = 0; k < len; k++) v[r * len + k] += pattern[k];
}
// operators/wilson.h:132 caxy_shuffle_pattern(scale, shuffle, n, in, out, nsites):
// out[s*n+i] = scale[i] * in[s*n+shuffle[i]]
template <typename T, typename P> inline void caxy_shuffle_pattern(const P* scale, const int* shuffle, int n, const T* in, T* out, int nsites)
{
  for (int s = 0; s < nsites; s++) for (int i = 0; i < n; i++) out[s * n + i] = scale[i] * in[s * n + shuffle[i]];
}

// ------------------------------------------------------------- reductions --
template <typename T> struct RealReducer { typedef T type; };
template <typename T> struct RealReducer<complex<T> > { typedef T type; };
template <typename T> struct Reducer { typedef T type; };
template <typename T> struct ComplexBase
{
  static inline T conj(T x) { return x; }
  static inline T real(T x) { return x; }
};
template <typename T> struct ComplexBase<complex<T> >
{
  static inline complex<T> conj(complex<T> x) { return std::conj(x); }
  static inline T real(complex<T> x) { return std::real(x); }
};

// multigrid/stateful_multigrid.h:880
template <typename T> inline typename RealReducer<T>::type norm2sq(const T* v, int n)
{
  typename RealReducer<T>::type s = 0;
  for (int i = 0; i < n; i++) s += ComplexBase<T>::real(ComplexBase<T>::conj(v[i]) * v[i]);
  return s;
}
// tests/n13_wilson_kcycle/wilson_kcycle.cpp:471
template <typename T> inline typename RealReducer<T>::type diffnorm2sq(const T* a, const T* b, int n)
{
  typename RealReducer<T>::type s = 0;
  for (int i = 0; i < n; i++) { T d = a[i] - b[i]; s += ComplexBase<T>::real(ComplexBase<T>::conj(d) * d); }
  return s;
}
// stencil/stencil_2d.h:411
template <typename T> inline typename RealReducer<T>::type norminf(const T* v, int n)
{
  typename RealReducer<T>::type m = 0;
  for (int i = 0; i < n; i++) { typename RealReducer<T>::type a = abs(v[i]); if (a > m) m = a; }
  return m;
}
// multigrid/stateful_multigrid.h:904: <a|b>, conjugate on the first argument.
template <typename T> inline T dot(const T* a, const T* b, int n)
{
  T s = 0;
  for (int i = 0; i < n; i++) s += ComplexBase<T>::conj(a[i]) * b[i];
  return s;
}
template <typename T> inline typename RealReducer<T>::type re_dot(const T* a, const T* b, int n)
{
  return ComplexBase<T>::real(dot(a, b, n));
}
template <typename T> inline T sum_vector(const T* v, int n) { T s = 0; for (int i = 0; i < n; i++) s += v[i]; return s; }
// tests/n13_wilson_kcycle/wilson_kcycle.cpp:383
template <typename T> inline void normalize(T* v, int n)
{
  typename RealReducer<T>::type nrm = sqrt(norm2sq(v, n));
  for (int i = 0; i < n; i++) v[i] /= nrm;
}
// tests/n13_wilson_kcycle/wilson_kcycle.cpp:348 orthogonal(v, against, n): v -= <against|v>/<against|against> against
template <typename T> inline void orthogonal(T* v, const T* against, int n)
{
  T c = dot(against, v, n) / norm2sq(against, n);
  for (int i = 0; i < n; i++) v[i] -= c * against[i];
}

// -------------------------------------------------------------------- RNG --
// Draw order / scaling UNPINNED: sources are always generated once and fed to
// both oracle and GPU as arrays.  tests/n13_wilson_kcycle/wilson_kcycle.cpp:341
template <typename T> inline void gaussian(complex<T>* v, int n, std::mt19937& gen, T dev = T(1))
{
  std::normal_distribution<T> dist(0.0, dev);
  for (int i = 0; i < n; i++) { T re = dist(gen); T im = dist(gen); v[i] = complex<T>(re, im); }
}
inline void gaussian(double* v, int n, std::mt19937& gen, double dev = 1.0)
{
  std::normal_distribution<double> dist(0.0, dev);
  for (int i = 0; i < n; i++) v[i] = dist(gen);
}
template <typename T> inline void gaussian_real(complex<T>* v, int n, std::mt19937& gen, T dev = T(1))
{
  std::normal_distribution<T> dist(0.0, dev);
  for (int i = 0; i < n; i++) v[i] = complex<T>(dist(gen), 0);
}
// u1/u1_utils.h:193 random_uniform(v, n, gen, lo, hi)
template <typename T> inline void random_uniform(complex<T>* v, int n, std::mt19937& gen, T lo, T hi)
{
  std::uniform_real_distribution<T> dist(lo, hi);
  for (int i = 0; i < n; i++) v[i] = complex<T>(dist(gen), 0);
}
inline void random_uniform(double* v, int n, std::mt19937& gen, double lo, double hi)
{
  std::uniform_real_distribution<double> dist(lo, hi);
  for (int i = 0; i < n; i++) v[i] = dist(gen);
}

// The batched small-matrix routines live in generic_local_matrix.h in
// quantum-linalg; the reference includes all three headers together
// (stencil/stencil_2d.h:10-12).
#include "generic_local_matrix.h"

#endif
