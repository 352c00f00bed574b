// quantum-mg on B200 -- device-resident BLAS-1 with the call signatures the reference uses from
// its quantum-linalg dependency ("blas/generic_vector.h"; SURVEY.md 8c lists the call sites).
//
// Every pointer handed to these functions is DEVICE memory obtained from allocate_vector
// (cudaMalloc, or managed memory under qmg_set_alloc_mode(1) / QMG_MANAGED=1 so that unmodified
// reference drivers that index vectors from host code keep working).  Each call is one launch of a
// hand-written sm_100a kernel in libqmg_b200.so (include/qmg_b200.h); reductions return their
// value to the host.  There is no host implementation: without a GPU every call fails loudly.
#ifndef QMG_B200_GENERIC_VECTOR
#define QMG_B200_GENERIC_VECTOR

#include <cmath>
#include <complex>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../qmg_b200.h"

// the reference relies on these names being visible unqualified (stencil/stencil_2d.h:17, :447)
using namespace std;

namespace qmg_host {

typedef std::complex<double> cplx;

inline void check(int rc, const char* what)
{
  if (rc != 0)
  {
    std::cout << "[QMG-ERROR]: " << what << " failed: " << qmg_last_error() << "\n";
    std::cout.flush();
    std::abort();   // no CPU fallback: a failed kernel call is fatal
  }
}
inline qmg_cplx* P(cplx* p) { return reinterpret_cast<qmg_cplx*>(p); }
inline const qmg_cplx* P(const cplx* p) { return reinterpret_cast<const qmg_cplx*>(p); }

// host <-> device staging for the few routines whose definition is a host loop (RNG fills, I/O)
inline void upload(cplx* dev, const cplx* host, long n) { check(qmg_memcpy_h2d(dev, host, sizeof(cplx) * (size_t)n), "upload"); }
inline void download(cplx* host, const cplx* dev, long n) { check(qmg_memcpy_d2h(host, dev, sizeof(cplx) * (size_t)n), "download"); }

} // namespace qmg_host

#define QMG_CHK(call) qmg_host::check((call), #call)

// ---------------------------------------------------------------- alloc ----
// stencil/stencil_2d.h:220, :299 (deallocate sets the pointer to 0)
template <typename T> inline T* allocate_vector(long n)
{
  void* p = 0;
  QMG_CHK(qmg_malloc(&p, sizeof(T) * (size_t)(n > 0 ? n : 1)));
  return reinterpret_cast<T*>(p);
}
template <typename T> inline void deallocate_vector(T** v) { if (*v != 0) { QMG_CHK(qmg_free((void*)*v)); } *v = 0; }

// ------------------------------------------------------------ fill/copy ----
typedef std::complex<double> qmg_cd;

// templates, because drivers also spell the element type out (tests/n07_free_laplace_mg/free_laplace_mg.cpp:480:
// zero_vector<complex<double>>(e, n)) and call them on real phase fields (tests/n13_wilson_kcycle/wilson_kcycle.cpp:203)
template <typename T> inline void zero_vector(T* v, long n) { QMG_CHK(qmg_zero_bytes(v, sizeof(T) * (size_t)(n > 0 ? n : 0))); }
template <typename T> inline void copy_vector(T* dst, const T* src, long n) { if (n > 0 && dst != src) QMG_CHK(qmg_memcpy_d2d(dst, src, sizeof(T) * (size_t)n)); }
template <typename T> inline void copy_vector(T* dst, T* src, long n) { copy_vector<T>(dst, static_cast<const T*>(src), n); }
template <typename U> inline void constant_vector(qmg_cd* v, U val, long n) { qmg_cd a(val); QMG_CHK(qmg_constant(qmg_host::P(v), a.real(), a.imag(), n)); }
inline void zero_vector_blas(qmg_cd* v, int stride, long n) { QMG_CHK(qmg_zero_strided(qmg_host::P(v), stride, n)); }
template <typename U> inline void constant_vector_blas(qmg_cd* v, int stride, U val, long n) { qmg_cd a(val); QMG_CHK(qmg_constant_strided(qmg_host::P(v), stride, a.real(), a.imag(), n)); }
// operators/wilson.h:110 (same stride on both sides) and transfer/transfer.h:560 (two strides)
inline void copy_vector_blas(qmg_cd* dst, const qmg_cd* src, int stride, long n) { QMG_CHK(qmg_caxy_strided(1.0, 0.0, qmg_host::P(src), stride, qmg_host::P(dst), stride, n, 0)); }
inline void copy_vector_blas(qmg_cd* dst, int dstride, const qmg_cd* src, int sstride, long n) { QMG_CHK(qmg_caxy_strided(1.0, 0.0, qmg_host::P(src), sstride, qmg_host::P(dst), dstride, n, 0)); }

// ---------------------------------------------------------- elementwise ----
inline void conj_vector(qmg_cd* v, long n) { QMG_CHK(qmg_conj(qmg_host::P(v), n)); }
inline void abs_vector(qmg_cd* v, long n) { QMG_CHK(qmg_elementwise(0, qmg_host::P(v), n)); }
inline void arg_vector(qmg_cd* v, long n) { QMG_CHK(qmg_elementwise(4, qmg_host::P(v), n)); }
inline void polar(qmg_cd* v, long n) { QMG_CHK(qmg_polar(qmg_host::P(v), n)); }
inline void cinvx(qmg_cd* v, long n) { QMG_CHK(qmg_cinvx(qmg_host::P(v), n)); }
inline void cxty(const qmg_cd* x, qmg_cd* y, long n) { QMG_CHK(qmg_cxty(qmg_host::P(x), qmg_host::P(y), n)); }

// transfer/transfer.h:583, operators/staggered.h:61: an arbitrary HOST callback per element.  Setup-path only:
// the vector is staged through host memory (download, loop, upload).
inline void arb_local_function_vector(qmg_cd* v, void (*fcn)(int, qmg_cd&, void*), void* extra, long n)
{
  std::vector<qmg_cd> h((size_t)n);
  qmg_host::download(h.data(), v, n);
  for (long i = 0; i < n; i++) fcn((int)i, h[i], extra);
  qmg_host::upload(v, h.data(), n);
}

// ------------------------------------------------------------ axpy family --
template <typename U> inline void cax(U a_, qmg_cd* x, long n) { qmg_cd a(a_); QMG_CHK(qmg_cax(a.real(), a.imag(), qmg_host::P(x), n)); }
template <typename U> inline void cax_blas(U a_, qmg_cd* x, int stride, long n) { qmg_cd a(a_); QMG_CHK(qmg_cax_strided(a.real(), a.imag(), qmg_host::P(x), stride, n)); }
template <typename U> inline void caxy(U a_, const qmg_cd* x, qmg_cd* y, long n) { qmg_cd a(a_); QMG_CHK(qmg_caxy(a.real(), a.imag(), qmg_host::P(x), qmg_host::P(y), n)); }
template <typename U> inline void caxy_blas(U a_, const qmg_cd* x, int xs, qmg_cd* y, int ys, long n) { qmg_cd a(a_); QMG_CHK(qmg_caxy_strided(a.real(), a.imag(), qmg_host::P(x), xs, qmg_host::P(y), ys, n, 0)); }
template <typename U> inline void caxpy(U a_, const qmg_cd* x, qmg_cd* y, long n) { qmg_cd a(a_); QMG_CHK(qmg_caxpy(a.real(), a.imag(), qmg_host::P(x), qmg_host::P(y), n)); }
template <typename U> inline void caxpy_blas(U a_, const qmg_cd* x, int xs, qmg_cd* y, int ys, long n) { qmg_cd a(a_); QMG_CHK(qmg_caxy_strided(a.real(), a.imag(), qmg_host::P(x), xs, qmg_host::P(y), ys, n, 1)); }
// stencil/stencil_2d.h:903: elements offset, offset+stride, ... below size
template <typename U> inline void caxpy_stride(U a_, const qmg_cd* x, qmg_cd* y, long size, int offset, int stride)
{
  qmg_cd a(a_);
  const long count = (size - offset + stride - 1) / stride;
  QMG_CHK(qmg_caxy_strided(a.real(), a.imag(), qmg_host::P(x + offset), stride, qmg_host::P(y + offset), stride, count, 1));
}
inline void cxpy(const qmg_cd* x, qmg_cd* y, long n) { QMG_CHK(qmg_caxpy(1.0, 0.0, qmg_host::P(x), qmg_host::P(y), n)); }
inline void cxpyz(const qmg_cd* x, const qmg_cd* y, qmg_cd* z, long n) { QMG_CHK(qmg_caxpbyz(1.0, 0.0, qmg_host::P(x), 1.0, 0.0, qmg_host::P(y), qmg_host::P(z), n)); }
template <typename U> inline void cxpay(const qmg_cd* x, U a_, qmg_cd* y, long n) { qmg_cd a(a_); QMG_CHK(qmg_caxpby(1.0, 0.0, qmg_host::P(x), a.real(), a.imag(), qmg_host::P(y), n)); }
template <typename U, typename W> inline void caxpby(U a_, const qmg_cd* x, W b_, qmg_cd* y, long n)
{ qmg_cd a(a_), b(b_); QMG_CHK(qmg_caxpby(a.real(), a.imag(), qmg_host::P(x), b.real(), b.imag(), qmg_host::P(y), n)); }
template <typename U, typename W> inline void caxpbyz(U a_, const qmg_cd* x, W b_, const qmg_cd* y, qmg_cd* z, long n)
{ qmg_cd a(a_), b(b_); QMG_CHK(qmg_caxpbyz(a.real(), a.imag(), qmg_host::P(x), b.real(), b.imag(), qmg_host::P(y), qmg_host::P(z), n)); }
template <typename U, typename W> inline void caxpbypz(U a_, const qmg_cd* x, W b_, const qmg_cd* y, qmg_cd* z, long n)
{ qmg_cd a(a_), b(b_); QMG_CHK(qmg_caxpbypz(a.real(), a.imag(), qmg_host::P(x), b.real(), b.imag(), qmg_host::P(y), qmg_host::P(z), n)); }

// --------------------------------------------------------------- patterns --
// stencil/stencil_2d.h:1526 (pattern lives on the host)
inline void capx_pattern(const qmg_cd* pattern, int len, qmg_cd* v, long nrepeat)
{ QMG_CHK(qmg_cmat_add_pattern(reinterpret_cast<const double*>(pattern), len, qmg_host::P(v), nrepeat)); }
// real pattern (operators/coarse.h:702-721 passes a double array)
inline void capx_pattern(const double* pattern, int len, qmg_cd* v, long nrepeat)
{
  std::vector<qmg_cd> c((size_t)len);
  for (int i = 0; i < len; i++) c[i] = qmg_cd(pattern[i], 0.0);
  capx_pattern(c.data(), len, v, nrepeat);
}
// operators/wilson.h:132 (scale / shuffle live on the host); in == out is allowed for the identity shuffle
inline void caxy_shuffle_pattern(const double* scale, const int* shuffle, int n, const qmg_cd* in, qmg_cd* out, long nsites)
{ QMG_CHK(qmg_shuffle_pattern(scale, shuffle, n, qmg_host::P(in), qmg_host::P(out), nsites)); }

// ------------------------------------------------------------- reductions --
template <typename T> struct RealReducer { typedef T type; };
template <typename T> struct RealReducer<complex<T> > { typedef T type; };
template <typename T> struct Reducer { typedef T type; };
template <typename T> struct ComplexBase { static inline T conj(T x) { return x; } static inline T real(T x) { return x; } };
template <typename T> struct ComplexBase<complex<T> >
{
  static inline complex<T> conj(complex<T> x) { return std::conj(x); }
  static inline T real(complex<T> x) { return std::real(x); }
};

inline double norm2sq(const qmg_cd* v, long n) { double r = 0.0; QMG_CHK(qmg_norm2sq(qmg_host::P(v), n, &r)); return r; }
inline double diffnorm2sq(const qmg_cd* a, const qmg_cd* b, long n) { double r = 0.0; QMG_CHK(qmg_diffnorm2sq(qmg_host::P(a), qmg_host::P(b), n, &r)); return r; }
inline double norminf(const qmg_cd* v, long n) { double r = 0.0; QMG_CHK(qmg_norminf(qmg_host::P(v), n, &r)); return r; }
// <a|b>, conjugate on the first argument (multigrid/stateful_multigrid.h:904)
inline qmg_cd dot(const qmg_cd* a, const qmg_cd* b, long n) { double r[2] = {0.0, 0.0}; QMG_CHK(qmg_dot(qmg_host::P(a), qmg_host::P(b), n, r)); return qmg_cd(r[0], r[1]); }
inline double re_dot(const qmg_cd* a, const qmg_cd* b, long n) { return real(dot(a, b, n)); }
inline void normalize(qmg_cd* v, long n) { const double nrm = sqrt(norm2sq(v, n)); cax(1.0 / nrm, v, n); }
// tests/n13_wilson_kcycle/wilson_kcycle.cpp:348: v -= <against|v>/<against|against> against
inline void orthogonal(qmg_cd* v, const qmg_cd* against, long n) { const qmg_cd c = dot(against, v, n) / norm2sq(against, n); caxpy(-c, against, v, n); }

// -------------------------------------------------------------------- RNG --
// Host draws in the quantum-linalg order (re, im per element from one std::mt19937 stream),
// staged to the device: a driver seeded like the reference's produces the same vectors on both.
// Exception: when the lattice is sharded over several GPUs, for vectors beyond 2^22 elements, or with QMG_DEVICE_RNG=1,
// the fill is the counter-based device generator keyed by two words drawn from `gen` -- a slab cannot reproduce its part
// of one serial mt19937 stream without drawing the whole lattice, and host draws of 10^8 normals take seconds.  The
// device stream is indexed by the GLOBAL element, so 1 GPU and N GPUs draw the same vector.
namespace qmg_host {
inline bool device_rng(long n)
{
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("QMG_DEVICE_RNG"); forced = (e != 0 && e[0] == '1') ? 1 : 0; }
  return forced == 1 || qmg_comm_size() > 1 || n > (1L << 22);
}
}
inline void gaussian(qmg_cd* v, long n, std::mt19937& gen, double dev = 1.0)
{
  if (qmg_host::device_rng(n))
  {
    const unsigned long long hi = gen(), lo = gen();
    QMG_CHK(qmg_gaussian(qmg_host::P(v), n, (hi << 32) | lo, 0ULL, dev));
    return;
  }
  std::vector<qmg_cd> h((size_t)n);
  std::normal_distribution<double> dist(0.0, dev);
  for (long i = 0; i < n; i++) { const double re = dist(gen); const double im = dist(gen); h[i] = qmg_cd(re, im); }
  qmg_host::upload(v, h.data(), n);
}
inline void gaussian_real(qmg_cd* v, long n, std::mt19937& gen, double dev = 1.0)
{
  std::vector<qmg_cd> h((size_t)n);
  std::normal_distribution<double> dist(0.0, dev);
  for (long i = 0; i < n; i++) h[i] = qmg_cd(dist(gen), 0.0);
  qmg_host::upload(v, h.data(), n);
}
// real fields (phases: tests/n04_staggered_test/staggered_test.cpp:144)
inline void random_uniform(double* v, long n, std::mt19937& gen, double lo, double hi)
{
  std::vector<double> h((size_t)n);
  std::uniform_real_distribution<double> dist(lo, hi);
  for (long i = 0; i < n; i++) h[i] = dist(gen);
  qmg_host::check(qmg_memcpy_h2d(v, h.data(), sizeof(double) * (size_t)n), "upload");
}
inline void gaussian(double* v, long n, std::mt19937& gen, double dev = 1.0)
{
  std::vector<double> h((size_t)n);
  std::normal_distribution<double> dist(0.0, dev);
  for (long i = 0; i < n; i++) h[i] = dist(gen);
  qmg_host::check(qmg_memcpy_h2d(v, h.data(), sizeof(double) * (size_t)n), "upload");
}
inline void random_uniform(qmg_cd* v, long n, std::mt19937& gen, double lo, double hi)
{
  std::vector<qmg_cd> h((size_t)n);
  std::uniform_real_distribution<double> dist(lo, hi);
  for (long i = 0; i < n; i++) h[i] = qmg_cd(dist(gen), 0.0);
  qmg_host::upload(v, h.data(), n);
}
// counter-based device fill (no host staging) for large synthetic sources
inline void gaussian_device(qmg_cd* v, long n, unsigned long long seed, unsigned long long stream_id, double dev = 1.0)
{ QMG_CHK(qmg_gaussian(qmg_host::P(v), n, seed, stream_id, dev)); }

#include "generic_local_matrix.h"

#endif
