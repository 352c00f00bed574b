// BLAS-1 on device vectors of complex<double>: the quantum-linalg calls the
// reference makes on the solve path, plus the fused
// Krylov updates.  All kernels are HBM-bound streaming kernels: 16-byte
// (double2) coalesced accesses, grid-stride loops sized to fill every SM,
// warp-shuffle tree reductions finished deterministically by the last block.
#include "qmg_launch.cuh"

using namespace qmg;

#define CD(p) reinterpret_cast<cd*>(p)
#define CCD(p) reinterpret_cast<const cd*>(p)

extern "C" {

int qmg_zero(qmg_cplx* x, long n)
{
  QMG_REQUIRE_INIT();
  if (n <= 0) return 0;
  QMG_CUDA(cudaMemsetAsync(x, 0, sizeof(cd) * (size_t)n, rt().stream));
  rt().launches++;
  return 0;
}

int qmg_copy(qmg_cplx* dst, const qmg_cplx* src, long n)
{
  QMG_REQUIRE_INIT();
  if (n <= 0 || dst == src) return 0;
  QMG_CUDA(cudaMemcpyAsync(dst, src, sizeof(cd) * (size_t)n, cudaMemcpyDeviceToDevice, rt().stream));
  rt().launches++;
  return 0;
}

int qmg_constant(qmg_cplx* x_, double re, double im, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_); const cd v = cmake(re, im);
  return launch_ew(n, [=] __device__(long i) { x[i] = v; });
}

int qmg_cax(double ar, double ai, qmg_cplx* x_, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_); const cd a = cmake(ar, ai);
  return launch_ew(n, [=] __device__(long i) { x[i] = cmul(a, x[i]); });
}

int qmg_caxy(double ar, double ai, const qmg_cplx* x_, qmg_cplx* y_, long n)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); cd* y = CD(y_); const cd a = cmake(ar, ai);
  return launch_ew(n, [=] __device__(long i) { y[i] = cmul(a, x[i]); });
}

int qmg_caxpy(double ar, double ai, const qmg_cplx* x_, qmg_cplx* y_, long n)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); cd* y = CD(y_); const cd a = cmake(ar, ai);
  return launch_ew(n, [=] __device__(long i) { cd t = y[i]; cfma(t, a, x[i]); y[i] = t; });
}

int qmg_caxpby(double ar, double ai, const qmg_cplx* x_, double br, double bi, qmg_cplx* y_, long n)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); cd* y = CD(y_); const cd a = cmake(ar, ai), b = cmake(br, bi);
  return launch_ew(n, [=] __device__(long i) { cd t = cmul(b, y[i]); cfma(t, a, x[i]); y[i] = t; });
}

int qmg_caxpbyz(double ar, double ai, const qmg_cplx* x_, double br, double bi, const qmg_cplx* y_, qmg_cplx* z_, long n)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); const cd* y = CCD(y_); cd* z = CD(z_); const cd a = cmake(ar, ai), b = cmake(br, bi);
  return launch_ew(n, [=] __device__(long i) { cd t = cmul(b, y[i]); cfma(t, a, x[i]); z[i] = t; });
}

int qmg_caxpbypz(double ar, double ai, const qmg_cplx* x_, double br, double bi, const qmg_cplx* y_, qmg_cplx* z_, long n)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); const cd* y = CCD(y_); cd* z = CD(z_); const cd a = cmake(ar, ai), b = cmake(br, bi);
  return launch_ew(n, [=] __device__(long i) { cd t = z[i]; cfma(t, b, y[i]); cfma(t, a, x[i]); z[i] = t; });
}

int qmg_cxty(const qmg_cplx* x_, qmg_cplx* y_, long n)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); cd* y = CD(y_);
  return launch_ew(n, [=] __device__(long i) { y[i] = cmul(x[i], y[i]); });
}

int qmg_conj(qmg_cplx* x_, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_);
  return launch_ew(n, [=] __device__(long i) { x[i] = cconj(x[i]); });
}

int qmg_cinvx(qmg_cplx* x_, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_);
  return launch_ew(n, [=] __device__(long i) { x[i] = cdiv(cmake(1.0, 0.0), x[i]); });
}

int qmg_polar(qmg_cplx* x_, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_);
  return launch_ew(n, [=] __device__(long i) { double s, c; sincos(x[i].x, &s, &c); x[i] = cmake(c, s); });
}

int qmg_elementwise(int op, qmg_cplx* x_, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_);
  switch (op)
  {
    case 0: return launch_ew(n, [=] __device__(long i) { x[i] = cmake(hypot(x[i].x, x[i].y), 0.0); });
    case 1: return launch_ew(n, [=] __device__(long i) { x[i] = cmake(1.0 / sqrt(x[i].x), 0.0); });
    case 2: return launch_ew(n, [=] __device__(long i) { x[i] = cmake(1.0 / hypot(x[i].x, x[i].y), 0.0); });
    case 3: return launch_ew(n, [=] __device__(long i) {
      const double r = 1.0 / sqrt(hypot(x[i].x, x[i].y)), th = atan2(x[i].y, x[i].x);
      double s, c; sincos(th, &s, &c); x[i] = cmake(r * c, r * s); });
    case 4: return launch_ew(n, [=] __device__(long i) { x[i] = cmake(atan2(x[i].y, x[i].x), 0.0); });
    default: return fail_msg("qmg_elementwise: unknown op");
  }
}

int qmg_zero_strided(qmg_cplx* x_, long stride, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_);
  return launch_ew(n, [=] __device__(long i) { x[i * stride] = cmake(0.0, 0.0); });
}

int qmg_constant_strided(qmg_cplx* x_, long stride, double re, double im, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_); const cd v = cmake(re, im);
  return launch_ew(n, [=] __device__(long i) { x[i * stride] = v; });
}

int qmg_caxy_strided(double ar, double ai, const qmg_cplx* x_, long xs, qmg_cplx* y_, long ys, long n, int accumulate)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); cd* y = CD(y_); const cd a = cmake(ar, ai);
  if (accumulate)
    return launch_ew(n, [=] __device__(long i) { cd t = y[i * ys]; cfma(t, a, x[i * xs]); y[i * ys] = t; });
  return launch_ew(n, [=] __device__(long i) { y[i * ys] = cmul(a, x[i * xs]); });
}

int qmg_cax_strided(double ar, double ai, qmg_cplx* x_, long stride, long n)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_); const cd a = cmake(ar, ai);
  return launch_ew(n, [=] __device__(long i) { x[i * stride] = cmul(a, x[i * stride]); });
}

struct ShufflePattern { double scale[64]; int shuffle[64]; };

int qmg_shuffle_pattern(const double* scale_host, const int* shuffle_host, int nc, const qmg_cplx* in_, qmg_cplx* out_, long nsites)
{
  QMG_REQUIRE_INIT();
  if (nc > 64) return fail_msg("qmg_shuffle_pattern: nc > 64 unsupported");
  ShufflePattern pat;
  for (int i = 0; i < nc; i++) { pat.scale[i] = scale_host[i]; pat.shuffle[i] = shuffle_host[i]; }
  const cd* in = CCD(in_); cd* out = CD(out_);
  return launch_ew(nsites * nc, [=] __device__(long e) {
    long s = e / nc; int i = (int)(e - s * nc);
    cd v = in[s * nc + pat.shuffle[i]];
    out[e] = cmake(pat.scale[i] * v.x, pat.scale[i] * v.y);
  });
}

// ---------------------------------------------------------------- reductions --

int qmg_dot(const qmg_cplx* x_, const qmg_cplx* y_, long n, double* result2)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); const cd* y = CCD(y_);
  return launch_reduce<2>(n, [=] __device__(long i, double (&acc)[2]) {
    cd a = x[i], b = y[i];
    acc[0] += a.x * b.x + a.y * b.y;
    acc[1] += a.x * b.y - a.y * b.x;
  }, result2);
}

int qmg_dot_norm(const qmg_cplx* x_, const qmg_cplx* y_, long n, double* result3)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); const cd* y = CCD(y_);
  return launch_reduce<3>(n, [=] __device__(long i, double (&acc)[3]) { dot_acc3(acc, x[i], y[i]); }, result3);
}

int qmg_norm2sq(const qmg_cplx* x_, long n, double* result)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_);
  return launch_reduce<1>(n, [=] __device__(long i, double (&acc)[1]) {
    cd a = x[i];
    acc[0] += a.x * a.x + a.y * a.y;
  }, result);
}

int qmg_diffnorm2sq(const qmg_cplx* x_, const qmg_cplx* y_, long n, double* result)
{
  QMG_REQUIRE_INIT();
  const cd* x = CCD(x_); const cd* y = CCD(y_);
  return launch_reduce<1>(n, [=] __device__(long i, double (&acc)[1]) {
    cd d = csub(x[i], y[i]);
    acc[0] += d.x * d.x + d.y * d.y;
  }, result);
}

} // extern "C"

namespace qmg {
// max-reduction for norminf: reuse the sum machinery on a monotone transform is
// not exact, so do a dedicated two-stage max.
__global__ void __launch_bounds__(kEwBlock) norminf_kernel(const cd* x, long n, double* partials, unsigned int* counter, double* result)
{
  __shared__ double smem[kEwBlock / 32];
  __shared__ bool is_last;
  double m = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    cd a = x[i];
    m = fmax(m, hypot(a.x, a.y));
  }
  auto block_max = [&](double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_down_d(v, o));
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32)
    {
      v = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_down_d(v, o));
    }
    __syncthreads();
    return v;
  };
  m = block_max(m);
  if (threadIdx.x == 0)
  {
    partials[blockIdx.x] = m;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last)
  {
    __threadfence();
    double v = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) v = fmax(v, __ldcg(&partials[b]));
    v = block_max(v);
    if (threadIdx.x == 0) result[0] = v;
    __syncthreads();
    publish_result(counter, result, 1, 1, threadIdx.x, blockDim.x);
  }
}
} // namespace qmg

extern "C" {

int qmg_norminf(const qmg_cplx* x_, long n, double* result)
{
  QMG_REQUIRE_INIT();
  Runtime& r = rt();
  int grid = ew_grid(n > 0 ? n : 1);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  norminf_kernel<<<grid, kEwBlock, 0, r.stream>>>(CCD(x_), n, r.d_partials, r.d_counter, r.d_result);
  QMG_LAUNCH_CHECK();
  return fetch_result(result, 1, 1);
}

} // extern "C"

namespace qmg {

template <int K> struct PtrPack { const cd* p[K]; };
template <int K> struct CoefPack { cd a[K]; };

template <int K>
static int multi_dot_pass(const qmg_cplx* const* xs, int k, const cd* y, long n, double* out2k)
{
  PtrPack<K> pk;
  for (int j = 0; j < K; j++) pk.p[j] = CCD(xs[j < k ? j : 0]);
  double tmp[2 * K];
  // all K + 1 loads of an element are issued before the first product (written as one loop they came out as load, use,
  // load, use ... -- two loads in flight per thread and 0.56 of the copy peak at k = 8, profiles/r03a_kernel_probe_blas.txt)
  int rc = launch_reduce<2 * K>(n, [=] __device__(long i, double (&acc)[2 * K]) {
    const cd b = __ldg(y + i);
    cd a[K];
#pragma unroll
    for (int j = 0; j < K; j++) a[j] = ld_stream(pk.p[j] + i);
#pragma unroll
    for (int j = 0; j < K; j++)
    {
      acc[2 * j] += a[j].x * b.x + a[j].y * b.y;
      acc[2 * j + 1] += a[j].x * b.y - a[j].y * b.x;
    }
  }, tmp);
  if (rc) return rc;
  for (int j = 0; j < k; j++) { out2k[2 * j] = tmp[2 * j]; out2k[2 * j + 1] = tmp[2 * j + 1]; }
  return 0;
}

template <int K>
static int multi_axpy_pass(const double* a_host, const qmg_cplx* const* xs, int k, cd* y, long n, const cd* x0 = nullptr)
{
  if (x0 == nullptr) x0 = y;
  PtrPack<K> pk; CoefPack<K> ck;
  for (int j = 0; j < K; j++)
  {
    pk.p[j] = CCD(xs[j < k ? j : 0]);
    ck.a[j] = j < k ? cmake(a_host[2 * j], a_host[2 * j + 1]) : cmake(0.0, 0.0);
  }
  return launch_ew(n, [=] __device__(long i) {
    cd t = x0[i];
#pragma unroll
    for (int j = 0; j < K; j++) cfma(t, ck.a[j], pk.p[j][i]);
    y[i] = t;
  });
}

} // namespace qmg

extern "C" {

int qmg_multi_dot(const qmg_cplx* const* xs_host, int k, const qmg_cplx* y_, long n, double* result2k)
{
  QMG_REQUIRE_INIT();
  const cd* y = CCD(y_);
  int done = 0;
  while (done < k)
  {
    int left = k - done, rc;
    if (left >= 8) { rc = multi_dot_pass<8>(xs_host + done, 8, y, n, result2k + 2 * done); done += 8; }
    else if (left > 4) { rc = multi_dot_pass<8>(xs_host + done, left, y, n, result2k + 2 * done); done += left; }
    else if (left > 2) { rc = multi_dot_pass<4>(xs_host + done, left, y, n, result2k + 2 * done); done += left; }
    else if (left == 2) { rc = multi_dot_pass<2>(xs_host + done, 2, y, n, result2k + 2 * done); done += 2; }
    else { rc = multi_dot_pass<1>(xs_host + done, 1, y, n, result2k + 2 * done); done += 1; }
    if (rc) return rc;
  }
  return 0;
}

int qmg_multi_axpyz(const double* a_host, const qmg_cplx* const* xs_host, int k, const qmg_cplx* x0_, qmg_cplx* y_, long n)
{
  QMG_REQUIRE_INIT();
  cd* y = CD(y_);
  const cd* x0 = CCD(x0_);
  if (k <= 0) return (x0 != nullptr && x0 != y) ? qmg_copy(y_, x0_, n) : 0;
  int done = 0;
  while (done < k)
  {
    int left = k - done, rc, take;
    if (left >= 8) { take = 8; rc = multi_axpy_pass<8>(a_host + 2 * done, xs_host + done, take, y, n, x0); }
    else if (left > 4) { take = left; rc = multi_axpy_pass<8>(a_host + 2 * done, xs_host + done, take, y, n, x0); }
    else if (left > 2) { take = left; rc = multi_axpy_pass<4>(a_host + 2 * done, xs_host + done, take, y, n, x0); }
    else if (left == 2) { take = 2; rc = multi_axpy_pass<2>(a_host + 2 * done, xs_host + done, take, y, n, x0); }
    else { take = 1; rc = multi_axpy_pass<1>(a_host + 2 * done, xs_host + done, take, y, n, x0); }
    if (rc) return rc;
    done += take;
    x0 = nullptr;   // later passes continue in place
  }
  return 0;
}

int qmg_multi_axpy(const double* a_host, const qmg_cplx* const* xs_host, int k, qmg_cplx* y_, long n)
{
  return qmg_multi_axpyz(a_host, xs_host, k, nullptr, y_, n);
}

int qmg_update_xr_norm(double ar, double ai, const qmg_cplx* p_, const qmg_cplx* q_, qmg_cplx* x_, qmg_cplx* r_, long n, double* result)
{
  QMG_REQUIRE_INIT();
  const cd* p = CCD(p_); const cd* q = CCD(q_); cd* x = CD(x_); cd* r = CD(r_);
  const cd a = cmake(ar, ai), ma = cmake(-ar, -ai);
  return launch_reduce<1>(n, [=] __device__(long i, double (&acc)[1]) {
    cd xi = x[i]; cfma(xi, a, p[i]); x[i] = xi;
    cd ri = r[i]; cfma(ri, ma, q[i]); r[i] = ri;
    acc[0] += ri.x * ri.x + ri.y * ri.y;
  }, result);
}

// One Krylov step with ONE host wait:  alpha = omega <q|r> / <q|q> ;  x += alpha p ;  r -= alpha q ;
// result4 = { |r|^2, Re<q|r>, Im<q|r>, <q|q> }.
// (MR: p = r, q = A r; GCR: p = direction, q = A p, omega = 1.)  The dot products stay on the device: the first kernel
// leaves them (all-reduced) in device memory, every thread of the second forms alpha from them exactly as the host would
// -- (omega d0) / d2, (omega d1) / d2, no contraction -- so the step is bit-identical to qmg_dot_norm followed by
// qmg_update_xr_norm with one round trip instead of two.  The dot products reach the host through a side slot of the
// mapped result buffer, written by element 0's thread.
int qmg_step_xr_norm(double omega, const qmg_cplx* p_, const qmg_cplx* q_, qmg_cplx* x_, qmg_cplx* r_, long n, double* result4)
{
  QMG_REQUIRE_INIT();
  if (n <= 0) return fail_msg("qmg_step_xr_norm: empty vector");
  Runtime& rtm = rt();
  const cd* p = CCD(p_); const cd* q = CCD(q_); cd* x = CD(x_); cd* r = CD(r_);
  double* dres = rtm.d_result + 64;                    // device slot of the dot products
  double* aux = rtm.h_result + 256;                    // mapped host slot the second kernel copies them to
  int rc = launch_reduce_keep<3>(n, [=] __device__(long i, double (&acc)[3]) { dot_acc3(acc, q[i], r[i]); }, dres);
  if (rc) return rc;
  double out[1];
  rc = launch_reduce<1>(n, [=] __device__(long i, double (&acc)[1]) {
    const double d0 = dres[0], d1 = dres[1], d2 = dres[2];
    const cd a = cmake(__ddiv_rn(__dmul_rn(omega, d0), d2), __ddiv_rn(__dmul_rn(omega, d1), d2));
    const cd ma = cmake(-a.x, -a.y);
    if (i == 0) { aux[0] = d0; aux[1] = d1; aux[2] = d2; __threadfence_system(); }
    cd pi = p[i];
    cd xi = x[i]; cfma(xi, a, pi); x[i] = xi;
    cd ri = r[i]; cfma(ri, ma, q[i]); r[i] = ri;
    acc[0] += ri.x * ri.x + ri.y * ri.y;
  }, out);
  if (rc) return rc;
  result4[0] = out[0];
  if (rtm.publish_now) { result4[1] = aux[0]; result4[2] = aux[1]; result4[3] = aux[2]; }
  else
  {
    // copy-and-synchronise mode: the stream has been synchronised by the fetch above
    QMG_CUDA(cudaMemcpy(result4 + 1, dres, sizeof(double) * 3, cudaMemcpyDeviceToHost));
  }
  return 0;
}

} // extern "C"

namespace qmg {
// GCR orthogonalisation with the coefficients formed ON THE DEVICE, the two basis updates in one pass, and the dot
// products of the step that follows folded in:
//   beta_j = -dots_j / apn_j                        (dots_j = <Ap_j|Ap_k> from the multi-dot pass, apn_j = |Ap_j|^2 of step j)
//   Ap_k = (first ? Ap_k : Ap_k) + sum_j beta_j Ap_j ;  p_k = (first ? dir : p_k) + sum_j beta_j p_j
//   last pass only: { <Ap_k|r>, |Ap_k|^2 } -> result (device)
// One pass holds K vectors; the host chains passes of 8 in the order qmg_multi_axpyz uses, so the sums come out bit-identical.
template <int K, bool DOTS, bool WITHP>
__global__ void __launch_bounds__(kEwBlock) gcr_ortho_kernel(long n, PtrPack<K> Ap, PtrPack<K> P, int k, cd* apk, const cd* dir,
                                                             cd* pk, const cd* __restrict__ r, const double* __restrict__ dots,
                                                             const double* __restrict__ apn, double* partials, unsigned int* counter, double* result)
{
  __shared__ double smem[(kEwBlock / 32) * 3];
  cd beta[K];
#pragma unroll
  for (int j = 0; j < K; j++)
  {
    const double nr = (j < k) ? apn[j] : 1.0;
    beta[j] = (j < k) ? cmake(__ddiv_rn(-dots[2 * j], nr), __ddiv_rn(-dots[2 * j + 1], nr)) : cmake(0.0, 0.0);
  }
  double acc[3] = {0.0, 0.0, 0.0};
  const long stride = (long)gridDim.x * blockDim.x;
  // two sweeps, one per basis (each the loop of multi_axpy_pass, which runs at the copy peak; a single sweep over both
  // bases keeps 2 K + 3 loads per element alive, which ptxas serialised: 3.8 TB/s)
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    cd a[K];
#pragma unroll
    for (int j = 0; j < K; j++) a[j] = ld_stream(Ap.p[j] + i);
    cd t = apk[i];
    cd rr = DOTS ? __ldg(r + i) : cmake(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < K; j++) cfma(t, beta[j], a[j]);
    apk[i] = t;
    if (DOTS) dot_acc3(acc, t, rr);
  }
  if (WITHP)
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    cd b[K];
#pragma unroll
    for (int j = 0; j < K; j++) b[j] = ld_stream(P.p[j] + i);
    cd u = dir[i];
#pragma unroll
    for (int j = 0; j < K; j++) cfma(u, beta[j], b[j]);
    pk[i] = u;
  }
  if (DOTS) grid_reduce_finish<3>(acc, smem, partials, counter, result);
}

template <int K, bool WITHP>
static int gcr_ortho_pass(long n, const qmg_cplx* const* Ap, const qmg_cplx* const* P, int k, cd* apk, const cd* dir, cd* pk, const cd* r,
                          const double* dots, const double* apn, bool with_dots, double* result_dev)
{
  PtrPack<K> a, b;
  for (int j = 0; j < K; j++) { a.p[j] = CCD(Ap[j < k ? j : 0]); b.p[j] = WITHP ? CCD(P[j < k ? j : 0]) : nullptr; }
  Runtime& rtm = rt();
  const int grid = with_dots ? resident_grid(gcr_ortho_kernel<K, true, WITHP>, n) : resident_grid(gcr_ortho_kernel<K, false, WITHP>, n);
  if (with_dots)
  {
    gcr_ortho_kernel<K, true, WITHP><<<grid, kEwBlock, 0, rtm.stream>>>(n, a, b, k, apk, dir, pk, r, dots, apn, rtm.d_partials, rtm.d_counter, result_dev);
    QMG_LAUNCH_CHECK();
    return skip_result(result_dev, 3);
  }
  gcr_ortho_kernel<K, false, WITHP><<<grid, kEwBlock, 0, rtm.stream>>>(n, a, b, k, apk, dir, pk, r, dots, apn, nullptr, nullptr, nullptr);
  QMG_LAUNCH_CHECK();
  return 0;
}
} // namespace qmg

extern "C" {

// Orthogonalise the new GCR direction against the k stored ones and prepare the step, without a host round trip:
//   dots_dev[2j..] = <Ap[j]|Ap_k>  (multi-dot, kept on the device) ; beta_j = -dots_j / apn_dev[j] ;
//   Ap_k += sum beta_j Ap[j] ; p_k = dir + sum beta_j p[j] ; { <Ap_k|r>, |Ap_k|^2 } left where qmg_krylov_step with
//   QMG_STEP_DOTS_READY expects them.  Same sums in the same order as qmg_multi_dot + host division + 2 qmg_multi_axpyz +
//   the dot pass of qmg_krylov_step: bit-identical.  k = 0: only the dot products (p_k = dir is the caller's copy or alias).
int qmg_gcr_orthogonalize(const qmg_cplx* const* Ap_host, const qmg_cplx* const* p_host, int k, qmg_cplx* Apk_, const qmg_cplx* dir_, qmg_cplx* pk_,
                          const qmg_cplx* r_, long n, double* dots_dev, const double* apn_dev)
{
  QMG_REQUIRE_INIT();
  if (n <= 0) return fail_msg("qmg_gcr_orthogonalize: empty vector");
  if (k < 1) return fail_msg("qmg_gcr_orthogonalize: k >= 1 stored directions needed");
  cd* apk = CD(Apk_); const cd* dir = CCD(dir_); cd* pk = CD(pk_); const cd* r = CCD(r_);
  double* dres = rt().d_result + 64;
  int rc;
  // 1. the k projections, left on the device (passes of at most 8 like qmg_multi_dot)
  int done = 0;
  while (done < k)
  {
    const int left = k - done;
    int take;
    const cd* y = apk;
#define QMG_MDOT_KEEP(KK) { PtrPack<KK> pk_; for (int j = 0; j < KK; j++) pk_.p[j] = CCD(Ap_host[done + (j < take ? j : 0)]); \
      rc = launch_reduce_keep<2 * KK>(n, [=] __device__(long i, double (&acc)[2 * KK]) { const cd b = __ldg(y + i); cd a[KK]; \
        _Pragma("unroll") for (int j = 0; j < KK; j++) a[j] = ld_stream(pk_.p[j] + i); \
        _Pragma("unroll") for (int j = 0; j < KK; j++) { acc[2 * j] += a[j].x * b.x + a[j].y * b.y; acc[2 * j + 1] += a[j].x * b.y - a[j].y * b.x; } }, dots_dev + 2 * done); }
    if (left >= 8) { take = 8; QMG_MDOT_KEEP(8) }
    else if (left > 4) { take = left; QMG_MDOT_KEEP(8) }
    else if (left > 2) { take = left; QMG_MDOT_KEEP(4) }
    else if (left == 2) { take = 2; QMG_MDOT_KEEP(2) }
    else { take = 1; QMG_MDOT_KEEP(1) }
#undef QMG_MDOT_KEEP
    if (rc) return rc;
    done += take;
  }
  // 2. both basis updates per pass of 8; the last pass also forms the step's dot products
  done = 0;
  const cd* src_dir = dir;
  do
  {
    const int left = k - done;
    const int take = left >= 8 ? 8 : left;
    const bool last = (done + take >= k);
    const qmg_cplx* const* A = Ap_host + done; const qmg_cplx* const* Pp = (p_host != nullptr) ? p_host + done : nullptr;
    const double* dd = dots_dev + 2 * done; const double* nn = apn_dev + done;
    if (p_host != nullptr)
    {
      if (take > 4) rc = gcr_ortho_pass<8, true>(n, A, Pp, take, apk, src_dir, pk, r, dd, nn, last, dres);
      else if (take > 2) rc = gcr_ortho_pass<4, true>(n, A, Pp, take, apk, src_dir, pk, r, dd, nn, last, dres);
      else if (take == 2) rc = gcr_ortho_pass<2, true>(n, A, Pp, take, apk, src_dir, pk, r, dd, nn, last, dres);
      else rc = gcr_ortho_pass<1, true>(n, A, Pp, take, apk, src_dir, pk, r, dd, nn, last, dres);
    }
    else
    {
      // the directions are kept unorthogonalised (the solver forms x from them once, at the end): only the A p basis moves
      if (take > 4) rc = gcr_ortho_pass<8, false>(n, A, A, take, apk, nullptr, nullptr, r, dd, nn, last, dres);
      else if (take > 2) rc = gcr_ortho_pass<4, false>(n, A, A, take, apk, nullptr, nullptr, r, dd, nn, last, dres);
      else if (take == 2) rc = gcr_ortho_pass<2, false>(n, A, A, take, apk, nullptr, nullptr, r, dd, nn, last, dres);
      else rc = gcr_ortho_pass<1, false>(n, A, A, take, apk, nullptr, nullptr, r, dd, nn, last, dres);
    }
    if (rc) return rc;
    done += take;
    src_dir = pk;        // later passes continue in place
  } while (done < k);
  return 0;
}

// The general Krylov step behind the K-cycle's smoothers and (flexible) GCR solves: the same two kernels as
// qmg_step_xr_norm -- so every sum is formed in the same order and a step without flags is bit-identical to it -- with
// the start-up and wind-down work of a solve folded in:
//   alpha = omega <q|r_in> / <q|q>
//   t = (x_in ? x_in : 0) + alpha p ;  x_out = (acc ? acc + t : t)
//   r_out = r_in - alpha q ;  |r_out|^2                       (skipped with QMG_STEP_X_ONLY: no reduction, no host wait)
// x_in == NULL is the first step of a solve from a zero start (x is written, never read, so nobody has to zero it);
// r_in != r_out is the same first step reading the right-hand side in place of a copied residual; acc folds the
// "lhs += z" that follows a smoother into its last step.  QMG_STEP_WANT_RNORM adds |r_in|^2 (the |b|^2 every solver
// needs) to the dot-product pass, which reads r_in anyway.
// result5 = { |r_out|^2, Re<q|r_in>, Im<q|r_in>, <q|q>, |r_in|^2 }; with QMG_STEP_X_ONLY nothing is returned.
int qmg_krylov_step(double omega, const qmg_cplx* p_, const qmg_cplx* q_, const qmg_cplx* x_in_, qmg_cplx* x_out_,
                    const qmg_cplx* r_in_, qmg_cplx* r_out_, const qmg_cplx* acc_, long n, int flags, double* result5, double* qq_dev)
{
  QMG_REQUIRE_INIT();
  if (n <= 0) return fail_msg("qmg_krylov_step: empty vector");
  Runtime& rtm = rt();
  const cd* p = CCD(p_); const cd* q = CCD(q_); const cd* xin = CCD(x_in_); cd* x = CD(x_out_);
  const cd* rin = CCD(r_in_); cd* r = CD(r_out_); const cd* accv = CCD(acc_);
  double* dres = rtm.d_result + 64;                    // device slot of the dot products
  double* aux = rtm.h_result + 256;                    // mapped host slot the second kernel copies them to
  const bool want_rnorm = (flags & QMG_STEP_WANT_RNORM) != 0;
  int rc = 0;
  if (flags & QMG_STEP_DOTS_READY)
  {
    // <q|r_in>, |q|^2 are already in the device slot (qmg_gcr_orthogonalize left them there)
    if (want_rnorm) return fail_msg("qmg_krylov_step: QMG_STEP_DOTS_READY and QMG_STEP_WANT_RNORM exclude each other");
  }
  else if (want_rnorm)
    rc = launch_reduce_keep<4>(n, [=] __device__(long i, double (&acc)[4]) {
      const cd a = q[i], b = rin[i];
      double a3[3] = { acc[0], acc[1], acc[2] };
      dot_acc3(a3, a, b);
      acc[0] = a3[0]; acc[1] = a3[1]; acc[2] = a3[2];
      acc[3] += b.x * b.x + b.y * b.y;
    }, dres);
  else
    rc = launch_reduce_keep<3>(n, [=] __device__(long i, double (&acc)[3]) { dot_acc3(acc, q[i], rin[i]); }, dres);
  if (rc) return rc;
  if (flags & QMG_STEP_R_ONLY)
  {
    // GCR with the solution formed once at the end: only the residual recurrence runs per step
    double out1[1];
    rc = launch_reduce<1>(n, [=] __device__(long i, double (&acc)[1]) {
      const double d0 = dres[0], d1 = dres[1], d2 = dres[2];
      const cd a = cmake(__ddiv_rn(__dmul_rn(omega, d0), d2), __ddiv_rn(__dmul_rn(omega, d1), d2));
      const cd ma = cmake(-a.x, -a.y);
      if (i == 0) { aux[0] = d0; aux[1] = d1; aux[2] = d2; if (qq_dev != nullptr) qq_dev[0] = d2; __threadfence_system(); }
      cd ri = rin[i];
      cfma(ri, ma, q[i]); r[i] = ri;
      acc[0] += ri.x * ri.x + ri.y * ri.y;
    }, out1);
    if (rc) return rc;
    result5[0] = out1[0];
    if (rtm.publish_now) { result5[1] = aux[0]; result5[2] = aux[1]; result5[3] = aux[2]; result5[4] = 0.0; }
    else { QMG_CUDA(cudaMemcpy(result5 + 1, dres, sizeof(double) * 3, cudaMemcpyDeviceToHost)); result5[4] = 0.0; }
    return 0;
  }
  if (flags & QMG_STEP_NO_NORM)
  {
    if (flags & (QMG_STEP_X_ONLY | QMG_STEP_R_ONLY | QMG_STEP_WANT_RNORM)) return fail_msg("qmg_krylov_step: QMG_STEP_NO_NORM excludes X_ONLY, R_ONLY and WANT_RNORM");
    // the element arithmetic of the general step below, without its reduction
    return launch_ew(n, [=] __device__(long i) {
      const double d0 = dres[0], d1 = dres[1], d2 = dres[2];
      const cd a = cmake(__ddiv_rn(__dmul_rn(omega, d0), d2), __ddiv_rn(__dmul_rn(omega, d1), d2));
      const cd ma = cmake(-a.x, -a.y);
      cd pi = p[i];
      cd ri = rin[i];
      cd xi = (xin != nullptr) ? xin[i] : cmake(0.0, 0.0);
      cfma(xi, a, pi);
      if (accv != nullptr) xi = cadd(accv[i], xi);
      x[i] = xi;
      cfma(ri, ma, q[i]); r[i] = ri;
    });
  }
  if (flags & QMG_STEP_X_ONLY)
    return launch_ew(n, [=] __device__(long i) {
      const double d0 = dres[0], d1 = dres[1], d2 = dres[2];
      const cd a = cmake(__ddiv_rn(__dmul_rn(omega, d0), d2), __ddiv_rn(__dmul_rn(omega, d1), d2));
      cd xi = (xin != nullptr) ? xin[i] : cmake(0.0, 0.0);
      cfma(xi, a, p[i]);
      if (accv != nullptr) xi = cadd(accv[i], xi);
      x[i] = xi;
    });
  double out[1];
  rc = launch_reduce<1>(n, [=] __device__(long i, double (&acc)[1]) {
    const double d0 = dres[0], d1 = dres[1], d2 = dres[2];
    const cd a = cmake(__ddiv_rn(__dmul_rn(omega, d0), d2), __ddiv_rn(__dmul_rn(omega, d1), d2));
    const cd ma = cmake(-a.x, -a.y);
    if (i == 0) { aux[0] = d0; aux[1] = d1; aux[2] = d2; if (want_rnorm) aux[3] = dres[3]; if (qq_dev != nullptr) qq_dev[0] = d2; __threadfence_system(); }
    cd pi = p[i];
    cd ri = rin[i];
    cd xi = (xin != nullptr) ? xin[i] : cmake(0.0, 0.0);
    cfma(xi, a, pi);
    if (accv != nullptr) xi = cadd(accv[i], xi);
    x[i] = xi;
    cfma(ri, ma, q[i]); r[i] = ri;
    acc[0] += ri.x * ri.x + ri.y * ri.y;
  }, out);
  if (rc) return rc;
  result5[0] = out[0];
  if (rtm.publish_now) { result5[1] = aux[0]; result5[2] = aux[1]; result5[3] = aux[2]; result5[4] = want_rnorm ? aux[3] : 0.0; }
  else
  {
    // copy-and-synchronise mode: the stream has been synchronised by the fetch above
    QMG_CUDA(cudaMemcpy(result5 + 1, dres, sizeof(double) * (want_rnorm ? 4 : 3), cudaMemcpyDeviceToHost));
    if (!want_rnorm) result5[4] = 0.0;
  }
  return 0;
}

} // extern "C"

// ------------------------------------------------------- two MR steps from a zero start in two passes --
// The K-cycle's smoother is MR(2) from a zero start: r1 = r0 - a1 A r0, x = a1 r0 + a2 r1, r2 = r1 - a2 A r1.  Applying A to
// q1 = A r0 instead of to r1 (A r1 = q1 - a1 A q1: the same number of applies) leaves nothing between the two applies, and both
// step lengths follow from ONE pass of dot products over r0, q1 = A r0, p2 = A q1:
//   a1 = w <q1|r0> / <q1|q1> ;  <q2|r1> = <q1|r0> - a1 <q1|q1> - conj(a1) <p2|r0> + |a1|^2 <p2|q1> ;
//   <q2|q2> = <q1|q1> - 2 Re(conj(a1) <p2|q1>) + |a1|^2 <p2|p2> ;  a2 = w <q2|r1> / <q2|q2>
//   x = (a1 + a2) r0 - a1 a2 q1 ;  r2 = r0 - (a1 + a2) q1 + a1 a2 p2
// -- 3 + 5 vector passes and one host wait instead of 13 passes and two.  The iterates are those of MR(2) up to rounding.
extern "C" {

// out9 = { Re<q1|r0>, Im<q1|r0>, <q1|q1>, Re<p2|r0>, Im<p2|r0>, Re<p2|q1>, Im<p2|q1>, <p2|p2>, <r0|r0> }   (<x|y> = sum conj(x) y)
int qmg_mr2_gram(const qmg_cplx* r0_, const qmg_cplx* q1_, const qmg_cplx* p2_, long n, double* out9)
{
  QMG_REQUIRE_INIT();
  if (n <= 0) return fail_msg("qmg_mr2_gram: empty vector");
  const cd* r0 = CCD(r0_); const cd* q1 = CCD(q1_); const cd* p2 = CCD(p2_);
  return launch_reduce<9>(n, [=] __device__(long i, double (&acc)[9]) {
    const cd r = r0[i], q = q1[i], p = p2[i];
    acc[0] += q.x * r.x + q.y * r.y; acc[1] += q.x * r.y - q.y * r.x;
    acc[2] += q.x * q.x + q.y * q.y;
    acc[3] += p.x * r.x + p.y * r.y; acc[4] += p.x * r.y - p.y * r.x;
    acc[5] += p.x * q.x + p.y * q.y; acc[6] += p.x * q.y - p.y * q.x;
    acc[7] += p.x * p.x + p.y * p.y;
    acc[8] += r.x * r.x + r.y * r.y;
  }, out9);
}

// x_out = (acc ? acc : 0) + cx0 r0 + cx1 q1 ;  r_out = r0 + cr1 q1 + cr2 p2 (if r_out != NULL; may alias r0).  Coefficients: 2 doubles each.
int qmg_mr2_update(const double* cx0, const double* cx1, const double* cr1, const double* cr2, const qmg_cplx* r0_, const qmg_cplx* q1_, const qmg_cplx* p2_,
                   const qmg_cplx* acc_, qmg_cplx* x_out_, qmg_cplx* r_out_, long n)
{
  QMG_REQUIRE_INIT();
  if (n <= 0) return fail_msg("qmg_mr2_update: empty vector");
  const cd* r0 = CCD(r0_); const cd* q1 = CCD(q1_); const cd* p2 = CCD(p2_); const cd* accv = CCD(acc_); cd* x = CD(x_out_); cd* rr = CD(r_out_);
  const cd a0 = cmake(cx0[0], cx0[1]), a1 = cmake(cx1[0], cx1[1]), b1 = cmake(cr1[0], cr1[1]), b2 = cmake(cr2[0], cr2[1]);
  return launch_ew(n, [=] __device__(long i) {
    const cd r = r0[i], q = q1[i];
    cd xi = cmul(a0, r);
    cfma(xi, a1, q);
    if (accv != nullptr) xi = cadd(accv[i], xi);
    x[i] = xi;
    if (rr != nullptr)
    {
      cd t = r;
      cfma(t, b1, q);
      cfma(t, b2, p2[i]);
      rr[i] = t;
    }
  });
}

} // extern "C"

// ------------------------------------------------------- BiCGstab(L) sweeps with their vector traffic cut in half --
// One sweep of BiCGstab(L) as the reference's set-up runs it (tests/n13_wilson_kcycle/wilson_kcycle.cpp:359, L = 6) is 2 L
// operator applies and, written call by call, 280 vector reads / writes of BLAS-1 -- twice the bytes of the applies on the
// fine level.  Three kernels bring that to 148 without changing one floating-point operation of an element:
//  * qmg_bicgstab_replay.  In the BiCG part only the TOP vectors r_j, u_j of step j feed the applies and the dot products;
//    the updates of the lower ones (u_i = r_i - beta_j u_i, r_i -= alpha_j u_{i+1}, i < j, and x += alpha_j u_0) only have
//    to be complete when the MR part starts.  The solver performs the top updates at once and this kernel replays the L
//    steps of the lower ones for an element in registers: 2 L reads + 2 L - 1 writes instead of 3 L (L - 1) + 3 L.
//  * qmg_bicgstab_mgs_pass.  The modified Gram-Schmidt of the MR part in right-looking order (the same updates of every
//    r_j in the same order, every projection taken of the same updated vector): pass i forms tau_ij = <r_i|r_j> / sigma_i
//    on the device from the sums of pass i - 1, subtracts tau_ij r_i from ALL later r_j in one sweep, and, r_{i+1} being
//    final now, accumulates sigma_{i+1}, <r_{i+1}|r_0> and <r_{i+1}|r_j> (j > i + 1) in the same sweep.
//  * qmg_bicgstab_finish.  x += sum c_j r_j, r_0 -= sum g_j r_j and |r_0|^2 in one pass over r_0 .. r_L.
namespace qmg {
constexpr int kBicgMaxL = 7;                 // pass 0 reduces 2 L + 1 sums
constexpr int kMgsStride = 2 * kBicgMaxL + 2;     // doubles per pass in the device table of the MGS sums

template <int L> struct BicgReplayPack { cd* r[L]; cd* u[L]; cd* x; cd alpha[L]; cd beta[L]; };

template <int L>
__global__ void __launch_bounds__(kEwBlock) bicg_replay_kernel(long n, BicgReplayPack<L> a)
{
  const cd one = cmake(1.0, 0.0);
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
  {
    cd R[L > 1 ? L - 1 : 1], U[L];
#pragma unroll
    for (int k = 0; k < L - 1; k++) R[k] = a.r[k][e];
#pragma unroll
    for (int k = 0; k < L; k++) U[k] = a.u[k][e];
    cd X = a.x[e];
#pragma unroll
    for (int j = 0; j < L; j++)
    {
      const cd mb = cmake(-a.beta[j].x, -a.beta[j].y), ma = cmake(-a.alpha[j].x, -a.alpha[j].y);
      // u_i = r_i - beta_j u_i  (qmg_caxpby(1, r_i, -beta, u_i)), i < j
#pragma unroll
      for (int k = 0; k < j; k++) { cd t = cmul(mb, U[k]); cfma(t, one, R[k]); U[k] = t; }
      // r_i -= alpha_j u_{i+1}  (qmg_caxpy(-alpha, u_{i+1}, r_i)), i < j: u_j is the step's top vector, as stored
#pragma unroll
      for (int k = 0; k < j; k++) cfma(R[k], ma, U[k + 1]);
      // x += alpha_j u_0
      cfma(X, a.alpha[j], U[0]);
    }
#pragma unroll
    for (int k = 0; k < L - 1; k++) { a.r[k][e] = R[k]; a.u[k][e] = U[k]; }
    a.x[e] = X;
  }
}

template <int M> struct MgsPack { const cd* ri; const cd* r0; cd* v[M]; };

// pass over r_i (null in pass 0) and the M vectors after it; prev: the sums of the pass before (null in pass 0)
template <int M>
static int mgs_pass(long n, const MgsPack<M>& pk, const double* prev, double* out_dev)
{
  const MgsPack<M> a = pk;
  return launch_reduce_keep<2 * M + 1>(n, [=] __device__(long e, double (&acc)[2 * M + 1]) {
    cd v[M];
#pragma unroll
    for (int m = 0; m < M; m++) v[m] = a.v[m][e];
    const cd b0 = __ldg(a.r0 + e);
    if (prev != nullptr)
    {
      const cd ri = __ldg(a.ri + e);
      const double sg = prev[0];
#pragma unroll
      for (int m = 0; m < M; m++)
      {
        // tau = <r_i|r_j> / sigma_i as the host forms it (complex / double), r_j -= tau r_i as qmg_caxpy(-tau, r_i, r_j)
        const cd mt = cmake(-__ddiv_rn(prev[3 + 2 * m], sg), -__ddiv_rn(prev[4 + 2 * m], sg));
        cfma(v[m], mt, ri);
        a.v[m][e] = v[m];
      }
    }
    const cd y = v[0];
    acc[0] += y.x * y.x + y.y * y.y;
    acc[1] += y.x * b0.x + y.y * b0.y;
    acc[2] += y.x * b0.y - y.y * b0.x;
#pragma unroll
    for (int m = 1; m < M; m++)
    {
      acc[1 + 2 * m] += y.x * v[m].x + y.y * v[m].y;
      acc[2 + 2 * m] += y.x * v[m].y - y.y * v[m].x;
    }
  }, out_dev);
}

template <int L> struct BicgFinishPack { const cd* r[L + 1]; cd* x; cd* r0; cd cx[L]; cd cr[L]; };
} // namespace qmg

extern "C" {

int qmg_bicgstab_max_l(void) { return kBicgMaxL; }

// r_host, u_host: HOST arrays of L device pointers r_0 .. r_{L-1}, u_0 .. u_{L-1}; alpha_host, beta_host: 2 L doubles each.
int qmg_bicgstab_replay(int L, qmg_cplx* const* r_host, qmg_cplx* const* u_host, qmg_cplx* x_, const double* alpha_host, const double* beta_host, long n)
{
  QMG_REQUIRE_INIT();
  if (L < 1 || L > kBicgMaxL) return fail_msg("qmg_bicgstab_replay: 1 <= L <= 7");
  if (n <= 0) return fail_msg("qmg_bicgstab_replay: empty vector");
#define QMG_REPLAY(LL) case LL: { BicgReplayPack<LL> a; for (int k = 0; k < LL; k++) { a.r[k] = CD(r_host[k]); a.u[k] = CD(u_host[k]); \
      a.alpha[k] = cmake(alpha_host[2 * k], alpha_host[2 * k + 1]); a.beta[k] = cmake(beta_host[2 * k], beta_host[2 * k + 1]); } a.x = CD(x_); \
      bicg_replay_kernel<LL><<<ew_grid(n), kEwBlock, 0, rt().stream>>>(n, a); break; }
  switch (L) { QMG_REPLAY(1) QMG_REPLAY(2) QMG_REPLAY(3) QMG_REPLAY(4) QMG_REPLAY(5) QMG_REPLAY(6) QMG_REPLAY(7) default: break; }
#undef QMG_REPLAY
  QMG_LAUNCH_CHECK();
  return 0;
}

// The whole modified Gram-Schmidt of the MR part: L passes, no host wait in between, one read-back at the end.
// r_host: L + 1 device pointers r_0 .. r_L (r_1 .. r_L are orthogonalised in place).  sums_host: L rows of 2 L + 2 doubles;
// row i - 1 = { sigma_i, Re <r_i|r_0>, Im <r_i|r_0>, Re <r_i|r_{i+1}>, Im <r_i|r_{i+1}>, ... , <r_i|r_L> } with every r final
// on the left and updated against r_1 .. r_{i-1} on the right: tau_ij = <r_i|r_j> / sigma_i, gamma'_i = <r_i|r_0> / sigma_i.
int qmg_bicgstab_mgs(int L, qmg_cplx* const* r_host, long n, double* sums_host)
{
  QMG_REQUIRE_INIT();
  if (L < 1 || L > kBicgMaxL) return fail_msg("qmg_bicgstab_mgs: 1 <= L <= 7");
  if (n <= 0) return fail_msg("qmg_bicgstab_mgs: empty vector");
  // the table of the passes' sums: a block of the caching allocator (returned below; the stream orders its next use after this one)
  void* tab_v = nullptr;
  int rc = qmg_malloc(&tab_v, sizeof(double) * kBicgMaxL * kMgsStride);
  if (rc) return rc;
  double* tab = static_cast<double*>(tab_v);
  for (int i = 0; i < L && rc == 0; i++)
  {
    // pass i: r_i (i >= 1) is subtracted from r_{i+1} .. r_L, then the sums of r_{i+1}
    const int M = L - i;
    const double* prev = (i > 0) ? tab + (size_t)(i - 1) * kMgsStride : nullptr;
    double* out = tab + (size_t)i * kMgsStride;
#define QMG_MGS(MM) case MM: { MgsPack<MM> pk; pk.ri = (i > 0) ? CCD(r_host[i]) : nullptr; pk.r0 = CCD(r_host[0]); \
      for (int m = 0; m < MM; m++) pk.v[m] = CD(r_host[i + 1 + m]); rc = mgs_pass<MM>(n, pk, prev, out); break; }
    switch (M) { QMG_MGS(1) QMG_MGS(2) QMG_MGS(3) QMG_MGS(4) QMG_MGS(5) QMG_MGS(6) QMG_MGS(7) default: break; }
#undef QMG_MGS
  }
  if (rc) { qmg_free(tab_v); return rc; }
  cudaError_t ce = cudaMemcpyAsync(sums_host, tab, sizeof(double) * L * kMgsStride, cudaMemcpyDeviceToHost, rt().stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(rt().stream);
  qmg_free(tab_v);
  if (ce != cudaSuccess) return fail_msg(cudaGetErrorString(ce));
  return 0;
}

// x += sum_{j<L} cx_j r_j ; r_0 -= ... written as r_0 += sum_{j=1..L} cr_j r_j ; result = |r_0|^2.
// r_host: L + 1 device pointers; cx_host: 2 L doubles (coefficients of r_0 .. r_{L-1}); cr_host: 2 L doubles (of r_1 .. r_L).
// The sums run in the order of qmg_multi_axpy (one chain per output, j ascending), so x and r_0 come out as from the two calls.
int qmg_bicgstab_finish(int L, qmg_cplx* const* r_host, qmg_cplx* x_, const double* cx_host, const double* cr_host, long n, double* result)
{
  QMG_REQUIRE_INIT();
  if (L < 1 || L > kBicgMaxL) return fail_msg("qmg_bicgstab_finish: 1 <= L <= 7");
  if (n <= 0) return fail_msg("qmg_bicgstab_finish: empty vector");
  int rc = 1;
#define QMG_FIN(LL) case LL: { BicgFinishPack<LL> a; for (int k = 0; k <= LL; k++) a.r[k] = CCD(r_host[k]); a.x = CD(x_); a.r0 = CD(r_host[0]); \
      for (int k = 0; k < LL; k++) { a.cx[k] = cmake(cx_host[2 * k], cx_host[2 * k + 1]); a.cr[k] = cmake(cr_host[2 * k], cr_host[2 * k + 1]); } \
      rc = launch_reduce<1>(n, [=] __device__(long e, double (&acc)[1]) { \
        cd v[LL + 1]; \
        _Pragma("unroll") for (int k = 0; k <= LL; k++) v[k] = ld_stream(a.r[k] + e); \
        cd xx = a.x[e]; \
        _Pragma("unroll") for (int k = 0; k < LL; k++) cfma(xx, a.cx[k], v[k]); \
        a.x[e] = xx; \
        cd rr = v[0]; \
        _Pragma("unroll") for (int k = 0; k < LL; k++) cfma(rr, a.cr[k], v[k + 1]); \
        a.r0[e] = rr; \
        acc[0] += rr.x * rr.x + rr.y * rr.y; }, result); break; }
  switch (L) { QMG_FIN(1) QMG_FIN(2) QMG_FIN(3) QMG_FIN(4) QMG_FIN(5) QMG_FIN(6) QMG_FIN(7) default: break; }
#undef QMG_FIN
  return rc;
}

} // extern "C"

// ------------------------------------------------------- time-slice reductions --
// sum over x and colour for every row y (reductions/reductions.h:24-92: norm2sq / re_dot / dot per time slice, the
// correlator measurement of tests/n16_wilson_kcycle_heatbath).  In the even-odd layout row y is two contiguous spans of
// (X/2) nc elements, one per parity: one CTA per row streams both and writes out[y] (deterministic tree sum).
namespace qmg {
template <int OP>   // 0 |a|^2, 1 Re conj(a) b, 2 conj(a) b
__global__ void __launch_bounds__(256) timeslice_kernel(const cd* __restrict__ a, const cd* __restrict__ b, int rowlen, long half, double* __restrict__ out)
{
  __shared__ double smem[8 * 2];
  const int y = blockIdx.x;
  double acc[2] = {0.0, 0.0};
  for (int p = 0; p < 2; p++)
  {
    const cd* ra = a + (size_t)p * half + (size_t)y * rowlen;
    const cd* rb = (OP == 0) ? ra : b + (size_t)p * half + (size_t)y * rowlen;
    for (int i = threadIdx.x; i < rowlen; i += blockDim.x)
    {
      const cd u = ra[i];
      if (OP == 0) acc[0] += u.x * u.x + u.y * u.y;
      else
      {
        const cd v = rb[i];
        acc[0] += u.x * v.x + u.y * v.y;
        if (OP == 2) acc[1] += u.x * v.y - u.y * v.x;
      }
    }
  }
  block_sum<2>(acc, smem);
  if (threadIdx.x == 0)
  {
    if (OP == 2) { out[2 * y] = acc[0]; out[2 * y + 1] = acc[1]; }
    else out[y] = acc[0];
  }
}
} // namespace qmg

// op: 0 norm2sq_cv_timeslice(a), 1 redot_cv_timeslice(a, b), 2 dot_cv_timeslice(a, b); host_out: Y doubles (op 2: 2 Y).
// On a y-slab the rows are this rank's rows (no exchange: every row lives on one rank).
extern "C" int qmg_timeslice_reduce(int op, const qmg_cplx* a_, const qmg_cplx* b_, int X, int Y, int nc, double* host_out)
{
  QMG_REQUIRE_INIT();
  if (X < 2 || Y < 1 || (X & 1) || nc < 1) return fail_msg("qmg_timeslice_reduce: bad lattice");
  if (op < 0 || op > 2 || (op > 0 && b_ == nullptr)) return fail_msg("qmg_timeslice_reduce: bad op / missing second vector");
  Runtime& r = rt();
  const int rowlen = (X / 2) * nc; const long half = (long)rowlen * Y;
  const size_t nout = (size_t)Y * (op == 2 ? 2 : 1);
  double* d_out = ensure_partials(nout > (size_t)kMaxRedBlocks * kMaxRedWidth ? nout : (size_t)kMaxRedBlocks * kMaxRedWidth);
  if (d_out == nullptr) return 1;
  if (op == 0) timeslice_kernel<0><<<Y, 256, 0, r.stream>>>(CCD(a_), nullptr, rowlen, half, d_out);
  else if (op == 1) timeslice_kernel<1><<<Y, 256, 0, r.stream>>>(CCD(a_), CCD(b_), rowlen, half, d_out);
  else timeslice_kernel<2><<<Y, 256, 0, r.stream>>>(CCD(a_), CCD(b_), rowlen, half, d_out);
  QMG_LAUNCH_CHECK();
  QMG_CUDA(cudaMemcpyAsync(host_out, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, r.stream));
  QMG_CUDA(cudaStreamSynchronize(r.stream));
  return 0;
}

// ------------------------------------------------------------------ gaussian --
extern "C" int qmg_gaussian(qmg_cplx* x_, long n, uint64_t seed, uint64_t stream_id, double dev)
{
  QMG_REQUIRE_INIT();
  cd* x = CD(x_);
  // y-slab sharding: element i of the local even-odd field is element  p * (N half) + rank * half + i'  of the global one
  // (i = p * half + i'), so the ranks together draw exactly the vector a single GPU would draw for the whole lattice.
  const long half = n / 2;
  const long nranks = qmg_comm_size(), rank = qmg_comm_rank();
  const bool split = nranks > 1 && (n % 2 == 0);
  return launch_ew(n, [=] __device__(long i) {
    long gi = i;
    if (split) { const long p = i >= half ? 1 : 0; gi = p * nranks * half + rank * half + (i - p * half); }
    double n0, n1;
    philox_normal2(seed, (uint64_t)gi, stream_id, n0, n1);
    x[i] = cmake(dev * n0, dev * n1);
  });
}
