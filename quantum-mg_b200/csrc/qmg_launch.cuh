// Launch helpers shared by the streaming kernels: grid-stride element-wise and reduction launches (lambda bodies),
// and the counter-based Philox generator.
#pragma once
#include "qmg_common.cuh"

namespace qmg {

constexpr int kEwBlock = 256;

// <a|b> and |a|^2 accumulated with pinned roundings, so that every kernel that forms these sums over the same
// element -> thread assignment produces the same bits (dot_norm, the Krylov step kernels, the fused GCR orthogonalisation)
__device__ __forceinline__ void dot_acc3(double (&acc)[3], const cd a, const cd b)
{
  acc[0] = __dadd_rn(acc[0], __fma_rn(a.y, b.y, __dmul_rn(a.x, b.x)));
  acc[1] = __dadd_rn(acc[1], __fma_rn(-a.y, b.x, __dmul_rn(a.x, b.y)));
  acc[2] = __dadd_rn(acc[2], __fma_rn(a.y, a.y, __dmul_rn(a.x, a.x)));
}

template <class F>
__global__ void __launch_bounds__(kEwBlock) ew_kernel(long n, F f)
{
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

template <int W, class F>
__global__ void __launch_bounds__(kEwBlock) reduce_kernel(long n, F f, double* partials, unsigned int* counter, double* result)
{
  __shared__ double smem[(kEwBlock / 32) * W];
  double acc[W];
#pragma unroll
  for (int w = 0; w < W; w++) acc[w] = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i, acc);
  grid_reduce_finish<W>(acc, smem, partials, counter, result);
}

static inline int ew_grid(long n)
{
  long want = (n + kEwBlock - 1) / kEwBlock;
  long cap = (long)rt().sm_count * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <class F> static inline int launch_ew(long n, F f)
{
  if (n <= 0) return 0;
  ew_kernel<<<ew_grid(n), kEwBlock, 0, rt().stream>>>(n, f);
  QMG_LAUNCH_CHECK();
  return 0;
}

// Grid of a reducing kernel: never more blocks than are resident at once.  The wide reductions (multi-dot: 16 accumulators,
// 76 registers) fit 3 blocks per SM, not 8; 1184 blocks were then 2.67 waves with a third of the GPU idle during the last one.
template <class K> static inline int resident_grid(K kernel, long n)
{
  static int per_sm[64] = { 0 };
  const int dev = rt().device & 63;
  if (per_sm[dev] == 0)
  {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kEwBlock, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
    per_sm[dev] = nb;
  }
  int grid = ew_grid(n > 0 ? n : 1);
  const long cap = (long)per_sm[dev] * rt().sm_count;
  if (grid > cap) grid = (int)cap;
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  return grid;
}

template <int W, class F> static inline int launch_reduce(long n, F f, double* host_out)
{
  Runtime& r = rt();
  const int grid = resident_grid(reduce_kernel<W, F>, n);
  reduce_kernel<W><<<grid, kEwBlock, 0, r.stream>>>(n, f, r.d_partials, r.d_counter, r.d_result);
  QMG_LAUNCH_CHECK();
  double sink[W];
  return fetch_result(host_out ? host_out : sink, W);     // every reduction is fetched: host and device count them alike
}


__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
  for (int round = 0; round < 10; round++)
  {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}


// one standard normal pair from one Philox block keyed by (seed, counter): Box-Muller on two 32-bit uniforms
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, double& n0, double& n1)
{
  uint32_t c[4] = { (uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32) };
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u1 = ((double)c[0] + 0.5) * (1.0 / 4294967296.0);
  const double u2 = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, co;
  sincospi(2.0 * u2, &s, &co);
  n0 = rad * co; n1 = rad * s;
}

// A reduction whose result stays on the device (all-reduced over the ranks like any other) for the NEXT kernel to read:
// nothing is fetched, the host only keeps its reduction count in step.  result_dev: W doubles of device memory.
template <int W, class F> static inline int launch_reduce_keep(long n, F f, double* result_dev)
{
  Runtime& r = rt();
  const int grid = resident_grid(reduce_kernel<W, F>, n);
  reduce_kernel<W><<<grid, kEwBlock, 0, r.stream>>>(n, f, r.d_partials, r.d_counter, result_dev);
  QMG_LAUNCH_CHECK();
  return skip_result(result_dev, W);
}

} // namespace qmg
