// Shared device/host helpers for libqmg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/qmg_b200.h"

namespace qmg {

// ---- global runtime state (one device / one driving host thread per process) ----
struct Runtime
{
  bool ready = false;
  int device = -1;
  int sm_count = 148;
  cudaStream_t stream = 0;
  long launches = 0;
  int profile = 0;                  // per-entry-point timing enabled
  int managed = 0;                  // qmg_malloc hands out managed memory (reference drivers index vectors on the host)
  // reduction scratch: per-block partials + a completion counter, and a pinned result slot
  double* d_partials = nullptr;     // at least kMaxRedBlocks * kMaxRedWidth doubles (grown on demand)
  size_t partials_cap = 0;          // capacity of d_partials in doubles
  unsigned int* d_counter = nullptr;
  double* d_result = nullptr;       // kMaxRedWidth doubles
  double* h_result = nullptr;       // pinned
  void** d_ptrs = nullptr;          // small device table for pointer arrays (multi-dot etc.)
  double* d_scalars = nullptr;      // small device table for coefficient arrays
  std::string error;
};
Runtime& rt();

constexpr int kMaxRedBlocks = 1184;   // 148 SMs * 8
constexpr int kMaxRedWidth = 130;     // doubles per block partial (64 complex + 2)
constexpr int kMaxPtrs = 256;

int fail(const char* what, cudaError_t e, const char* file, int line);
int fail_msg(const char* msg);

#define QMG_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return qmg::fail(#call, e__, __FILE__, __LINE__); } while (0)
// Optional per-entry-point timing (QMG_PROFILE=1 or qmg_profile_enable(1)): every C-ABI call is bracketed by stream
// synchronisations and its wall time is accumulated under the function's name; qmg_profile_report() prints the table.
// Off by default: the scope object is then a single predictable branch.
struct ProfScope
{
  const char* name; double t0; bool on;
  explicit ProfScope(const char* n);
  ~ProfScope();
};
#define QMG_REQUIRE_INIT() if (!qmg::rt().ready) { int r__ = qmg_init(-1); if (r__) return r__; } qmg::ProfScope prof_scope__(__func__)
#define QMG_LAUNCH_CHECK() do { qmg::rt().launches++; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return qmg::fail("kernel launch", e__, __FILE__, __LINE__); } while (0)

// ---- complex<double> as double2 -------------------------------------------------
typedef double2 cd;
__host__ __device__ __forceinline__ cd cmake(double r, double i) { cd z; z.x = r; z.y = i; return z; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cconj(cd a) { return cmake(a.x, -a.y); }
// acc += a*b
__device__ __forceinline__ void cfma(cd& acc, cd a, cd b)
{
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void cfma_conj(cd& acc, cd a, cd b)
{
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ cd cdiv(cd a, cd b)
{
  double d = b.x * b.x + b.y * b.y;
  return cmake((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// streaming (read-once) 16-byte load: matrices never get reused inside one apply
__device__ __forceinline__ cd ld_stream(const cd* p) { return __ldcs(p); }
// read-only cached load: vectors are re-read by the neighbouring sites
__device__ __forceinline__ cd ld_keep(const cd* p) { return __ldg(p); }

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cd shfl_xor_c(cd v, int m) { return cmake(shfl_xor_d(v.x, m), shfl_xor_d(v.y, m)); }
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
  return v;
}

// Block-level sum of W doubles per thread; the result is valid in thread 0.
// smem must hold (blockDim.x/32) * W doubles.
template <int W>
__device__ __forceinline__ void block_sum(double (&v)[W], double* smem)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int w = 0; w < W; w++) v[w] = warp_sum(v[w]);
  if (lane == 0)
  {
#pragma unroll
    for (int w = 0; w < W; w++) smem[warp * W + w] = v[w];
  }
  __syncthreads();
  if (warp == 0)
  {
#pragma unroll
    for (int w = 0; w < W; w++)
    {
      double t = (lane < nwarp) ? smem[lane * W + w] : 0.0;
      v[w] = warp_sum(t);
    }
  }
  __syncthreads();
}

// Deterministic grid reduction tail: every block stores its W partials; the
// last block to arrive sums them in block order and writes result[0..W).
template <int W>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[W], double* smem, double* partials, unsigned int* counter, double* result)
{
  block_sum<W>(v, smem);
  __shared__ bool is_last;
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int w = 0; w < W; w++) partials[(size_t)blockIdx.x * W + w] = v[w];
    __threadfence();
    unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last)
  {
    __threadfence();
    double acc[W];
#pragma unroll
    for (int w = 0; w < W; w++) acc[w] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
    {
#pragma unroll
      for (int w = 0; w < W; w++) acc[w] += __ldcg(&partials[(size_t)b * W + w]);
    }
    block_sum<W>(acc, smem);
    if (threadIdx.x == 0)
    {
#pragma unroll
      for (int w = 0; w < W; w++) result[w] = acc[w];
      *counter = 0u;
    }
  }
}

// fetch `count` doubles of the reduction result to the host (synchronises the stream)
int fetch_result(double* host_out, int count, int op_max = 0);   // all-reduced over the ranks when sharded
// make sure the per-block partial scratch holds `ndoubles`; returns nullptr on failure
double* ensure_partials(size_t ndoubles);
int reduction_grid(long n_items, int items_per_block);

} // namespace qmg
