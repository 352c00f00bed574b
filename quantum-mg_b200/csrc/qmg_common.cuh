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
  unsigned int* d_counter = nullptr; // first word of the RedState block (see below)
  double* d_result = nullptr;       // kMaxRedWidth doubles
  double* h_result = nullptr;       // pinned, mapped: the last block of every reduction writes its result here
  unsigned long long* h_flag = nullptr;   // pinned, mapped: number of reductions published so far
  unsigned long long* h_err = nullptr;    // pinned, mapped: non-zero once a kernel gave up waiting for a peer (1 + the peer's rank; +256 for a halo row)
  unsigned long long red_seq = 0;   // host mirror of that number (reductions launched)
  int publish = 1;                  // 1: results arrive through h_result / h_flag; 0 (QMG_PUBLISH=0): copy + stream synchronise
  int bicgstab_fused = 1;           // BiCGstab(L) sweeps through the fused kernels (QMG_BICGSTAB_FUSED=0: call by call)
  int tile_kernel = 1;              // gamma5-hermitian applies use the shared-memory tile kernel (QMG_TILE=0: streaming HERM kernel)
  int publish_now = 1;              // publish, and the kernels hold the final (all-reduced) values themselves
  void** d_ptrs = nullptr;          // small device table for pointer arrays (multi-dot etc.)
  double* d_scalars = nullptr;      // small device table for coefficient arrays
  std::string error;
};
Runtime& rt();

constexpr int kMaxRedBlocks = 1184;   // 148 SMs * 8
constexpr int kMaxRedWidth = 130;     // doubles per block partial (64 complex + 2)
constexpr int kMaxPtrs = 256;
constexpr int kMaxRanks = 8;          // GPUs of one NVSwitch node
constexpr int kMailWidth = 2 * kMaxPtrs;   // doubles one rank can publish per reduction

// Device-resident state of the reduction tail, shared by every reducing kernel through its `counter` argument (the block
// starts with the completion counter).  The LAST block of a reduction, after summing the per-block partials:
//   1. sharded with peer mailboxes (p2p): stores its values into slot [seq & 1][rank] of EVERY rank's mailbox with plain
//      NVLink peer stores, raises its flag there, waits until all nranks flags of its own mailbox show this sequence
//      number and sums the nranks contributions in rank order -- the all-reduce happens inside the reducing kernel, every
//      rank obtains bit-identical values, and no collective kernel is launched;
//   2. writes the final values and the sequence number into mapped pinned host memory, where the host thread is
//      polling: no device-to-host copy, no stream synchronisation.
// Two slots are enough: a rank can only start reduction n+2 after every rank has raised its flag for n+1, i.e. after
// every rank has finished reading slot n.
struct RedState
{
  unsigned int counter; unsigned int pad;
  unsigned long long seq;           // reductions published by THIS process so far (host handshake; private to the process)
  unsigned long long coll_seq;      // reductions all-reduced through the mailboxes since qmg_comm_init: the SAME on every rank, it
                                    // picks the mailbox slot and is the flag value (ranks may have reduced different numbers of
                                    // times before they joined the communicator)
  int nranks, rank, p2p, publish;
  double* mail[kMaxRanks];          // mail[r]: rank r's mailbox (own: local pointer; others: IPC-mapped peer memory)
  double* host_out;
  unsigned long long* host_flag;
  unsigned long long* host_err;
  long long watchdog_cycles;        // how long a last block waits for its peers before it gives up (QMG_P2P_TIMEOUT_S, default 120 s)
};
// mailbox layout: data[2][kMaxRanks][kMailWidth] doubles, then flags[2][kMaxRanks] (unsigned long long)
constexpr size_t kMailDataDoubles = (size_t)2 * kMaxRanks * kMailWidth;
constexpr size_t kMailBytes = sizeof(double) * kMailDataDoubles + sizeof(unsigned long long) * 2 * kMaxRanks;
// The same IPC block continues with the halo mailbox: flags[2][2] (sequence numbers; [slot][0] = row -1 from the lower
// neighbour, [slot][1] = row Y from the upper neighbour), then data[2][2][kHaloCap] complex.
constexpr size_t kHaloOffset = (kMailBytes + 255) & ~(size_t)255;
constexpr size_t kHaloCap = (size_t)1 << 18;          // complex elements per row slot (4 MB): X * dof of one boundary row
constexpr size_t kHaloDataOffset = kHaloOffset + 256;
constexpr size_t kPeerBlockBytes = kHaloDataOffset + sizeof(double) * 2 * 4 * kHaloCap;

int fail(const char* what, cudaError_t e, const char* file, int line);
int fail_msg(const char* msg);

#define QMG_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return qmg::fail(#call, e__, __FILE__, __LINE__); } while (0)
// Optional per-entry-point timing (QMG_PROFILE=1 or qmg_profile_enable(1)): every C-ABI call is bracketed by stream
// synchronisations and its wall time is accumulated under the function's name; qmg_profile_report() prints the table.
// Off by default: the scope object is then a single predictable branch.
struct ProfScope
{
  const char* name; double t0; bool on;
  char tag[48];                      // optional detail appended to the name (e.g. the lattice an apply ran on)
  explicit ProfScope(const char* n);
  ~ProfScope();
  void detail(int nc, int X, int Y);
};
#define QMG_REQUIRE_INIT() if (!qmg::rt().ready) { int r__ = qmg_init(-1); if (r__) return r__; } qmg::ProfScope prof_scope__(__func__)
#define QMG_LAUNCH_CHECK() do { qmg::rt().launches++; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return qmg::fail("kernel launch", e__, __FILE__, __LINE__); } while (0)

// ---- complex<double> as double2 -------------------------------------------------
typedef double2 cd;
__host__ __device__ __forceinline__ cd cmake(double r, double i) { cd z; z.x = r; z.y = i; return z; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmake(a.x - b.x, a.y - b.y); }
// a*b with its roundings spelt out (one product rounded, the other fused -- what nvcc's contraction made of the plain expression),
// so that every kernel that multiplies the same numbers gets the same bits whatever surrounds the call
__device__ __forceinline__ cd cmul(cd a, cd b) { return cmake(fma(a.x, b.x, -__dmul_rn(a.y, b.y)), fma(a.y, b.x, __dmul_rn(a.x, b.y))); }
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

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p)
{ unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v)
{ asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ double ld_sys_f64(const double* p)
{ double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_sys_f64(double* p, double v)
{ asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }

// Called by ALL threads of the last block once result[0..W) holds this rank's sums (written by one of its threads and
// followed by __syncthreads()).  tid / nthreads: flattened thread index and count of the block.
__device__ __forceinline__ void publish_result(unsigned int* counter, double* result, int W, int op_max, int tid, int nthreads)
{
  RedState* st = reinterpret_cast<RedState*>(counter);
  const unsigned long long seq = st->seq + 1;
  if (st->p2p)
  {
    const unsigned long long cseq = st->coll_seq + 1;
    const int nr = st->nranks, me = st->rank, buf = (int)(cseq & 1);
    for (int i = tid; i < W * nr; i += nthreads)
    {
      const int r = i / W, w = i - r * W;
      st_sys_f64(st->mail[r] + ((size_t)buf * kMaxRanks + me) * kMailWidth + w, result[w]);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < nr)
    {
      unsigned long long* peer_flags = reinterpret_cast<unsigned long long*>(st->mail[tid] + kMailDataDoubles);
      st_sys_u64(peer_flags + buf * kMaxRanks + me, cseq);
      const unsigned long long* my_flags = reinterpret_cast<const unsigned long long*>(st->mail[me] + kMailDataDoubles);
      const long long t0 = clock64();
      while (ld_sys_u64(my_flags + buf * kMaxRanks + tid) < cseq)
        if (clock64() - t0 > st->watchdog_cycles)
        {
          // a peer is gone: no trap (that would poison the context) -- flag the error for the host, which turns it into an
          // error code at the next fetch, and let the kernel retire with whatever arrived
          st_sys_u64(st->host_err, 1ull + (unsigned long long)tid);
          break;
        }
    }
    __syncthreads();
    for (int w = tid; w < W; w += nthreads)
    {
      const double* slot = st->mail[me] + (size_t)buf * kMaxRanks * kMailWidth + w;
      double acc = ld_sys_f64(slot);
      for (int r = 1; r < nr; r++) { const double v = ld_sys_f64(slot + (size_t)r * kMailWidth); acc = op_max ? fmax(acc, v) : acc + v; }
      result[w] = acc;
    }
    __syncthreads();
    if (tid == 0) st->coll_seq = cseq;
  }
  if (st->publish)
  {
    for (int w = tid; w < W; w += nthreads) st->host_out[w] = result[w];
    __threadfence_system();
    __syncthreads();
  }
  if (tid == 0)
  {
    st->seq = seq;
    st->counter = 0u;
    if (st->publish) st_sys_u64(st->host_flag, seq);
  }
}

// Deterministic grid reduction tail: every block stores its W partials; the
// last block to arrive sums them in block order and writes result[0..W).
template <int W>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[W], double* smem, double* partials, unsigned int* counter, double* result)
{
  static_assert(W <= kMailWidth && W <= kMaxRedWidth, "reduction wider than the result slots");
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
    }
    __syncthreads();
    publish_result(counter, result, W, 0, threadIdx.x, blockDim.x);
  }
}

// fetch `count` doubles of the reduction result to the host (synchronises the stream)
int fetch_result(double* host_out, int count, int op_max = 0);   // all-reduced over the ranks when sharded
// count a reduction whose result is consumed on the device (NCCL fallback: all-reduce it in place, stream-ordered)
int skip_result(double* result_dev, int count);
// make sure the per-block partial scratch holds `ndoubles`; returns nullptr on failure
double* ensure_partials(size_t ndoubles);
int reduction_grid(long n_items, int items_per_block);

} // namespace qmg
