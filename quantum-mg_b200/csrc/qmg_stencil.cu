// K1: the stencil apply -- THE hot kernel.
//
//   out(x) (=|+=) [clover(x) + diag shifts] in(x) + sum_mu hopping_mu(x) in(x+mu)
//
// replaces Stencil2D::apply_M and its pieces
// (/root/reference/stencil/stencil_2d.h:666-936): one fused pass instead of
// 1 zero + 1 clover cMATxpy + 8 cshift + 8 half-volume cMATxpy + 2 caxpy.
//
// Mapping: one lane per stored matrix ELEMENT.  Lane (site, c1, c2) streams
// clover[c1][c2] and the four hopping_mu[c1][c2] of its site -- consecutive lanes
// read consecutive 16-byte elements, so every warp-level load covers one
// contiguous 512-byte span (4 lines) and each link is fetched from HBM exactly
// once -- multiplies by in(neighbour)[c2] (re-used across c1 and across
// neighbouring sites through L1/L2) and the nc partial products of a row are
// summed with a log2(nc)-step xor-shuffle tree.  The c2 == 0 lane writes
// out[site][c1].  No shared memory, no divergence inside a row.
//
// HBM-bound: 16*(nc^2*(n_clover+4) + 2nc) algorithmic bytes per site
// (Wilson 384 B, coarse nc=8 5376 B) for 8*nc^2*(n_clover+4) flops.
#include "qmg_comm.cuh"
#include <stdlib.h>

namespace qmg {

struct StencilKArgs
{
  const cd* clover;     // nullptr if absent / not applied
  const cd* hop;        // nullptr if absent / not applied
  const cd* in;
  cd* out;
  const cd* dotw;       // fused <out|dotw> or nullptr
  const cd* resid;      // residual epilogue: out = resid - A in (nullptr: out = A in)
  const cd* halo_ym;    // input row y=-1 from the lower rank, or nullptr (periodic)
  const cd* halo_yp;    // input row y=Y from the upper rank, or nullptr
  Geom g;
  long size_cm;         // V*nc*nc: stride between hopping directions
  cd diag[2][2];        // [parity][dof half]: shift +- eo_shift +- dof_shift (+1 for identity clover)
  int use_diag;
  int hop_to[2];        // hop_to[p]: apply hopping into parity p
  int dir_mask;
  int accumulate;
  int p_begin;          // first parity written
  int n_par;            // parities written; the parity is the FASTEST block index so that the even and the odd
                        // output rows y are in flight together and each input row is fetched from HBM once
  int herm;             // gamma5-hermitian link set: backward hops are read from the neighbours' forward blocks
  const cd* hop_ym;     // herm + y-slab: row -1 of the +y hopping blocks (from the lower rank), layout (parity, x/2, nc nc)
  int y_off, y_stride, y_cnt;   // rows of this launch: y = y_off + i * y_stride, i < y_cnt (all rows: 0, 1, Y; the two
                                // slab-boundary rows of a sharded apply: 0, Y-1, 2; its interior: 1, 1, Y-2)
  // matrix-free Wilson apply (nc = 2, the whole operator): the stored blocks ARE the Wilson blocks of this gauge field
  const cd* mf_gauge;           // [mu V + site], mu = x, y, on the nc = 1 lattice; nullptr: read the stored blocks
  const cd* mf_gauge_ym;        // y-slab: row -1 of U_y (layout (parity, x/2)), from the lower rank
  double mf_w;                  // Wilson parameter
};

// HERM: for an operator with D^dag = gamma5 D gamma5 (Wilson and its Galerkin coarsenings with chirality-preserving
// transfers) the backward block is determined by the forward block of the neighbour,
//   H_{-mu}(x)[a][b] = s_a s_b conj( H_{+mu}(x - mu)[b][a] ),   s = +1 on the top half of the dof, -1 on the bottom half,
// so only clover, H_{+x}, H_{+y} are fetched from HBM (3 of the 5 blocks); the neighbour's forward block was read by the
// neighbouring site moments ago (same row for -x, previous row for -y) and comes out of L2.  The lane that owns element
// (c1, c2) reads element (c2, c1) of that block: the same 16 nc^2 bytes per site, permuted inside the block.
// RESID: the residual epilogue out = resid - A in.  A template parameter, not a run-time test: the plain instantiation must stay
// the straight-line body ptxas schedules with every load ahead of the first FMA (a run-time `if (a.resid)` in the epilogue made
// it interleave load, use, load, use: 3.61 -> 4.07 ms on the 8192^2 Wilson apply, profiles/r03h_ncu_stencil_wilson8192_regression.txt)
template <int NC, bool REDUCE, bool HERM, bool RESID>
__global__ void __launch_bounds__(256) stencil_kernel(const StencilKArgs a, double* partials, unsigned int* counter, double* result)
{
  constexpr int LPS = NC * NC;   // lanes per site
  const int bxi = (a.n_par == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int col = bxi * blockDim.x + threadIdx.x;   // element inside the row
  const int yi = blockIdx.y * blockDim.y + threadIdx.y;
  const int y = a.y_off + yi * a.y_stride;
  const int p = a.p_begin + ((a.n_par == 2) ? (blockIdx.x & 1) : 0);
  const int k = col / LPS;
  const int c = col % LPS;
  const int c1 = c / NC, c2 = c % NC;
  const bool active = (k < a.g.xh) && (yi < a.y_cnt);

  // Straight-line body: every address is formed first, then all (up to 11) 16-byte loads are issued back to back as
  // predicated loads with no branch in between, so one thread keeps ~176 B in flight and a full SM ~350 KB -- the
  // latency-bandwidth product of HBM3e needs ~45 KB per SM.  (A branch per direction serialises the round trips.)
  const cd zero = cmake(0.0, 0.0);
  const unsigned h = active ? (unsigned)y * a.g.xh + k : 0u;
  const size_t site = (size_t)p * a.g.half + h;
  const int q = 1 - p;
  const bool hop = active && (a.hop != nullptr) && (a.hop_to[p] != 0);
  const bool m0 = hop && (a.dir_mask & 1), m1 = hop && (a.dir_mask & 2), m2 = hop && (a.dir_mask & 4), m3 = hop && (a.dir_mask & 8);
  const bool has_cl = active && (a.clover != nullptr);
  const bool has_dg = active && a.use_diag && (c1 == c2);
  const bool writer = active && (c2 == 0);
  const size_t idx = site * NC + c1;

  const int sft = (y + p) & 1;
  int kxp = k + sft; kxp = (kxp == a.g.xh) ? 0 : kxp;
  int kxm = k - 1 + sft; kxm = (kxm < 0) ? a.g.xh - 1 : kxm;
  const int yp1 = (y + 1 == a.g.Y) ? 0 : y + 1;
  const int ym1 = (y == 0) ? a.g.Y - 1 : y - 1;
  const cd* in_q = a.in + (size_t)q * a.g.half * NC + c2;
  const cd* s0 = in_q + ((size_t)y * a.g.xh + kxp) * NC;
  const cd* s2 = in_q + ((size_t)y * a.g.xh + kxm) * NC;
  const cd* s1 = (a.halo_yp != nullptr && y == a.g.Y - 1) ? a.halo_yp + ((size_t)q * a.g.xh + k) * NC + c2 : in_q + ((size_t)yp1 * a.g.xh + k) * NC;
  const cd* s3 = (a.halo_ym != nullptr && y == 0) ? a.halo_ym + ((size_t)q * a.g.xh + k) * NC + c2 : in_q + ((size_t)ym1 * a.g.xh + k) * NC;
  const cd* hp = a.hop + site * LPS + c;

  cd H0, H1, H2, H3;
  if (HERM)
  {
    // forward blocks stay cacheable (the neighbour in -mu direction re-reads them); the re-read is their last use
    const int ct = c2 * NC + c1;
    const cd* hb2 = a.hop + ((size_t)q * a.g.half + (size_t)y * a.g.xh + kxm) * LPS + ct;
    const cd* hb3 = (a.hop_ym != nullptr && y == 0) ? a.hop_ym + ((size_t)q * a.g.xh + k) * LPS + ct
                                                    : a.hop + a.size_cm + ((size_t)q * a.g.half + (size_t)ym1 * a.g.xh + k) * LPS + ct;
    H0 = m0 ? ld_keep(hp) : zero;
    H1 = m1 ? ld_keep(hp + a.size_cm) : zero;
    const cd B2 = m2 ? ld_stream(hb2) : zero;
    const cd B3 = m3 ? ld_stream(hb3) : zero;
    const double sg = ((2 * c1 < NC) == (2 * c2 < NC)) ? 1.0 : -1.0;
    H2 = cmake(sg * B2.x, -sg * B2.y);
    H3 = cmake(sg * B3.x, -sg * B3.y);
  }
  else
  {
    H0 = m0 ? ld_stream(hp) : zero;
    H1 = m1 ? ld_stream(hp + a.size_cm) : zero;
    H2 = m2 ? ld_stream(hp + 2 * a.size_cm) : zero;
    H3 = m3 ? ld_stream(hp + 3 * a.size_cm) : zero;
  }
  const cd CL = has_cl ? ld_stream(a.clover + site * LPS + c) : zero;
  const cd V0 = m0 ? ld_keep(s0) : zero;
  const cd V1 = m1 ? ld_keep(s1) : zero;
  const cd V2 = m2 ? ld_keep(s2) : zero;
  const cd V3 = m3 ? ld_keep(s3) : zero;
  const cd VC = (has_cl || has_dg) ? ld_keep(a.in + site * NC + c2) : zero;
  const cd OLD = (writer && a.accumulate) ? a.out[idx] : zero;
  const cd RB = (RESID && writer) ? ld_stream(a.resid + idx) : zero;
  const cd DG = has_dg ? a.diag[p][(2 * c2 >= NC && NC > 1) ? 1 : 0] : zero;

  cd acc = zero;
  cfma(acc, CL, VC);
  cfma(acc, DG, VC);
  cfma(acc, H0, V0);
  cfma(acc, H1, V1);
  cfma(acc, H2, V2);
  cfma(acc, H3, V3);

  // sum the nc partial products of each matrix row (lanes c2 = 0..NC-1 are consecutive)
#pragma unroll
  for (int o = NC / 2; o > 0; o >>= 1) { cd t = shfl_xor_c(acc, o); acc = cadd(acc, t); }

  double red[3] = {0.0, 0.0, 0.0};
  if (writer)
  {
    acc = cadd(acc, OLD);
    if (RESID) acc = csub(RB, acc);
    a.out[idx] = acc;
    if (REDUCE)
    {
      red[2] = acc.x * acc.x + acc.y * acc.y;
      if (a.dotw != nullptr)
      {
        const cd w = ld_keep(a.dotw + idx);
        red[0] = acc.x * w.x + acc.y * w.y;
        red[1] = acc.x * w.y - acc.y * w.x;
      }
    }
  }
  if (REDUCE)
  {
    // flatten the block for the shared reduction helpers
    __shared__ double smem[8 * 3];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int lane = tid & 31, warp = tid >> 5, nwarp = (nthreads + 31) >> 5;
#pragma unroll
    for (int w = 0; w < 3; w++) red[w] = warp_sum(red[w]);
    if (lane == 0) { smem[warp * 3 + 0] = red[0]; smem[warp * 3 + 1] = red[1]; smem[warp * 3 + 2] = red[2]; }
    __syncthreads();
    __shared__ bool is_last;
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (warp == 0)
    {
#pragma unroll
      for (int w = 0; w < 3; w++) { double t = (lane < nwarp) ? smem[lane * 3 + w] : 0.0; red[w] = warp_sum(t); }
      if (lane == 0)
      {
        partials[(size_t)bid * 3 + 0] = red[0]; partials[(size_t)bid * 3 + 1] = red[1]; partials[(size_t)bid * 3 + 2] = red[2];
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == nblocks - 1);
      }
    }
    __syncthreads();
    if (is_last)
    {
      __threadfence();
      double accr[3] = {0.0, 0.0, 0.0};
      for (unsigned int b = tid; b < nblocks; b += nthreads)
      {
        accr[0] += __ldcg(&partials[(size_t)b * 3 + 0]);
        accr[1] += __ldcg(&partials[(size_t)b * 3 + 1]);
        accr[2] += __ldcg(&partials[(size_t)b * 3 + 2]);
      }
#pragma unroll
      for (int w = 0; w < 3; w++) accr[w] = warp_sum(accr[w]);
      __syncthreads();
      if (lane == 0) { smem[warp * 3 + 0] = accr[0]; smem[warp * 3 + 1] = accr[1]; smem[warp * 3 + 2] = accr[2]; }
      __syncthreads();
      if (warp == 0)
      {
#pragma unroll
        for (int w = 0; w < 3; w++) { double t = (lane < nwarp) ? smem[lane * 3 + w] : 0.0; accr[w] = warp_sum(t); }
        if (lane == 0) { result[0] = accr[0]; result[1] = accr[1]; result[2] = accr[2]; }
      }
      __syncthreads();
      publish_result(counter, result, 3, 0, tid, nthreads);
    }
  }
}

// ---- tile kernel for gamma5-hermitian link sets --------------------------------------------------------------------
// The streaming kernel above sits on the L2 -> SM fabric (every one of the 5 blocks of a site crosses it once, and with
// HERM the backward blocks cross it a second time out of L2).  Here a CTA owns a patch of TY rows x 2 TK columns, stages
// the FORWARD blocks (+x, +y) of its sites -- plus the one column to the left and the one row below that its backward hops
// need -- and the input spinors of the patch and its 1-site ring in shared memory with cp.async, and every lane then
// takes its forward elements [c1][c2] and its backward elements s s conj([c2][c1]) from shared memory: per site
// 1 (clover) + 2 (1 + halo share) blocks cross the fabric instead of 5, and every spinor once per CTA instead of once
// per lane group.  Block rows are padded to NC + 1 elements so that both the row-wise and the transposed reads of a
// quarter warp fall into distinct banks.  Same lane <-> matrix element mapping and the same shuffle tree as above, so the
// sums are formed in the same order.
// SPLIT threads share one (site, column): each walks NC / SPLIT of the rows (more warps per staged byte)
template <int NC, int TK, int TY, int SPLIT = 1> struct TileDims
{
  static const int S = TY * 2 * TK;                               // sites of the patch
  static const int CAP = 256 * SPLIT;
  static const int THREADS = (S * NC * SPLIT >= CAP) ? CAP : S * NC * SPLIT;
  static const int PASSES = S * NC * SPLIT / THREADS;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem_src) : "memory");
}

// the three pieces of a patch's life: stage (asynchronous), fetch the clover column (registers), compute
template <int NC, int TK, int TY, int SPLIT = 1> struct Tile
{
  static const int S = TileDims<NC, TK, TY, SPLIT>::S, NT = TileDims<NC, TK, TY, SPLIT>::THREADS, PASSES = TileDims<NC, TK, TY, SPLIT>::PASSES;
  static const int R = NC / SPLIT;             // rows per thread
  static const int LPS = NC * NC;
  static const int RS = NC + 1;                // padded row stride of a staged block
  static const int BS = NC * RS;               // elements per staged block
  static const int NHX = S + TY, NHY = S + 2 * TK, VK = TK + 2;
  static const int BUF = (NHX + NHY) * BS + (TY + 2) * 2 * VK * NC;     // complex elements of one staging buffer

  __device__ static __forceinline__ void stage(const StencilKArgs& a, cd* buf, int k0, int y0, int tid)
  {
    cd* sHx = buf; cd* sHy = sHx + (size_t)NHX * BS; cd* sV = sHy + (size_t)NHY * BS;
    const int xh = a.g.xh, Y = a.g.Y;
    const size_t half = a.g.half;
    const cd* hopx = a.hop;
    const cd* hopy = a.hop + a.size_cm;
    for (int e = tid; e < NHX * LPS; e += NT)
    {
      const int slot = e / LPS, c = e - slot * LPS;
      int ty, p, k;
      if (slot < S) { ty = slot / (2 * TK); const int r = slot - ty * 2 * TK; p = r / TK; k = k0 + (r - p * TK); }
      else { ty = slot - S; p = 1 - ((y0 + ty) & 1); k = (k0 == 0) ? xh - 1 : k0 - 1; }      // left neighbour of the sft = 0 site of this row
      const size_t site = (size_t)p * half + (size_t)(y0 + ty) * xh + k;
      cp_async16(sHx + (size_t)slot * BS + (c / NC) * RS + (c % NC), hopx + site * LPS + c);
    }
    for (int e = tid; e < NHY * LPS; e += NT)
    {
      const int slot = e / LPS, c = e - slot * LPS;
      int y, p, k;
      if (slot < S) { const int ty = slot / (2 * TK); const int r = slot - ty * 2 * TK; p = r / TK; k = k0 + (r - p * TK); y = y0 + ty; }
      else { const int r = slot - S; p = r / TK; k = k0 + (r - p * TK); y = (y0 == 0) ? Y - 1 : y0 - 1; }
      const size_t site = (size_t)p * half + (size_t)y * xh + k;
      cp_async16(sHy + (size_t)slot * BS + (c / NC) * RS + (c % NC), hopy + site * LPS + c);
    }
    for (int e = tid; e < (TY + 2) * 2 * VK * NC; e += NT)
    {
      const int c = e % NC; int r = e / NC;
      const int kk = r % VK; r /= VK;
      const int p = r & 1, ry = r >> 1;
      int y = y0 + ry - 1; y = (y < 0) ? Y - 1 : ((y >= Y) ? 0 : y);
      int k = k0 + kk - 1; k = (k < 0) ? xh - 1 : ((k >= xh) ? 0 : k);
      cp_async16(sV + e, a.in + ((size_t)p * half + (size_t)y * xh + k) * NC + c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // clover column c2 of this thread's site(s), straight from global memory while a tile is in flight
  // (also the element of the residual epilogue's right-hand side this thread will need at the very end: fetched here, while
  // the tile is in flight, instead of as a dependent load after the butterfly)
  __device__ static __forceinline__ void clover(const StencilKArgs& a, int k0, int y0, int tid, cd (&CLc)[PASSES][R], cd (&RBc)[PASSES])
  {
    const int c2 = tid % NC, hf = (tid / NC) % SPLIT;
    const cd zero = cmake(0.0, 0.0);
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++)
    {
      const int slot = ps * (NT / (NC * SPLIT)) + tid / (NC * SPLIT);
      const int ty = slot / (2 * TK); const int r = slot - ty * 2 * TK; const int p = r / TK, tk = r - p * TK;
      const size_t site = (size_t)p * a.g.half + (size_t)(y0 + ty) * a.g.xh + (k0 + tk);
#pragma unroll
      for (int i = 0; i < R; i++) CLc[ps][i] = (a.clover != nullptr) ? ld_stream(a.clover + site * LPS + (hf * R + i) * NC + c2) : zero;
      RBc[ps] = (a.resid != nullptr && (SPLIT == 1 || (c2 % SPLIT) == 0)) ? ld_stream(a.resid + site * NC + hf * R + c2 / SPLIT) : zero;
    }
  }

  // One thread per (site, column c2).  It holds in(neighbour)[c2] in a register and walks down column c2 of each block,
  // accumulating all NC output rows; the NC threads of a site then exchange partial sums with a transposing butterfly
  // (NC - 1 complex shuffles) that leaves row t on thread t.  Per block a thread reads NC matrix elements and ONE spinor
  // element from shared memory (an element-per-lane mapping reads one of each per element: twice the shared-memory
  // traffic, which is what bounds this kernel).
  __device__ static __forceinline__ void compute(const StencilKArgs& a, const cd* buf, int k0, int y0, int tid, const cd (&CLc)[PASSES][R], const cd (&RBc)[PASSES])
  {
    const cd* sHx = buf; const cd* sHy = sHx + (size_t)NHX * BS; const cd* sV = sHy + (size_t)NHY * BS;
    const int c2 = tid % NC, hf = (tid / NC) % SPLIT;
    const cd zero = cmake(0.0, 0.0);
    const bool top = (2 * c2 < NC);
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++)
    {
      const int slot = ps * (NT / (NC * SPLIT)) + tid / (NC * SPLIT);
      const int ty = slot / (2 * TK); const int r = slot - ty * 2 * TK; const int p = r / TK, tk = r - p * TK;
      const int q = 1 - p, y = y0 + ty, sft = (y + p) & 1;
      const int nx = sft ? (ty * 2 + q) * TK + tk : (tk > 0 ? (ty * 2 + q) * TK + tk - 1 : S + ty);
      const int ny = (ty > 0) ? ((ty - 1) * 2 + q) * TK + tk : S + q * TK + tk;
      const cd* vrow = sV + ((size_t)((ty + 1) * 2) * VK) * NC + c2;       // row ty of the patch, parity 0, kk = 0
      const cd VC = vrow[((size_t)p * VK + tk + 1) * NC];
      const cd V0 = vrow[((size_t)q * VK + tk + 1 + sft) * NC];
      const cd V2 = vrow[((size_t)q * VK + tk + sft) * NC];
      const cd V1 = vrow[((size_t)(2 + q) * VK + tk + 1) * NC];
      const cd V3 = vrow[((ptrdiff_t)q * VK + tk + 1 - 2 * VK) * NC];
      const cd* hx = sHx + (size_t)slot * BS + (hf * R) * RS + c2;   // column c2, rows of this thread: element [c1][c2] at + i * RS
      const cd* hy = sHy + (size_t)slot * BS + (hf * R) * RS + c2;
      const cd* bx = sHx + (size_t)nx * BS + c2 * RS + hf * R;       // row c2 of the neighbour's block: element [c2][c1] at + i
      const cd* by = sHy + (size_t)ny * BS + c2 * RS + hf * R;
      cd acc[R];
#pragma unroll
      for (int i = 0; i < R; i++)
      {
        const int c1 = hf * R + i;
        cd t = zero;
        cfma(t, CLc[ps][i], VC);
        cfma(t, hx[i * RS], V0);
        cfma(t, hy[i * RS], V1);
        // backward blocks: s_{c1} s_{c2} conj(B[c2][c1])
        const double sg = ((2 * c1 < NC) == top) ? 1.0 : -1.0;
        const cd b2 = bx[i], b3 = by[i];
        cfma(t, cmake(sg * b2.x, -sg * b2.y), V2);
        cfma(t, cmake(sg * b3.x, -sg * b3.y), V3);
        acc[i] = t;
      }
      if (a.use_diag)
      {
        // diag shift on row c2 of this thread's column: a predicated add keeps acc[] in registers (no dynamic indexing)
        const cd dg = a.diag[p][top ? 0 : 1];
#pragma unroll
        for (int i = 0; i < R; i++) if (hf * R + i == c2) cfma(acc[i], dg, VC);
      }
      // butterfly over the NC threads (c2) that share this site and row set: transposing while more than one row is
      // left on a thread, a plain pairwise sum afterwards; row hf * R + (c2 / SPLIT) ends up on thread c2
      int left = R;
#pragma unroll
      for (int off = NC / 2; off > 0; off >>= 1)
      {
        if (left > 1)
        {
          const int hl = left / 2;
          const bool upper = (c2 & off) != 0;
#pragma unroll
          for (int i = 0; i < R / 2; i++)
            if (i < hl)
            {
              const cd send = upper ? acc[i] : acc[i + hl];
              const cd keep = upper ? acc[i + hl] : acc[i];
              acc[i] = cadd(keep, shfl_xor_c(send, off));
            }
          left = hl;
        }
        else acc[0] = cadd(acc[0], shfl_xor_c(acc[0], off));
      }
      if (SPLIT == 1 || (c2 % SPLIT) == 0)
      {
        const int row = hf * R + c2 / SPLIT;
        const size_t site = (size_t)p * a.g.half + (size_t)y * a.g.xh + (k0 + tk);
        cd res = acc[0];
        if (a.accumulate) res = cadd(res, a.out[site * NC + row]);
        if (a.resid != nullptr) res = csub(RBc[ps], res);
        a.out[site * NC + row] = res;
      }
    }
  }
};

// one patch per CTA
template <int NC, int TK, int TY, int SPLIT>
__global__ void __launch_bounds__(TileDims<NC, TK, TY, SPLIT>::THREADS, (SPLIT > 1 ? 1024 / TileDims<NC, TK, TY, SPLIT>::THREADS : 1)) stencil_tile_kernel(const StencilKArgs a)
{
  typedef Tile<NC, TK, TY, SPLIT> T;
  extern __shared__ cd tile_smem[];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * TK, y0 = a.y_off + blockIdx.y * TY;
  T::stage(a, tile_smem, k0, y0, tid);
  cd CLc[T::PASSES][T::R];
  cd RBc[T::PASSES];
  T::clover(a, k0, y0, tid, CLc, RBc);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  T::compute(a, tile_smem, k0, y0, tid, CLc, RBc);
}

// (A persistent, double-buffered flavour -- a CTA walking over patches, the next patch's cp.async traffic filling a second
// buffer while it computes -- was measured too: 3.04 ms against 2.74 ms for this one at nc = 8 on 2048^2; with 97-108 KB per
// buffer pair it leaves 8 warps per SM to do the arithmetic.  Two independent one-patch CTAs per SM overlap better.)

template <int NC, int TK, int TY> static size_t tile_smem_bytes()
{
  constexpr int S = TileDims<NC, TK, TY>::S;
  return sizeof(cd) * ((size_t)(S + TY + S + 2 * TK) * NC * (NC + 1) + (size_t)(TY + 2) * 2 * (TK + 2) * NC);
}

// A full (all pieces, both parities, all directions) out-of-place or accumulating apply of a gamma5-hermitian link set on
// a lattice whose dimensions the patch divides ...
template <int TK, int TY> static bool tile_shape_fits(const StencilKArgs& a, int n_par)
{
  return a.herm && n_par == 2 && a.hop != nullptr && a.hop_to[0] && a.hop_to[1] && a.dir_mask == 15 &&
         a.g.xh % TK == 0 && a.g.Y % TY == 0 && (const void*)a.in != (const void*)a.out;
}
// ... over ALL rows of a periodic lattice (rows that wrap read their neighbours inside this lattice)
template <int TK, int TY> static bool tile_applicable(const StencilKArgs& a, int n_par)
{
  return tile_shape_fits<TK, TY>(a, n_par) && a.halo_ym == nullptr && a.halo_yp == nullptr && a.hop_ym == nullptr &&
         a.y_off == 0 && a.y_stride == 1 && a.y_cnt == a.g.Y;
}

template <int NC, int TK, int TY, int SPLIT = 1> static int launch_tile(const StencilKArgs& a)
{
  static bool configured[64] = { false };      // the attribute is per device (qmg_init may move to another one)
  const size_t smem = tile_smem_bytes<NC, TK, TY>();
  const int dev = rt().device & 63;
  if (!configured[dev]) { QMG_CUDA(cudaFuncSetAttribute(stencil_tile_kernel<NC, TK, TY, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured[dev] = true; }
  dim3 grid(a.g.xh / TK, a.y_cnt / TY, 1);       // rows [y_off, y_off + y_cnt): whole patches
  if (grid.y > 65535) return fail_msg("qmg_stencil_apply: Y too large for the launch grid");
  stencil_tile_kernel<NC, TK, TY, SPLIT><<<grid, TileDims<NC, TK, TY, SPLIT>::THREADS, smem, rt().stream>>>(a);
  QMG_LAUNCH_CHECK();
  return 0;
}

// ---- the same patch staged by the TMA engine ---------------------------------------------------------------------------
// The cp.async flavour above spends ~10 LDGSTS plus their index arithmetic per thread on staging (64 % issue utilisation,
// profiles/r03a_ncu_tile_kernel_cp_async.txt).  Here the patch arrives by BULK asynchronous copies -- cp.async.bulk
// global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP / SYNCS) -- issued by the lanes of ONE warp:
// the forward blocks of the TK sites of one (row, parity) are TK nc^2 contiguous complex numbers in the reference's layout
// (4 KB at nc = 8), so a patch is 2 TY + TY copies for +x (rows + left halo column), 2 TY + 2 for +y (rows + the halo row
// below), and the spinor rows with their 1-site ring; no thread computes a staging address in the hot loop and nothing
// passes through registers.  Bulk copies cannot pad, so the blocks sit UNPADDED (row stride nc) and the backward products
// are re-mapped to stay conflict-free: the thread of output row a walks DOWN column a of the neighbour's block --
// consecutive lanes read consecutive elements -- against broadcast spinor elements, and hands its complete row sum to the
// lane that owns row a in the forward butterfly.  Forward products keep the column mapping of the cp.async kernel.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one bulk copy global -> shared of `bytes` (multiple of 16, both addresses 16-byte aligned), completing on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int NC, int TK, int TY, int SPLIT> struct TmaTile
{
  static const int S = TileDims<NC, TK, TY, SPLIT>::S, NT = TileDims<NC, TK, TY, SPLIT>::THREADS, PASSES = TileDims<NC, TK, TY, SPLIT>::PASSES;
  static const int R = NC / SPLIT;
  static const int LPS = NC * NC;              // elements per block, unpadded
  static const int NHX = S + TY, NHY = S + 2 * TK, VK = TK + 2;
  static const int NV = (TY + 2) * 2 * VK * NC;
  static const unsigned BYTES = (unsigned)(sizeof(cd) * ((size_t)(NHX + NHY) * LPS + NV));
  static const size_t SMEM = 128 + BYTES;      // mbarrier in the first 128 bytes
  // copies: +x rows 2 TY, +x halo column TY, +y rows 2 TY, +y halo row 2, spinor (row, parity) pairs 2 (TY + 2)
  static const int NCOPY = 2 * TY + TY + 2 * TY + 2 + 2 * (TY + 2);

  // Issued by ONE warp: lane l takes copies l, l + 32, ...  Two halves with their own completion barriers and their own
  // destination regions, so that the ring kernel can recycle them separately: X = +x blocks (rows, left halo column) and the
  // spinor rows with their ring; Y = +y blocks (rows, halo row below).
  static const unsigned BYTES_Y = (unsigned)(sizeof(cd) * (size_t)NHY * LPS);
  static const unsigned BYTES_X = BYTES - BYTES_Y;
  static const int NCOPY_X = 2 * TY + TY + 2 * (TY + 2), NCOPY_Y = 2 * TY + 2;
  __device__ static __forceinline__ void stage_x(const StencilKArgs& a, cd* sHx, cd* sV, unsigned long long* bar, int k0, int y0, int lane)
  {
    const int xh = a.g.xh, Y = a.g.Y;
    const size_t half = a.g.half;
    const cd* hopx = a.hop;
    for (int c = lane; c < NCOPY_X; c += 32)
    {
      int j = c;
      if (j < 2 * TY)
      {
        const int ty = j >> 1, p = j & 1;        // the TK sites of (row ty, parity p): one contiguous run of TK blocks
        bulk_g2s(sHx + (size_t)(j * TK) * LPS, hopx + ((size_t)p * half + (size_t)(y0 + ty) * xh + k0) * LPS, TK * LPS * sizeof(cd), bar);
        continue;
      }
      j -= 2 * TY;
      if (j < TY)
      {
        const int ty = j, p = 1 - ((y0 + ty) & 1), k = (k0 == 0) ? xh - 1 : k0 - 1;     // left neighbour of the sft = 0 site of this row
        bulk_g2s(sHx + (size_t)(S + ty) * LPS, hopx + ((size_t)p * half + (size_t)(y0 + ty) * xh + k) * LPS, LPS * sizeof(cd), bar);
        continue;
      }
      j -= TY;
      {
        // spinor row ry - 1 of the patch (ry = 0 .. TY + 1), parity p, sites k0 - 1 .. k0 + TK: contiguous except across the x wrap
        const int ry = j >> 1, p = j & 1;
        int y = y0 + ry - 1; y = (y < 0) ? Y - 1 : ((y >= Y) ? 0 : y);
        const cd* row = a.in + ((size_t)p * half + (size_t)y * xh) * NC;
        cd* dst = sV + (size_t)((ry * 2 + p) * VK) * NC;
        const bool wl = (k0 == 0), wr = (k0 + TK == xh);
        const int kk0 = wl ? 1 : 0, kk1 = wr ? VK - 1 : VK;     // [kk0, kk1): the run that does not wrap
        bulk_g2s(dst + kk0 * NC, row + (size_t)(k0 - 1 + kk0) * NC, (unsigned)((kk1 - kk0) * NC * sizeof(cd)), bar);
        if (wl) bulk_g2s(dst, row + (size_t)(xh - 1) * NC, NC * sizeof(cd), bar);
        if (wr) bulk_g2s(dst + (size_t)(VK - 1) * NC, row, NC * sizeof(cd), bar);
      }
    }
  }
  __device__ static __forceinline__ void stage_y(const StencilKArgs& a, cd* sHy, unsigned long long* bar, int k0, int y0, int lane)
  {
    const int xh = a.g.xh, Y = a.g.Y;
    const size_t half = a.g.half;
    const cd* hopy = a.hop + a.size_cm;
    for (int j = lane; j < NCOPY_Y; j += 32)
    {
      if (j < 2 * TY)
      {
        const int ty = j >> 1, p = j & 1;
        bulk_g2s(sHy + (size_t)(j * TK) * LPS, hopy + ((size_t)p * half + (size_t)(y0 + ty) * xh + k0) * LPS, TK * LPS * sizeof(cd), bar);
      }
      else
      {
        const int p = j - 2 * TY, y = (y0 == 0) ? Y - 1 : y0 - 1;       // the row below the patch
        bulk_g2s(sHy + (size_t)(S + p * TK) * LPS, hopy + ((size_t)p * half + (size_t)y * xh + k0) * LPS, TK * LPS * sizeof(cd), bar);
      }
    }
  }
  // the one-patch kernel: everything into one buffer [Hx | Hy | V], one barrier
  __device__ static __forceinline__ void stage(const StencilKArgs& a, cd* buf, unsigned long long* bar, int k0, int y0, int lane)
  {
    cd* sHx = buf; cd* sHy = sHx + (size_t)NHX * LPS; cd* sV = sHy + (size_t)NHY * LPS;
    stage_x(a, sHx, sV, bar, k0, y0, lane);
    stage_y(a, sHy, bar, k0, y0, lane);
  }

  __device__ static __forceinline__ void compute(const StencilKArgs& a, const cd* buf, int k0, int y0, int tid, const cd (&CLc)[PASSES][R], const cd (&RBc)[PASSES])
  {
    const cd* sHx = buf; const cd* sHy = sHx + (size_t)NHX * LPS; const cd* sV = sHy + (size_t)NHY * LPS;
    const int c2 = tid % NC, hf = (tid / NC) % SPLIT;
    const cd zero = cmake(0.0, 0.0);
    const bool top = (2 * c2 < NC);
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++)
    {
      const int slot = ps * (NT / (NC * SPLIT)) + tid / (NC * SPLIT);
      const int ty = slot / (2 * TK); const int r = slot - ty * 2 * TK; const int p = r / TK, tk = r - p * TK;
      const int q = 1 - p, y = y0 + ty, sft = (y + p) & 1;
      const int nx = sft ? (ty * 2 + q) * TK + tk : (tk > 0 ? (ty * 2 + q) * TK + tk - 1 : S + ty);
      const int ny = (ty > 0) ? ((ty - 1) * 2 + q) * TK + tk : S + q * TK + tk;
      const cd* vrow = sV + ((size_t)((ty + 1) * 2) * VK) * NC;            // row ty of the patch, parity 0, kk = 0
      const cd VC = vrow[((size_t)p * VK + tk + 1) * NC + c2];
      const cd V0 = vrow[((size_t)q * VK + tk + 1 + sft) * NC + c2];
      const cd V1 = vrow[((size_t)(2 + q) * VK + tk + 1) * NC + c2];
      const cd* v2 = vrow + ((size_t)q * VK + tk + sft) * NC + hf * R;                      // in(x - x^)[rows of this thread]: broadcast reads
      const cd* v3 = vrow + ((ptrdiff_t)q * VK + tk + 1 - 2 * VK) * NC + hf * R;            // in(x - y^)
      const cd* hx = sHx + (size_t)slot * LPS + (hf * R) * NC + c2;    // forward: column c2, rows of this thread: [c1][c2] at + i NC
      const cd* hy = sHy + (size_t)slot * LPS + (hf * R) * NC + c2;
      const cd* bx = sHx + (size_t)nx * LPS + (hf * R) * NC + c2;      // backward: column a = c2 of the neighbour's block, rows b of this thread
      const cd* by = sHy + (size_t)ny * LPS + (hf * R) * NC + c2;
      cd acc[R];
      cd backx = zero, backy = zero;      // two independent chains
#pragma unroll
      for (int i = 0; i < R; i++)
      {
        cd t = zero;
        cfma(t, CLc[ps][i], VC);
        cfma(t, hx[i * NC], V0);
        cfma(t, hy[i * NC], V1);
        acc[i] = t;
        // s_a s_b conj(B[b][a]) in(x - mu)[b], b = hf R + i: the sign is applied once below
        cfma_conj(backx, bx[i * NC], v2[i]);
        cfma_conj(backy, by[i * NC], v3[i]);
      }
      cd back = cadd(backx, backy);
      // rows b of this thread all lie in one half of the dof when R divides NC / 2: one sign per thread
      static_assert((NC / 2) % R == 0, "TmaTile: the rows of a thread must not straddle the two chiral halves");
      const double sg = (top == (2 * hf * R < NC)) ? 1.0 : -1.0;
      back = cmake(sg * back.x, sg * back.y);
      // complete the sum over b across the SPLIT threads of (site, a)
#pragma unroll
      for (int off = NC * (SPLIT / 2); off >= NC; off >>= 1) back = cadd(back, shfl_xor_c(back, off));
      // row a = c2 of the forward partial sums lives on the thread with hf = c2 / R, as acc[c2 % R]: add the backward row sum there
#pragma unroll
      for (int i = 0; i < R; i++) if (hf * R + i == c2) acc[i] = cadd(acc[i], back);
      if (a.use_diag)
      {
        const cd dg = a.diag[p][top ? 0 : 1];
#pragma unroll
        for (int i = 0; i < R; i++) if (hf * R + i == c2) cfma(acc[i], dg, VC);
      }
      int left = R;
#pragma unroll
      for (int off = NC / 2; off > 0; off >>= 1)
      {
        if (left > 1)
        {
          const int hl = left / 2;
          const bool upper = (c2 & off) != 0;
#pragma unroll
          for (int i = 0; i < R / 2; i++)
            if (i < hl)
            {
              const cd send = upper ? acc[i] : acc[i + hl];
              const cd keep = upper ? acc[i + hl] : acc[i];
              acc[i] = cadd(keep, shfl_xor_c(send, off));
            }
          left = hl;
        }
        else acc[0] = cadd(acc[0], shfl_xor_c(acc[0], off));
      }
      if (SPLIT == 1 || (c2 % SPLIT) == 0)
      {
        const int row = hf * R + c2 / SPLIT;
        const size_t site = (size_t)p * a.g.half + (size_t)y * a.g.xh + (k0 + tk);
        cd res = acc[0];
        if (a.accumulate) res = cadd(res, a.out[site * NC + row]);
        if (a.resid != nullptr) res = csub(RBc[ps], res);
        a.out[site * NC + row] = res;
      }
    }
  }
};

template <int NC, int TK, int TY, int SPLIT>
__global__ void __launch_bounds__(TileDims<NC, TK, TY, SPLIT>::THREADS, (SPLIT > 1 ? 1024 / TileDims<NC, TK, TY, SPLIT>::THREADS : 1)) stencil_tma_kernel(const StencilKArgs a)
{
  typedef TmaTile<NC, TK, TY, SPLIT> T;
  typedef Tile<NC, TK, TY, SPLIT> T0;          // the clover / residual register prefetch is shared with the cp.async kernel
  extern __shared__ __align__(128) unsigned char tma_smem[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(tma_smem);
  cd* buf = reinterpret_cast<cd*>(tma_smem + 128);
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * TK, y0 = a.y_off + blockIdx.y * TY;
  // warp 0 sets the barrier up and issues the copies at once; the other warps only meet the barrier after their clover
  // loads are in flight (the __syncthreads orders the initialisation before everybody's wait)
  if (tid < 32)
  {
    if (tid == 0)
    {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(bar, T::BYTES);
    }
    __syncwarp();
    T::stage(a, buf, bar, k0, y0, tid);
  }
  cd CLc[T::PASSES][T::R];
  cd RBc[T::PASSES];
  T0::clover(a, k0, y0, tid, CLc, RBc);
  __syncthreads();
  mbar_wait(bar, 0);
  T::compute(a, buf, k0, y0, tid, CLc, RBc);
}

template <int NC, int TK, int TY, int SPLIT = 1> static int launch_tma(const StencilKArgs& a)
{
  static bool configured[64] = { false };
  const size_t smem = TmaTile<NC, TK, TY, SPLIT>::SMEM;
  const int dev = rt().device & 63;
  if (!configured[dev]) { QMG_CUDA(cudaFuncSetAttribute(stencil_tma_kernel<NC, TK, TY, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured[dev] = true; }
  dim3 grid(a.g.xh / TK, a.y_cnt / TY, 1);
  if (grid.y > 65535) return fail_msg("qmg_stencil_apply: Y too large for the launch grid");
  stencil_tma_kernel<NC, TK, TY, SPLIT><<<grid, TileDims<NC, TK, TY, SPLIT>::THREADS, smem, rt().stream>>>(a);
  QMG_LAUNCH_CHECK();
  return 0;
}

// ---- persistent, warp-specialised flavour: one CTA per SM, a ring of NSTAGE patch buffers filled by a producer warp --------
// Both one-patch-per-CTA kernels above wait for their tile with two CTAs per SM (registers and shared memory both stop at
// two), and that wait is what bounds them.  Here ONE resident CTA walks over patches b, b + gridDim, ...: a producer warp
// issues the bulk copies of a patch into the ring as soon as the buffer's previous tenant has been consumed (empty barrier),
// 512 consumer threads take patch after patch as its full barrier completes, and each consumer fetches what it reads straight
// from global memory for its NEXT patch (clover column, residual element, halo blocks) while it computes the current one.
//
// The stage holds the forward blocks of the patch's sites, the left halo column (+x) and the halo row below (+y) that the
// backward hops of its edge sites need, and the spinor tile cut down to the sites that are read (one ring site per row and
// parity, none on the rows above and below): 85 KB, two stages.  (Three 75 KB stages -- the halo row below read from global
// memory into registers one patch ahead by the threads that use it -- were built and measured: 2.81 ms against 2.30 ms,
// profiles/r05_ring_three_stages.txt.  With 17 warps the SM sub-partition that hosts five of them caps a thread at 96
// registers, and the clover column plus the halo rows cannot both be kept in flight a patch ahead inside that budget.
// L2 prefetches by the producer for the patch after the ones in the ring -- cp.async.bulk.prefetch.L2 per 4 KB run or
// prefetch.global.L2 per line, one to three patches ahead -- make it slower, 2.37 -> 2.83 - 2.95 ms, copies alone 1.9 -> 2.5.)
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
// bounded wait: false (and *err set) if the phase did not complete within ~2 s -- a broken pipeline must not hang the GPU
__device__ __forceinline__ bool mbar_wait_bounded(unsigned long long* bar, unsigned parity, unsigned long long* err)
{
  const unsigned addr = smem_u32(bar);
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  if (ok) return true;
  const long long t0 = clock64();
  for (;;)
  {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 4000000000LL) { st_sys_u64(err, 1000ull); return false; }      // reported by the next fetch / qmg_sync
  }
}

template <int NC, int TK, int TY, int SPLIT, int NSTAGE> struct RingCfg
{
  typedef TileDims<NC, TK, TY, SPLIT> D;
  static const int S = D::S, LPS = NC * NC, VK = TK + 1, R = NC / SPLIT;
  static const int GT = D::THREADS;                                // consumer threads = one patch
  static const int NTHREADS = GT + 32;
  // spinor tile: [row below: 2 parities x TK sites][rows of the patch: 2 parities x (TK + 1) sites][row above: 2 x TK]; a row
  // of the patch holds sites k0 - 1 + sh .. k0 + TK - 1 + sh of its parity P, sh = (y + 1 - P) & 1: the sites themselves plus
  // the ONE ring site that the other parity's +-x hops reach
  static const int NV = (4 * TK + TY * 2 * VK) * NC;
  static const size_t BHX = sizeof(cd) * (size_t)(S + TY) * LPS;   // +x blocks of the patch's sites, then the left halo column
  static const size_t BHY = sizeof(cd) * (size_t)(S + 2 * TK) * LPS;      // +y blocks of the patch's sites, then the halo row below (parity 0, parity 1)
  static const size_t BV = sizeof(cd) * (size_t)NV;
  static const size_t STAGE = BHX + BHY + BV;
  static_assert(STAGE % 128 == 0, "ring kernel: stage size");
  static const unsigned TX_BYTES = (unsigned)STAGE;
  static const size_t SMEM = 256 + NSTAGE * STAGE;
  static const int NCOPY = 2 * TY + TY + 2 * TY + 2 + 2 * (TY + 2);

  // Issued by the producer warp: lane l takes copies l, l + 32, ...  The TK blocks of one (row, parity) are contiguous in the
  // reference's layout (4 KB at nc = 8, TK = 4); a spinor row with its ring site is contiguous except across the x wrap.
  __device__ static __forceinline__ void stage(const StencilKArgs& a, unsigned char* buf, unsigned long long* bar, int k0, int y0, int lane)
  {
    cd* sHx = reinterpret_cast<cd*>(buf); cd* sHy = reinterpret_cast<cd*>(buf + BHX); cd* sV = reinterpret_cast<cd*>(buf + BHX + BHY);
    const int xh = a.g.xh, Y = a.g.Y;
    const size_t half = a.g.half;
    for (int c = lane; c < NCOPY; c += 32)
    {
      int j = c;
      if (j < 2 * TY)
      {
        const int ty = j >> 1, p = j & 1;
        bulk_g2s(sHx + (size_t)(j * TK) * LPS, a.hop + ((size_t)p * half + (size_t)(y0 + ty) * xh + k0) * LPS, TK * LPS * sizeof(cd), bar);
        continue;
      }
      j -= 2 * TY;
      if (j < TY)
      {
        const int ty = j, p = 1 - ((y0 + ty) & 1), k = (k0 == 0) ? xh - 1 : k0 - 1;     // left neighbour of the row's first site that hops back inside its parity index
        bulk_g2s(sHx + (size_t)(S + ty) * LPS, a.hop + ((size_t)p * half + (size_t)(y0 + ty) * xh + k) * LPS, LPS * sizeof(cd), bar);
        continue;
      }
      j -= TY;
      if (j < 2 * TY)
      {
        const int ty = j >> 1, p = j & 1;
        bulk_g2s(sHy + (size_t)(j * TK) * LPS, a.hop + a.size_cm + ((size_t)p * half + (size_t)(y0 + ty) * xh + k0) * LPS, TK * LPS * sizeof(cd), bar);
        continue;
      }
      j -= 2 * TY;
      if (j < 2)
      {
        const int p = j, y = (y0 == 0) ? Y - 1 : y0 - 1;       // the row below the patch
        bulk_g2s(sHy + (size_t)(S + p * TK) * LPS, a.hop + a.size_cm + ((size_t)p * half + (size_t)y * xh + k0) * LPS, TK * LPS * sizeof(cd), bar);
        continue;
      }
      j -= 2;
      {
        const int ry = j >> 1, p = j & 1;        // ry = 0: the row below, 1 .. TY: the patch, TY + 1: the row above
        int y = y0 + ry - 1; y = (y < 0) ? Y - 1 : ((y >= Y) ? 0 : y);
        const cd* row = a.in + ((size_t)p * half + (size_t)y * xh) * NC;
        if (ry == 0 || ry == TY + 1)
        {
          cd* dst = sV + (size_t)((ry == 0 ? 0 : 2 * TK + TY * 2 * VK) + p * TK) * NC;
          bulk_g2s(dst, row + (size_t)k0 * NC, TK * NC * sizeof(cd), bar);
        }
        else
        {
          const int sh = (y0 + ry - 1 + 1 - p) & 1;
          cd* dst = sV + (size_t)(2 * TK + ((ry - 1) * 2 + p) * VK) * NC;
          const int ks = k0 - 1 + sh;            // first site of the run of TK + 1
          if (ks < 0)
          {
            bulk_g2s(dst, row + (size_t)(xh - 1) * NC, NC * sizeof(cd), bar);
            bulk_g2s(dst + NC, row, TK * NC * sizeof(cd), bar);
          }
          else if (ks + VK > xh)
          {
            bulk_g2s(dst, row + (size_t)ks * NC, TK * NC * sizeof(cd), bar);
            bulk_g2s(dst + (size_t)TK * NC, row, NC * sizeof(cd), bar);
          }
          else bulk_g2s(dst, row + (size_t)ks * NC, VK * NC * sizeof(cd), bar);
        }
      }
    }
  }
};

// The consumers run the patch arithmetic for patch after patch with the SAME thread -> (site, column, rows) assignment, so
// everything that depends on the thread alone -- the element offsets of its nine operand streams inside a stage, its signs,
// which of its rows receives the backward sum, where its output element sits relative to the patch origin -- is worked out
// ONCE (a third of the consumer's instructions were this index arithmetic: 289 IMAD against 92 DFMA in the SASS) and a patch
// costs the loads, the products, the butterfly and one add for the global offset of its origin.  Needs the parity of the first
// row of every patch to be the same (TY even, so y0 = y_off + by TY has the parity of y_off).
template <int NC, int TK, int TY, int SPLIT> struct RingInv
{
  static const int R = NC / SPLIT, LPS = NC * NC, S = TileDims<NC, TK, TY, SPLIT>::S, VK = TK + 1;
  int oVC, oV0, oV1, ov2, ov3;       // element offsets into the spinor tile of the stage
  int ohf, obx, oby;                 // ... into its block regions: own forward blocks (same offset in +x and +y), backward neighbours
  long out_local, cl_local;          // output element / first clover element relative to the patch origin (units of elements)
  int back_row;                      // index i of acc[] that receives the backward row sum on this thread, or -1
  bool neg;                          // the backward sum of this thread's rows enters with a minus sign (rows and column in different chiral halves)
  bool writer;
  int dgi;                           // which of the four diagonal shifts applies to this thread's column

  __device__ __forceinline__ void init(const StencilKArgs& a, int tid, int y_parity)
  {
    static_assert((TY & 1) == 0, "RingInv: TY must be even");
    const int c2 = tid % NC, hf = (tid / NC) % SPLIT;
    const bool top = (2 * c2 < NC);
    const int slot = tid / (NC * SPLIT);
    const int ty = slot / (2 * TK); const int r = slot - ty * 2 * TK; const int p = r / TK, tk = r - p * TK;
    const int q = 1 - p, sft = (y_parity + ty + p) & 1;
    const int el = (hf * R) * NC + c2;                 // first element of this thread inside a block: [hf R + i][c2] at + i NC
    const int nx = sft ? (ty * 2 + q) * TK + tk : (tk > 0 ? (ty * 2 + q) * TK + tk - 1 : S + ty);
    const int ny = (ty > 0) ? ((ty - 1) * 2 + q) * TK + tk : S + q * TK + tk;
    // spinor tile (see RingCfg): a row of the patch with parity P starts at site k0 - 1 + sh_P; sh_q = sft, sh_p = 1 - sft
    const int rowp = (2 * TK + (ty * 2 + p) * VK) * NC, rowq = (2 * TK + (ty * 2 + q) * VK) * NC;
    oVC = rowp + (tk + sft) * NC + c2;                 // in(x)
    oV0 = rowq + (tk + 1) * NC + c2;                   // in(x + x^): site k0 + tk + sft of parity q
    ov2 = rowq + tk * NC + hf * R;                     // in(x - x^): site k0 + tk - 1 + sft; rows of this thread, broadcast reads
    oV1 = ((ty + 1 < TY) ? (2 * TK + ((ty + 1) * 2 + q) * VK + tk + sft) : (2 * TK + TY * 2 * VK + q * TK + tk)) * NC + c2;      // in(x + y^)
    ov3 = ((ty > 0) ? (2 * TK + ((ty - 1) * 2 + q) * VK + tk + sft) : (q * TK + tk)) * NC + hf * R;                            // in(x - y^)
    ohf = slot * LPS + el;
    obx = nx * LPS + el;
    oby = ny * LPS + el;
    neg = (top != (2 * hf * R < NC));
    back_row = (c2 >= hf * R && c2 < hf * R + R) ? c2 - hf * R : -1;
    writer = (SPLIT == 1 || (c2 % SPLIT) == 0);
    const int row = hf * R + c2 / SPLIT;
    // site = p half + (y0 + ty) xh + (k0 + tk): the patch origin contributes (y0 xh + k0), the rest is local
    const long site_local = (long)p * a.g.half + (long)ty * a.g.xh + tk;
    out_local = site_local * NC + row;
    cl_local = site_local * LPS + el;
    dgi = p * 2 + (top ? 0 : 1);
  }
};

template <int NC, int TK, int TY, int SPLIT, int NSTAGE>
__global__ void __launch_bounds__(RingCfg<NC, TK, TY, SPLIT, NSTAGE>::NTHREADS, 1)
stencil_ring_kernel(const StencilKArgs a, const int npatch, const int nbx, unsigned long long* err, const int dbg)
{
  typedef RingCfg<NC, TK, TY, SPLIT, NSTAGE> CFG;
  static_assert(TileDims<NC, TK, TY, SPLIT>::PASSES == 1, "ring kernel: one patch per pass of the consumer group");
  static_assert(CFG::SMEM <= 232448, "ring kernel: stages do not fit in shared memory");
  extern __shared__ __align__(128) unsigned char ring_smem[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring_smem);
  unsigned long long* empty = full + NSTAGE;
  const int tid = threadIdx.x;
  if (tid == 0)
  {
    for (int s = 0; s < NSTAGE; s++) { mbar_init(full + s, 1); mbar_init(empty + s, CFG::GT / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid >= CFG::GT)
  {
    // producer warp
    const int lane = tid - CFG::GT;
    int s = 0; unsigned ph = 1;          // waiting on a fresh barrier with parity 1 returns at once: the ring starts empty
    for (int t = blockIdx.x; t < npatch; t += gridDim.x)
    {
      const int by = t / nbx, bx = t - by * nbx;
      if (!mbar_wait_bounded(empty + s, ph, err)) return;
      if (dbg == 2) { if (lane == 0) mbar_arrive(full + s); }      // timing experiment 2: no copies, the consumers compute on whatever is there
      else
      {
        if (lane == 0) mbar_expect_tx(full + s, CFG::TX_BYTES);
        __syncwarp();
        CFG::stage(a, ring_smem + 256 + (size_t)s * CFG::STAGE, full + s, bx * TK, a.y_off + by * TY, lane);
      }
      if (++s == NSTAGE) { s = 0; ph ^= 1u; }
    }
    return;
  }
  // consumers
  RingInv<NC, TK, TY, SPLIT> inv;
  inv.init(a, tid, a.y_off & 1);
  constexpr int R = CFG::R, LPS = CFG::LPS;
  const cd zero = cmake(0.0, 0.0);
  const bool has_cl = (a.clover != nullptr), has_rb = (a.resid != nullptr) && inv.writer, acc_old = (a.accumulate != 0) && inv.writer;
  const int xh = a.g.xh;
  // this CTA's patch sequence without a division per patch: t advances by gridDim.x, (bx, by) by the matching (dx, dy)
  const int stride = (int)gridDim.x;
  const int dy = stride / nbx, dx = stride - dy * nbx;
  int t = blockIdx.x;
  int by = t / nbx, bx = t - by * nbx;
  cd CLc[R], CLn[R];
  cd RBc = zero, RBn = zero;
#pragma unroll
  for (int i = 0; i < R; i++) { CLc[i] = zero; CLn[i] = zero; }
  // the operands that come straight from global memory (clover column, residual element) travel in registers, one patch ahead
  auto fetch = [&](int fbx, int fby, cd (&CL)[R], cd& RB)
  {
    const long origin = (long)(a.y_off + fby * TY) * xh + fbx * TK;
    if (has_cl) { const cd* cp = a.clover + origin * LPS + inv.cl_local;
#pragma unroll
      for (int i = 0; i < R; i++) CL[i] = ld_stream(cp + i * NC); }
    if (has_rb) RB = ld_stream(a.resid + origin * NC + inv.out_local);
  };
  if (t < npatch) fetch(bx, by, CLc, RBc);
  int s = 0; unsigned ph = 0;
  for (; t < npatch; t += stride)
  {
    const long origin = (long)(a.y_off + by * TY) * xh + bx * TK;
    bx += dx; by += dy;
    if (bx >= nbx) { bx -= nbx; by++; }
    if (t < npatch - stride) fetch(bx, by, CLn, RBn);
    const unsigned char* buf = ring_smem + 256 + (size_t)s * CFG::STAGE;
    const cd* sHx = reinterpret_cast<const cd*>(buf);
    const cd* sHy = reinterpret_cast<const cd*>(buf + CFG::BHX);
    const cd* sV = reinterpret_cast<const cd*>(buf + CFG::BHX + CFG::BHY);
    cd acc[R];
    cd backx = zero, backy = zero;      // two independent chains
#pragma unroll
    for (int i = 0; i < R; i++) acc[i] = zero;
    if (!mbar_wait_bounded(full + s, ph, err)) return;
    const cd VC = sV[inv.oVC];
    if (dbg != 1)      // (timing experiment 1: copies only, nothing computed)
    {
      const cd V0 = sV[inv.oV0], V1 = sV[inv.oV1];
      const cd* hx = sHx + inv.ohf; const cd* hy = sHy + inv.ohf;
      const cd* bxp = sHx + inv.obx; const cd* byp = sHy + inv.oby;
      const cd* v2 = sV + inv.ov2; const cd* v3 = sV + inv.ov3;
#pragma unroll
      for (int i = 0; i < R; i++)
      {
        cd tt = zero;
        cfma(tt, CLc[i], VC);
        cfma(tt, hx[i * NC], V0);
        cfma(tt, hy[i * NC], V1);
        acc[i] = tt;
        // s_a s_b conj(B[b][a]) in(x - mu)[b], b = hf R + i: the sign is applied once below
        cfma_conj(backx, bxp[i * NC], v2[i]);
        cfma_conj(backy, byp[i * NC], v3[i]);
      }
    }
    // every operand of this patch is in registers: hand the stage back
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty + s);
    if (++s == NSTAGE) { s = 0; ph ^= 1u; }
    if (dbg != 1)
    {
      cd back = cadd(backx, backy);
      if (inv.neg) back = cmake(-back.x, -back.y);
#pragma unroll
      for (int off = NC * (SPLIT / 2); off >= NC; off >>= 1) back = cadd(back, shfl_xor_c(back, off));
      // the thread whose row set holds row a = c2 adds the backward row sum (and the diagonal shift on that row) to it
      if (a.use_diag) cfma(back, (&a.diag[0][0])[inv.dgi], VC);
#pragma unroll
      for (int i = 0; i < R; i++) if (inv.back_row == i) acc[i] = cadd(acc[i], back);
      const int c2 = tid % NC;
      int left = R;
#pragma unroll
      for (int off = NC / 2; off > 0; off >>= 1)
      {
        if (left > 1)
        {
          const int hl = left / 2;
          const bool upper = (c2 & off) != 0;
#pragma unroll
          for (int i = 0; i < R / 2; i++)
            if (i < hl)
            {
              const cd send = upper ? acc[i] : acc[i + hl];
              const cd keep = upper ? acc[i + hl] : acc[i];
              acc[i] = cadd(keep, shfl_xor_c(send, off));
            }
          left = hl;
        }
        else acc[0] = cadd(acc[0], shfl_xor_c(acc[0], off));
      }
      if (inv.writer)
      {
        cd* op = a.out + origin * NC + inv.out_local;
        cd res = acc[0];
        if (acc_old) res = cadd(res, *op);
        if (has_rb) res = csub(RBc, res);
        *op = res;
      }
    }
#pragma unroll
    for (int i = 0; i < R; i++) CLc[i] = CLn[i];
    RBc = RBn;
  }
}

static int ring_debug() { static int v = -1; if (v < 0) { const char* e = getenv("QMG_RING_DEBUG"); v = e ? atoi(e) : 0; } return v; }

template <int NC, int TK, int TY, int SPLIT, int NSTAGE> static int launch_ring(const StencilKArgs& a)
{
  typedef RingCfg<NC, TK, TY, SPLIT, NSTAGE> CFG;
  static bool configured[64] = { false };
  const int dev = rt().device & 63;
  if (!configured[dev])
  {
    QMG_CUDA(cudaFuncSetAttribute(stencil_ring_kernel<NC, TK, TY, SPLIT, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CFG::SMEM));
    configured[dev] = true;
  }
  const int nbx = a.g.xh / TK, nby = a.y_cnt / TY;
  const long npatch = (long)nbx * nby;
  if (npatch > 0x7fffffffL) return fail_msg("qmg_stencil_apply: lattice too large for the ring kernel");
  int grid = rt().sm_count;
  if (grid > npatch) grid = (int)npatch;
  stencil_ring_kernel<NC, TK, TY, SPLIT, NSTAGE><<<grid, CFG::NTHREADS, CFG::SMEM, rt().stream>>>(a, (int)npatch, nbx, rt().h_err, ring_debug());
  QMG_LAUNCH_CHECK();
  return 0;
}

// nc = 8 patches (QMG_TILE / qmg_set_tile_kernel).  1 (default): the persistent ring kernel -- 32-site patches, two stages,
// 512 loop-invariant consumer threads + a producer warp -- wherever every SM
// gets at least 8 patches, else the one-patch cp.async kernel; 9: always the ring; 3: always the cp.async kernel; 2: cp.async
// with one thread per column; 4: one-patch TMA kernel; 5 / 6: 16-site one-patch kernels.  (Ring kernels over 16-site patches
// with 2 - 3 consumer groups, modes 7 / 8 / 10 / 11 of an earlier build, were producer-bound -- 34 resp. 20 bulk copies per
// 16 sites -- and are gone, as is a ring whose x and y halves were recycled separately: the wait between the halves cost the
// consumers their instruction-level parallelism.  Timings: profiles/r03r_ring_kernel_variants.txt, r05_ring_three_stages.txt.)
static int launch_tile8(const StencilKArgs& a)
{
  const int mode = rt().tile_kernel;
  if (mode == 2) return launch_tile<8, 4, 4, 1>(a);
  if (mode == 4) return launch_tma<8, 4, 4, 2>(a);
  if (mode == 5) return launch_tile<8, 2, 4, 2>(a);
  if (mode == 6) return launch_tma<8, 2, 4, 2>(a);
  const long npatch = (long)(a.g.xh / 4) * (a.y_cnt / 4);
  if (mode == 9 || (mode == 1 && npatch >= 8L * rt().sm_count)) return launch_ring<8, 4, 4, 2, 2>(a);
  return launch_tile<8, 4, 4, 2>(a);
}

// (nc = 2, the Wilson fine level: a link-compressed patch kernel in the same style -- 4 rows x 32 sites per CTA, one lane per
// matrix element, forward and transposed backward elements out of shared memory, clover element straight from global memory,
// 256 instead of 384 bytes per site -- was built, passed the parity tests and was removed again: 5.39 ms on 8192^2 against
// 3.60 ms for the stored-block apply and 4.16 ms for the streaming link-compressed one.  A 64-byte block is four lanes' worth of
// data; per lane the staging and index arithmetic cost more instructions than the 128 bytes saved are worth, and the
// stored-block kernel already runs at 54 % issue utilisation.  At nc = 2 the stored blocks are the fast path; the K-cycle legs
// of bench.py switch only the nc = 8 levels to link-compressed applies.)

// Any nc below 16 (DWF Ls = 6 gives nc = 12; odd dof counts): one thread per (site, row), looping over the columns; the rows of a
// site are short enough for L1 to catch the sectors a warp's strided requests share (nc = 12 on 1024^2: 0.78 of the copy peak, the
// lane-group kernel below 0.72 - 0.75).
__global__ void __launch_bounds__(256) stencil_kernel_rows(const StencilKArgs a, int nc, int n_par)
{
  const long sub_half = (long)a.y_cnt * a.g.xh;     // sites per parity in this launch's row set
  const long rows = (long)n_par * sub_half * nc;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < rows; t += (long)gridDim.x * blockDim.x)
  {
    const int c1 = (int)(t % nc);
    const long s = t / nc;
    const int p = a.p_begin + (int)(s / sub_half);
    const long hs = s % sub_half;
    const int y = a.y_off + (int)(hs / a.g.xh) * a.y_stride, k = (int)(hs % a.g.xh);
    const unsigned h = (unsigned)y * a.g.xh + k;
    const size_t site = (size_t)p * a.g.half + h;
    const int q = 1 - p;
    const size_t lps = (size_t)nc * nc;
    cd acc = cmake(0.0, 0.0);
    if (a.clover != nullptr)
      for (int c2 = 0; c2 < nc; c2++) cfma(acc, a.clover[site * lps + (size_t)c1 * nc + c2], a.in[site * nc + c2]);
    if (a.use_diag) cfma(acc, a.diag[p][(2 * c1 >= nc && nc > 1) ? 1 : 0], a.in[site * nc + c1]);
    if (a.hop != nullptr && a.hop_to[p])
    {
      const cd* in_q = a.in + (size_t)q * a.g.half * nc;
      for (int mu = 0; mu < 4; mu++)
      {
        if (!((a.dir_mask >> mu) & 1)) continue;
        const cd* src;
        if (mu == 1 && a.halo_yp != nullptr && y == a.g.Y - 1) src = a.halo_yp + ((size_t)q * a.g.xh + k) * nc;
        else if (mu == 3 && a.halo_ym != nullptr && y == 0) src = a.halo_ym + ((size_t)q * a.g.xh + k) * nc;
        else src = in_q + (size_t)nbr_h(a.g, p, y, k, mu) * nc;
        const cd* m = a.hop + (size_t)mu * a.size_cm + site * lps + (size_t)c1 * nc;
        for (int c2 = 0; c2 < nc; c2++) cfma(acc, m[c2], src[c2]);
      }
    }
    const size_t idx = site * nc + c1;
    if (a.accumulate) acc = cadd(acc, a.out[idx]);
    if (a.resid != nullptr) acc = csub(a.resid[idx], acc);
    a.out[idx] = acc;
  }
}

// Any nc from 16 up (DWF Ls = 12, 24, 32 give nc = 24, 48, 64): G = 8 lanes per (site, row), lane g taking the columns g, g + G, ...
// of the row -- consecutive lanes read consecutive matrix and spinor elements, so a warp request covers 32 / G runs of G x 16
// contiguous bytes (whole sectors) instead of 32 single elements one matrix row apart -- and a log2(G)-step shuffle tree
// finishing the row sum; the five blocks of a row keep their own accumulators (10 loads in flight per step).  nc = 24 / 48:
// 0.62 -> 0.77 / 0.83 of the copy peak (profiles/r05p_kernel_probe_any_nc.txt).  (The first version ran one thread per row over all its columns.)
template <int G>
__global__ void __launch_bounds__(256, 2) stencil_kernel_generic(const StencilKArgs a, int nc, int n_par)
{
  const long sub_half = (long)a.y_cnt * a.g.xh;     // sites per parity in this launch's row set
  const long rows = (long)n_par * sub_half * nc;
  const int g = threadIdx.x % G;
  // a warp stays together (the row sums are finished with full-warp shuffles): lane groups beyond the last row recompute it and do not store
  for (long t0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) / G; ; t0 += ((long)gridDim.x * blockDim.x) / G)
  {
    const bool active = t0 < rows;
    if (!__any_sync(0xffffffffu, active)) break;
    const long t = active ? t0 : rows - 1;
    const int c1 = (int)(t % nc);
    const long s = t / nc;
    const int p = a.p_begin + (int)(s / sub_half);
    const long hs = s % sub_half;
    const int y = a.y_off + (int)(hs / a.g.xh) * a.y_stride, k = (int)(hs % a.g.xh);
    const unsigned h = (unsigned)y * a.g.xh + k;
    const size_t site = (size_t)p * a.g.half + h;
    const int q = 1 - p;
    const size_t lps = (size_t)nc * nc;
    // the five operand streams of the row (matrix row, spinor) with their own accumulators: 10 loads in flight per column step
    const cd* mp[5]; const cd* vp[5]; bool on[5];
    on[4] = (a.clover != nullptr);
    mp[4] = on[4] ? a.clover + site * lps + (size_t)c1 * nc : a.in;
    vp[4] = a.in + site * nc;
    const bool hop = (a.hop != nullptr) && a.hop_to[p];
    const cd* in_q = a.in + (size_t)q * a.g.half * nc;
#pragma unroll
    for (int mu = 0; mu < 4; mu++)
    {
      on[mu] = hop && ((a.dir_mask >> mu) & 1);
      const cd* src;
      if (mu == 1 && a.halo_yp != nullptr && y == a.g.Y - 1) src = a.halo_yp + ((size_t)q * a.g.xh + k) * nc;
      else if (mu == 3 && a.halo_ym != nullptr && y == 0) src = a.halo_ym + ((size_t)q * a.g.xh + k) * nc;
      else src = in_q + (size_t)nbr_h(a.g, p, y, k, mu) * nc;
      vp[mu] = src;
      mp[mu] = on[mu] ? a.hop + (size_t)mu * a.size_cm + site * lps + (size_t)c1 * nc : a.in;
    }
    cd part[5];
#pragma unroll
    for (int b = 0; b < 5; b++) part[b] = cmake(0.0, 0.0);
    for (int c2 = g; c2 < nc; c2 += G)
    {
      cd mm[5], vv[5];
#pragma unroll
      for (int b = 0; b < 5; b++) if (on[b]) { mm[b] = ld_stream(mp[b] + c2); vv[b] = vp[b][c2]; }
#pragma unroll
      for (int b = 0; b < 5; b++) if (on[b]) cfma(part[b], mm[b], vv[b]);
    }
    // clover first, then +x, +y, -x, -y: the order of the one-thread-per-row loop
    cd acc = part[4];
#pragma unroll
    for (int b = 0; b < 4; b++) acc = cadd(acc, part[b]);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc = cadd(acc, shfl_xor_c(acc, o));
    if (g == 0 && active)
    {
      if (a.use_diag) cfma(acc, a.diag[p][(2 * c1 >= nc && nc > 1) ? 1 : 0], a.in[site * nc + c1]);
      const size_t idx = site * nc + c1;
      if (a.accumulate) acc = cadd(acc, a.out[idx]);
      if (a.resid != nullptr) acc = csub(a.resid[idx], acc);
      a.out[idx] = acc;
    }
  }
}

static int build_args(const qmg_stencil_desc* st, int pieces, int dir_mask, qmg_cplx* lhs, const qmg_cplx* rhs, StencilKArgs& a, int& n_par)
{
  if (st == nullptr) return fail_msg("qmg_stencil_apply: null stencil");
  const bool single_site = (st->X == 1 && st->Y == 1);   // a volume-1 coarsest level: clover + shift only (stencil_2d.h:868-887)
  if (!single_site && (st->X < 2 || st->Y < 2 || (st->X & 1) || (st->Y & 1))) return fail_msg("qmg_stencil_apply: X and Y must be even and >= 2 (or a single site)");
  if (st->nc < 1) return fail_msg("qmg_stencil_apply: nc < 1");
  if ((const void*)lhs == (const void*)rhs && (pieces & (QMG_APPLY_CLOVER | QMG_APPLY_SHIFT | QMG_APPLY_IDENTITY_CLOVER)))
    return fail_msg("qmg_stencil_apply: in-place apply is only defined for pure hopping pieces");
  a.clover = (pieces & QMG_APPLY_CLOVER) ? reinterpret_cast<const cd*>(st->clover) : nullptr;
  const bool want_hop = pieces & (QMG_APPLY_HOP_TO_EVEN | QMG_APPLY_HOP_TO_ODD);
  a.hop = want_hop ? reinterpret_cast<const cd*>(st->hopping) : nullptr;
  a.in = reinterpret_cast<const cd*>(rhs);
  a.out = reinterpret_cast<cd*>(lhs);
  a.dotw = nullptr;
  a.resid = nullptr;
  a.halo_ym = reinterpret_cast<const cd*>(st->halo_ym);
  a.halo_yp = reinterpret_cast<const cd*>(st->halo_yp);
  a.herm = (st->gamma5_hermitian != 0 && st->nc % 2 == 0 && st->nc <= 32) ? 1 : 0;
  a.hop_ym = reinterpret_cast<const cd*>(st->hop_halo_ym);
  if (a.herm && comm().active && a.hop_ym == nullptr) return fail_msg("qmg_stencil_apply: a gamma5-hermitian link set on a y-slab needs hop_halo_ym (row -1 of the +y blocks)");
  a.g.xh = st->X / 2; a.g.Y = st->Y; a.g.half = (unsigned)(st->X / 2) * st->Y;
  if (single_site) { a.g.xh = 1; a.g.Y = 1; a.g.half = 1; a.hop = nullptr; }
  a.size_cm = (long)st->X * st->Y * st->nc * st->nc;
  const bool sh = pieces & QMG_APPLY_SHIFT;
  const double id = (pieces & QMG_APPLY_IDENTITY_CLOVER) ? 1.0 : 0.0;
  for (int p = 0; p < 2; p++)
    for (int hf = 0; hf < 2; hf++)
    {
      const double se = p ? -1.0 : 1.0, sd = hf ? -1.0 : 1.0;
      // dof_shift only exists for even nc (stencil_2d.h:897)
      const double dr = (st->nc % 2 == 0) ? st->dof_shift[0] : 0.0, di = (st->nc % 2 == 0) ? st->dof_shift[1] : 0.0;
      a.diag[p][hf] = sh ? cmake(id + st->shift[0] + se * st->eo_shift[0] + sd * dr, st->shift[1] + se * st->eo_shift[1] + sd * di)
                         : cmake(id, 0.0);
    }
  a.use_diag = (sh && (st->shift[0] != 0.0 || st->shift[1] != 0.0 || st->eo_shift[0] != 0.0 || st->eo_shift[1] != 0.0 ||
                       st->dof_shift[0] != 0.0 || st->dof_shift[1] != 0.0)) || id != 0.0;
  a.hop_to[0] = (pieces & QMG_APPLY_HOP_TO_EVEN) ? 1 : 0;
  a.hop_to[1] = (pieces & QMG_APPLY_HOP_TO_ODD) ? 1 : 0;
  a.dir_mask = dir_mask & 15;
  a.accumulate = (pieces & QMG_APPLY_ACCUMULATE) ? 1 : 0;
  a.p_begin = 0; n_par = 2;
  if (pieces & QMG_APPLY_EVEN_ROWS_ONLY) { a.p_begin = 0; n_par = 1; }
  if (pieces & QMG_APPLY_ODD_ROWS_ONLY) { a.p_begin = 1; n_par = 1; }
  if (single_site) { a.p_begin = 0; n_par = 1; }
  a.n_par = n_par;
  a.y_off = 0; a.y_stride = 1; a.y_cnt = a.g.Y;
  // matrix-free flavour: only the whole operator (every piece, every direction) of an nc = 2 set that carries its gauge field
  a.mf_gauge = nullptr; a.mf_gauge_ym = nullptr; a.mf_w = 0.0;
  if (st->wilson_gauge != nullptr && st->nc == 2 && !single_site && (pieces & QMG_APPLY_ALL) == QMG_APPLY_ALL &&
      !(pieces & QMG_APPLY_IDENTITY_CLOVER) && (dir_mask & 15) == 15 && st->clover != nullptr && st->hopping != nullptr &&
      (!comm().active || st->wilson_gauge_halo_ym != nullptr))
  {
    a.mf_gauge = reinterpret_cast<const cd*>(st->wilson_gauge);
    a.mf_gauge_ym = reinterpret_cast<const cd*>(st->wilson_gauge_halo_ym);
    a.mf_w = st->wilson_w;
  }
  return 0;
}

// ---- matrix-free Wilson apply (nc = 2, opt-in) ---------------------------------------------------------------------------
// The five stored 2 x 2 blocks of a Wilson site (320 of the apply's 384 bytes) are functions of four U(1) links: clover =
// 2w on the diagonal, hopping_mu[a][b] = coef_mu[a][b] u_mu with coef from {-w/2, +-1/2, +-i/2} (operators/wilson.h:153-209,
// csrc/qmg_setup.cu qmg_fill_wilson), u_{-mu}(x) = conj(u_mu(x - mu)).  When the caller vouches for that -- descriptor field
// wilson_gauge, set by Wilson2D::enable_matrix_free_apply after qmg_wilson_mf_deviation found the stored blocks EQUAL to the
// regenerated ones -- the whole-operator apply reads the links instead: 32 B of links + 32 B in + 32 B out = 96 B per site from
// DRAM (the backward links and the neighbour spinors were fetched by a neighbouring thread and come out of L1 / L2).  One thread
// per site forms the block elements with the fill kernel's cmul and runs the element kernel's chains -- per row c1 the partial
// sums of columns c2 = 0, 1 over clover, shift, +x, +y, -x, -y, then their sum -- so the output has the SAME BITS as the
// stored-block apply.  Pieces, variants (dagger, rbjacobi, Schur) and fused reductions keep reading the stored blocks.
__device__ __forceinline__ void ld256_keep(const cd* p, cd& e0, cd& e1)
{ asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(e0.x), "=d"(e0.y), "=d"(e1.x), "=d"(e1.y) : "l"(p)); }
__device__ __forceinline__ void ld256_stream(const cd* p, cd& e0, cd& e1)
{ asm("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(e0.x), "=d"(e0.y), "=d"(e1.x), "=d"(e1.y) : "l"(p)); }
__device__ __forceinline__ void st256(cd* p, cd e0, cd e1)
{ asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(e0.x), "d"(e0.y), "d"(e1.x), "d"(e1.y) : "memory"); }

// element [c1][c2] of the hopping block in direction mu (+x, +y, -x, -y) for link value u: the arithmetic of qmg_fill_wilson
__device__ __forceinline__ cd wilson_hop_element(int mu, int c1, int c2, cd u, double w)
{
  cd coef;
  if (c1 == c2) coef = cmake(-0.5 * w, 0.0);
  else if (mu == 0) coef = cmake(0.5, 0.0);
  else if (mu == 2) coef = cmake(-0.5, 0.0);
  else if (mu == 1) coef = (c1 == 0) ? cmake(0.0, -0.5) : cmake(0.0, 0.5);
  else coef = (c1 == 0) ? cmake(0.0, 0.5) : cmake(0.0, -0.5);
  return cmul(coef, u);
}

template <bool RESID>
__global__ void __launch_bounds__(256) wilson_mf_kernel(const StencilKArgs a)
{
  constexpr int NC = 2;
  const int bxi = (a.n_par == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int k = bxi * blockDim.x + threadIdx.x;
  const int yi = blockIdx.y * blockDim.y + threadIdx.y;
  const int y = a.y_off + yi * a.y_stride;
  const int p = a.p_begin + ((a.n_par == 2) ? (blockIdx.x & 1) : 0);
  if (!((k < a.g.xh) && (yi < a.y_cnt))) return;
  const cd zero = cmake(0.0, 0.0);
  const size_t V = 2 * (size_t)a.g.half;
  const size_t site = (size_t)p * a.g.half + (size_t)y * a.g.xh + k;
  const int q = 1 - p;
  const int sft = (y + p) & 1;
  int kxp = k + sft; kxp = (kxp == a.g.xh) ? 0 : kxp;
  int kxm = k - 1 + sft; kxm = (kxm < 0) ? a.g.xh - 1 : kxm;
  const int yp1 = (y + 1 == a.g.Y) ? 0 : y + 1;
  const int ym1 = (y == 0) ? a.g.Y - 1 : y - 1;
  const size_t qbase = (size_t)q * a.g.half;
  const cd* in_q = a.in + qbase * NC;
  const cd* s0 = in_q + ((size_t)y * a.g.xh + kxp) * NC;
  const cd* s2 = in_q + ((size_t)y * a.g.xh + kxm) * NC;
  const cd* s1 = (a.halo_yp != nullptr && y == a.g.Y - 1) ? a.halo_yp + ((size_t)q * a.g.xh + k) * NC : in_q + ((size_t)yp1 * a.g.xh + k) * NC;
  const cd* s3 = (a.halo_ym != nullptr && y == 0) ? a.halo_ym + ((size_t)q * a.g.xh + k) * NC : in_q + ((size_t)ym1 * a.g.xh + k) * NC;
  const cd* g2 = a.mf_gauge + qbase + (size_t)y * a.g.xh + kxm;
  const cd* g3 = (a.mf_gauge_ym != nullptr && y == 0) ? a.mf_gauge_ym + ((size_t)q * a.g.xh + k) : a.mf_gauge + V + qbase + (size_t)ym1 * a.g.xh + k;

  // every load ahead of the first product
  cd u[4];
  u[0] = ld_stream(a.mf_gauge + site);
  u[1] = ld_stream(a.mf_gauge + V + site);
  u[2] = ld_keep(g2);
  u[3] = ld_keep(g3);
  cd VC[2], VN[4][2];
  ld256_keep(a.in + site * NC, VC[0], VC[1]);
  ld256_keep(s0, VN[0][0], VN[0][1]);
  ld256_keep(s1, VN[1][0], VN[1][1]);
  ld256_keep(s2, VN[2][0], VN[2][1]);
  ld256_keep(s3, VN[3][0], VN[3][1]);
  cd OLD[2] = {zero, zero}, RB[2] = {zero, zero};
  if (a.accumulate) { OLD[0] = a.out[site * NC]; OLD[1] = a.out[site * NC + 1]; }
  if (RESID) ld256_stream(a.resid + site * NC, RB[0], RB[1]);
  u[2] = cconj(u[2]);
  u[3] = cconj(u[3]);

  cd res[2];
#pragma unroll
  for (int c1 = 0; c1 < 2; c1++)
  {
    cd part[2];
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++)
    {
      // the chain of the element kernel's lane (c1, c2): clover, diagonal shift, +x, +y, -x, -y
      const cd CL = (c1 == c2) ? cmake(2.0 * a.mf_w, 0.0) : zero;
      const cd DG = (a.use_diag && c1 == c2) ? a.diag[p][c2] : zero;
      cd acc = zero;
      cfma(acc, CL, VC[c2]);
      cfma(acc, DG, VC[c2]);
#pragma unroll
      for (int mu = 0; mu < 4; mu++) cfma(acc, wilson_hop_element(mu, c1, c2, u[mu], a.mf_w), VN[mu][c2]);
      part[c2] = acc;
    }
    cd r = cadd(part[0], part[1]);
    r = cadd(r, OLD[c1]);
    if (RESID) r = csub(RB[c1], r);
    res[c1] = r;
  }
  st256(a.out + site * NC, res[0], res[1]);
}

// The same arithmetic out of a shared-memory patch.  The kernel above reads 256 bytes per site through L2 (own and four
// neighbour spinors, four links, the output) for its 96 bytes of DRAM traffic and sits on the L2 -> SM fabric (9.3 TB/s at
// 1.84 ms on 8192^2), not on DRAM.  Here a CTA owns TY rows x 2 TK columns (both parities), every thread loads the spinor and
// the two links of ITS site once -- coalesced runs of TK sites -- plus one element of the patch's ring (spinors: one column
// left and right, one row above and below; U_x: the column to the left; U_y: the row below), and the five spinors and four
// links of a site come out of shared memory: ~115 bytes per site cross the fabric.  Whole patches only (rows and columns of
// the launch divisible by the patch, both parities, stride 1): everything else takes the kernel above.
template <int TK, int TY, bool RESID>
__global__ void __launch_bounds__(TK * TY * 2, 4) wilson_mf_tile_kernel(const StencilKArgs a)
{
  constexpr int NC = 2, VK = TK + 2, UK = TK + 1, NT = TK * TY * 2;
  // the two spinor components sit in planes of their own: consecutive threads then touch consecutive 16-byte words (as [..][c]
  // pairs every access was a two-way bank conflict: 94 M conflicts in 227 M wavefronts, L1 at 87 %, profiles/r06l_ncu_...)
  constexpr int PLANE = (TY + 2) * 2 * VK;
  __shared__ __align__(16) cd sV[2 * PLANE];                    // [c][ry = row + 1][parity][kk = k - k0 + 1]
  __shared__ __align__(16) cd sUx[TY * 2 * UK];                // [row][parity][kk = k - k0 + 1]  (kk = 0: the column to the left)
  __shared__ __align__(16) cd sUy[(TY + 1) * 2 * TK];          // [ry = row + 1][parity][k - k0]  (ry = 0: the row below)
  const int tid = threadIdx.x;
  const int tk = tid % TK, p = (tid / TK) & 1, ty = tid / (2 * TK);
  const int k0 = blockIdx.x * TK, y0 = a.y_off + blockIdx.y * TY;
  const int xh = a.g.xh, Y = a.g.Y;
  const size_t half = a.g.half, V = 2 * half;
  const int y = y0 + ty, k = k0 + tk;
  const size_t site = (size_t)p * half + (size_t)y * xh + k;
  const cd zero = cmake(0.0, 0.0);

  // own site: spinor, U_x, U_y
  cd v0, v1;
  ld256_keep(a.in + site * NC, v0, v1);
  const cd ux = ld_stream(a.mf_gauge + site), uy = ld_stream(a.mf_gauge + V + site);
  cd OLD[2] = {zero, zero}, RB[2] = {zero, zero};
  if (a.accumulate) { OLD[0] = a.out[site * NC]; OLD[1] = a.out[site * NC + 1]; }
  if (RESID) ld256_stream(a.resid + site * NC, RB[0], RB[1]);
  // the ring: spinor columns left / right (TY x 2 x 2), spinor rows below / above (2 x 2 x TK), U_x column left (TY x 2), U_y row below (2 x TK);
  // one element per thread (the patch has more sites than ring elements), its load issued together with the thread's own
  constexpr int N_VCOL = TY * 2 * 2, N_VROW = 2 * 2 * TK, N_UX = TY * 2, N_UY = 2 * TK;
  static_assert(N_VCOL + N_VROW + N_UX + N_UY <= NT, "wilson_mf_tile_kernel: one ring element per thread");
  const int kl = (k0 == 0) ? xh - 1 : k0 - 1, kr = (k0 + TK == xh) ? 0 : k0 + TK;
  const int yb = (y0 == 0) ? Y - 1 : y0 - 1, yt = (y0 + TY == Y) ? 0 : y0 + TY;
  const bool halo_b = (a.halo_ym != nullptr && y0 == 0), halo_t = (a.halo_yp != nullptr && y0 + TY == Y);
  cd h0 = zero, h1 = zero;
  cd* hdst = nullptr;          // shared-memory destination of this thread's ring element
  bool hpair = false;          // two complex (a spinor) or one (a link)
  {
    int j = tid;
    if (j < N_VCOL)
    {
      const int side = j & 1, pp = (j >> 1) & 1, r = j >> 2;
      ld256_keep(a.in + ((size_t)pp * half + (size_t)(y0 + r) * xh + (side ? kr : kl)) * NC, h0, h1);
      hdst = sV + ((r + 1) * 2 + pp) * VK + (side ? TK + 1 : 0); hpair = true;
    }
    else if ((j -= N_VCOL) < N_VROW)
    {
      const int kk = j % TK, pp = (j / TK) & 1, top = j / (2 * TK);
      const cd* src;
      if (top) src = halo_t ? a.halo_yp + ((size_t)pp * xh + k0 + kk) * NC : a.in + ((size_t)pp * half + (size_t)yt * xh + k0 + kk) * NC;
      else src = halo_b ? a.halo_ym + ((size_t)pp * xh + k0 + kk) * NC : a.in + ((size_t)pp * half + (size_t)yb * xh + k0 + kk) * NC;
      ld256_keep(src, h0, h1);
      hdst = sV + ((top ? TY + 1 : 0) * 2 + pp) * VK + kk + 1; hpair = true;
    }
    else if ((j -= N_VROW) < N_UX)
    {
      const int pp = j & 1, r = j >> 1;
      h0 = ld_keep(a.mf_gauge + (size_t)pp * half + (size_t)(y0 + r) * xh + kl);
      hdst = sUx + (r * 2 + pp) * UK;
    }
    else if ((j -= N_UX) < N_UY)
    {
      const int kk = j % TK, pp = j / TK;
      h0 = ld_keep((a.mf_gauge_ym != nullptr && y0 == 0) ? a.mf_gauge_ym + ((size_t)pp * xh + k0 + kk) : a.mf_gauge + V + (size_t)pp * half + (size_t)yb * xh + k0 + kk);
      hdst = sUy + pp * TK + kk;
    }
  }
  {
    cd* d = sV + ((ty + 1) * 2 + p) * VK + tk + 1;
    d[0] = v0; d[PLANE] = v1;
    sUx[(ty * 2 + p) * UK + tk + 1] = ux;
    sUy[((ty + 1) * 2 + p) * TK + tk] = uy;
    if (hdst != nullptr) { hdst[0] = h0; if (hpair) hdst[PLANE] = h1; }
  }
  __syncthreads();

  // the chains of the element kernel's lanes (c1, c2) -- clover, diagonal shift, +x, +y, -x, -y -- advanced direction by direction,
  // each direction's link and spinor taken from shared memory when its turn comes (the patch is small: registers, not
  // shared-memory latency, decide how many CTAs an SM holds)
  const int q = 1 - p, sft = (y + p) & 1;
  cd acc[2][2];
#pragma unroll
  for (int c1 = 0; c1 < 2; c1++)
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++)
    {
      const cd vc = c2 ? v1 : v0;
      const cd CL = (c1 == c2) ? cmake(2.0 * a.mf_w, 0.0) : zero;
      const cd DG = (a.use_diag && c1 == c2) ? a.diag[p][c2] : zero;
      cd t = zero;
      cfma(t, CL, vc);
      cfma(t, DG, vc);
      acc[c1][c2] = t;
    }
  const cd* rowq = sV + ((ty + 1) * 2 + q) * VK;
#pragma unroll
  for (int mu = 0; mu < 4; mu++)
  {
    cd um;
    const cd* vn;
    if (mu == 0) { um = ux; vn = rowq + (tk + 1 + sft); }                                                        // (q, y, k + sft)
    else if (mu == 1) { um = uy; vn = sV + ((ty + 2) * 2 + q) * VK + tk + 1; }                                  // (q, y + 1, k)
    else if (mu == 2) { um = cconj(sUx[(ty * 2 + q) * UK + tk + sft]); vn = rowq + (tk + sft); }                  // (q, y, k - 1 + sft) and its U_x
    else { um = cconj(sUy[(ty * 2 + q) * TK + tk]); vn = sV + (ty * 2 + q) * VK + tk + 1; }                     // (q, y - 1, k) and its U_y
    const cd n0 = vn[0], n1 = vn[PLANE];
#pragma unroll
    for (int c1 = 0; c1 < 2; c1++)
    {
      cfma(acc[c1][0], wilson_hop_element(mu, c1, 0, um, a.mf_w), n0);
      cfma(acc[c1][1], wilson_hop_element(mu, c1, 1, um, a.mf_w), n1);
    }
  }
  cd res[2];
#pragma unroll
  for (int c1 = 0; c1 < 2; c1++)
  {
    cd r = cadd(acc[c1][0], acc[c1][1]);
    r = cadd(r, OLD[c1]);
    if (RESID) r = csub(RB[c1], r);
    res[c1] = r;
  }
  st256(a.out + site * NC, res[0], res[1]);
}

static int wilson_mf_tile_mode() { static int v = -1; if (v < 0) { const char* e = getenv("QMG_MF_TILE"); v = (e != nullptr && e[0] == '0') ? 0 : 1; } return v; }

static int launch_wilson_mf(const StencilKArgs& a, int n_par)
{
  constexpr int MTK = 16, MTY = 8;      // (apply_sharded cuts a slab into bands of this height)
  if (wilson_mf_tile_mode() && n_par == 2 && a.y_stride == 1 && a.g.xh % MTK == 0 && a.y_cnt % MTY == 0 && a.y_cnt / MTY <= 65535)
  {
    dim3 grid(a.g.xh / MTK, a.y_cnt / MTY, 1);
    if (a.resid != nullptr) wilson_mf_tile_kernel<MTK, MTY, true><<<grid, MTK * MTY * 2, 0, rt().stream>>>(a);
    else wilson_mf_tile_kernel<MTK, MTY, false><<<grid, MTK * MTY * 2, 0, rt().stream>>>(a);
    QMG_LAUNCH_CHECK();
    return 0;
  }
  const int row_threads = a.g.xh;
  int bx = 256;
  while (bx > 32 && bx / 2 >= row_threads) bx /= 2;
  int by = 256 / bx;
  if (by > a.y_cnt) by = a.y_cnt;
  dim3 block(bx, by, 1);
  dim3 grid(((row_threads + bx - 1) / bx) * n_par, (a.y_cnt + by - 1) / by, 1);
  if (grid.y > 65535) return fail_msg("qmg_stencil_apply: Y too large for the launch grid");
  if (a.resid != nullptr) wilson_mf_kernel<true><<<grid, block, 0, rt().stream>>>(a);
  else wilson_mf_kernel<false><<<grid, block, 0, rt().stream>>>(a);
  QMG_LAUNCH_CHECK();
  return 0;
}

template <int NC>
static int launch_stencil(const StencilKArgs& a, int n_par, bool reduce)
{
  Runtime& r = rt();
  const int row_elems = a.g.xh * NC * NC;
  int bx = 256;
  while (bx > 32 && bx / 2 >= row_elems) bx /= 2;
  int by = 256 / bx;
  if (by > a.y_cnt) by = a.y_cnt;
  dim3 block(bx, by, 1);
  dim3 grid(((row_elems + bx - 1) / bx) * n_par, (a.y_cnt + by - 1) / by, 1);
  if (grid.y > 65535) return fail_msg("qmg_stencil_apply: Y too large for the launch grid");
  if (reduce)
  {
    double* partials = ensure_partials((size_t)grid.x * grid.y * grid.z * 3);
    if (partials == nullptr) return 1;
    stencil_kernel<NC, true, false, false><<<grid, block, 0, r.stream>>>(a, partials, r.d_counter, r.d_result);
  }
  else if (a.herm && NC > 1)
  {
    if (a.resid != nullptr) stencil_kernel<NC, false, (NC > 1), true><<<grid, block, 0, r.stream>>>(a, nullptr, nullptr, nullptr);
    else stencil_kernel<NC, false, (NC > 1), false><<<grid, block, 0, r.stream>>>(a, nullptr, nullptr, nullptr);
  }
  else if (a.resid != nullptr)
    stencil_kernel<NC, false, false, true><<<grid, block, 0, r.stream>>>(a, nullptr, nullptr, nullptr);
  else
    stencil_kernel<NC, false, false, false><<<grid, block, 0, r.stream>>>(a, nullptr, nullptr, nullptr);
  QMG_LAUNCH_CHECK();
  return 0;
}

static int dispatch_stencil(const StencilKArgs& a, int nc, int n_par, bool reduce)
{
  if (!reduce && a.mf_gauge != nullptr) return launch_wilson_mf(a, n_par);
  if (!reduce && a.herm && rt().tile_kernel)
  {
    // patch shapes: nc = 8: 8 x 4 sites (97 KB of shared memory, two CTAs per SM; 4x4, 8x2, 4x2 and 16x4 patches measured
    // 3-23 % slower, profiles/r02q_tile_shapes.txt); nc = 4: 16 x 8.  nc = 2 blocks are too small to win (the
    // column-wise clover loads waste half of every sector): the fine level keeps the streaming kernel.
    // nc = 8: two threads per (site, column), four rows each -- 32 instead of 16 warps per SM on the same staged bytes:
    // 3.00 -> 2.51 ms sustained on 2048^2 (profiles/r02w_tile_split.txt; QMG_TILE=2 selects the one-thread flavour)
    if (nc == 8 && tile_applicable<4, 4>(a, n_par)) return launch_tile8(a);
    if (nc == 4 && tile_applicable<8, 8>(a, n_par)) return launch_tile<4, 8, 8>(a);
  }
  switch (nc)
  {
    case 1: return launch_stencil<1>(a, n_par, reduce);
    case 2: return launch_stencil<2>(a, n_par, reduce);
    case 4: return launch_stencil<4>(a, n_par, reduce);
    case 8: return launch_stencil<8>(a, n_par, reduce);
    case 16: return launch_stencil<16>(a, n_par, reduce);
    case 32: return launch_stencil<32>(a, n_par, reduce);
    default: break;
  }
  if (reduce) return fail_msg("qmg_stencil_apply_dot: fused reduction needs nc in {1,2,4,8,16,32}");
  if (a.herm) return fail_msg("qmg_stencil_apply: the gamma5-hermitian apply needs nc in {2,4,8,16,32}");
  Runtime& r = rt();
  const long rows = (long)n_par * a.y_cnt * a.g.xh * nc;
  if (nc >= 16)
  {
    long blocks = (rows * 8 + 255) / 256, cap = (long)r.sm_count * 8;
    stencil_kernel_generic<8><<<(int)(blocks < cap ? blocks : cap), 256, 0, r.stream>>>(a, nc, n_par);
  }
  else
  {
    long blocks = (rows + 255) / 256, cap = (long)r.sm_count * 8;
    stencil_kernel_rows<<<(int)(blocks < cap ? blocks : cap), 256, 0, r.stream>>>(a, nc, n_par);
  }
  QMG_LAUNCH_CHECK();
  return 0;
}

// Does this apply read rows of rhs that live on the ring neighbours?  (sharded, a hop in +-y, no caller-supplied rows)
static bool needs_exchange(const StencilKArgs& a)
{
  return comm().active && a.hop != nullptr && (a.dir_mask & 10) != 0 && a.halo_ym == nullptr && a.halo_yp == nullptr &&
         (a.hop_to[0] || a.hop_to[1]) && a.g.half > 1;
}
// parities of rhs whose boundary rows are read: writing parity p hops from parity 1 - p
static int exchange_parities(const StencilKArgs& a, int n_par)
{
  int m = 0;
  for (int i = 0; i < n_par; i++) { const int p = a.p_begin + i; if (a.hop_to[p]) m |= 1 << (1 - p); }
  return m;
}

// Sharded apply: the two boundary rows of rhs travel on the exchange stream while the interior rows are computed;
// rows 0 and Y-1 follow once the neighbours' rows have arrived.
static int apply_sharded(StencilKArgs& a, int nc, int n_par)
{
  HaloRows rows;
  int rc = halo_exchange_begin(a.in, 2 * a.g.xh, a.g.Y, nc, exchange_parities(a, n_par), &rows);
  if (rc) return rc;
  // gamma5-hermitian link set: the rows that touch no other slab go through the shared-memory tile kernel, one band of
  // patch height at either end through the streaming kernel once the neighbours' rows have arrived
  constexpr int TY8 = 4, TK8 = 4;
  if (nc == 8 && rt().tile_kernel && tile_shape_fits<TK8, TY8>(a, n_par) && a.g.Y >= 3 * TY8)
  {
    StencilKArgs in = a;
    in.hop_ym = nullptr; in.y_off = TY8; in.y_stride = 1; in.y_cnt = a.g.Y - 2 * TY8;
    rc = launch_tile8(in); if (rc) return rc;
    rc = halo_exchange_end(); if (rc) return rc;
    a.halo_ym = rows.ym; a.halo_yp = rows.yp;
    a.y_off = 0; a.y_stride = 1; a.y_cnt = TY8;
    rc = dispatch_stencil(a, nc, n_par, false); if (rc) return rc;
    a.y_off = a.g.Y - TY8;
    return dispatch_stencil(a, nc, n_par, false);
  }
  // matrix-free Wilson set: whole patches of the patch kernel inside, one patch-high band at either end once the rows have arrived
  constexpr int MFY = 8, MFK = 16;
  if (a.mf_gauge != nullptr && n_par == 2 && a.g.xh % MFK == 0 && a.g.Y % MFY == 0 && a.g.Y >= 3 * MFY)
  {
    StencilKArgs in = a;
    in.y_off = MFY; in.y_stride = 1; in.y_cnt = a.g.Y - 2 * MFY;
    rc = dispatch_stencil(in, nc, n_par, false); if (rc) return rc;
    rc = halo_exchange_end(); if (rc) return rc;
    a.halo_ym = rows.ym; a.halo_yp = rows.yp;
    a.y_off = 0; a.y_stride = 1; a.y_cnt = MFY;
    rc = dispatch_stencil(a, nc, n_par, false); if (rc) return rc;
    a.y_off = a.g.Y - MFY;
    return dispatch_stencil(a, nc, n_par, false);
  }
  if (a.g.Y > 2)
  {
    a.y_off = 1; a.y_stride = 1; a.y_cnt = a.g.Y - 2;
    rc = dispatch_stencil(a, nc, n_par, false);
    if (rc) return rc;
  }
  rc = halo_exchange_end(); if (rc) return rc;
  a.halo_ym = rows.ym; a.halo_yp = rows.yp;
  a.y_off = 0; a.y_stride = a.g.Y - 1; a.y_cnt = 2;
  return dispatch_stencil(a, nc, n_par, false);
}

// ---- host-vector apply: upload, compute and download pipelined over row chunks -------------------------------------------
// The reference-facing call for callers whose vectors live in HOST memory (the reference's own drivers): rhs goes up
// chunk by chunk on one copy stream, each chunk of output rows is computed as soon as the rows it reads have landed, and
// comes down on a second copy stream, so PCIe carries traffic in both directions while the SMs work.
struct HostPipe
{
  cudaStream_t up = nullptr, down = nullptr;
  static const int kMaxChunks = 64;
  cudaEvent_t uploaded[kMaxChunks + 1], computed[kMaxChunks], start = nullptr;
  bool ready = false;
};
static HostPipe& host_pipe() { static HostPipe h; return h; }

static int host_pipe_init()
{
  HostPipe& h = host_pipe();
  if (h.ready) return 0;
  QMG_CUDA(cudaStreamCreateWithFlags(&h.up, cudaStreamNonBlocking));
  QMG_CUDA(cudaStreamCreateWithFlags(&h.down, cudaStreamNonBlocking));
  for (int i = 0; i <= HostPipe::kMaxChunks; i++) QMG_CUDA(cudaEventCreateWithFlags(&h.uploaded[i], cudaEventDisableTiming));
  for (int i = 0; i < HostPipe::kMaxChunks; i++) QMG_CUDA(cudaEventCreateWithFlags(&h.computed[i], cudaEventDisableTiming));
  QMG_CUDA(cudaEventCreateWithFlags(&h.start, cudaEventDisableTiming));
  h.ready = true;
  return 0;
}

// rows [y0, y0 + cnt) of both parity halves of an (parity, y, x/2, nc) vector
static int copy_rows(cd* dst, const cd* src, const Geom& g, int nc, int y0, int cnt, int parity_mask, cudaMemcpyKind kind, cudaStream_t s)
{
  const size_t rowlen = (size_t)g.xh * nc;
  for (int p = 0; p < 2; p++)
  {
    if (!((parity_mask >> p) & 1)) continue;
    const size_t off = ((size_t)p * g.Y + y0) * rowlen;
    QMG_CUDA(cudaMemcpyAsync(dst + off, src + off, sizeof(cd) * rowlen * cnt, kind, s));
  }
  return 0;
}

static int apply_host_pipelined(StencilKArgs& a, int nc, int n_par, cd* lhs_host, const cd* rhs_host, int rows_per_chunk)
{
  int rc = host_pipe_init(); if (rc) return rc;
  HostPipe& h = host_pipe();
  Runtime& r = rt();
  const int Y = a.g.Y;
  if (rows_per_chunk < 2) rows_per_chunk = 2;
  int nch = (Y + rows_per_chunk - 1) / rows_per_chunk;
  if (nch > HostPipe::kMaxChunks) { nch = HostPipe::kMaxChunks; rows_per_chunk = (Y + nch - 1) / nch; nch = (Y + rows_per_chunk - 1) / rows_per_chunk; }
  const int out_mask = (n_par == 2) ? 3 : (1 << a.p_begin);
  cd* d_in = const_cast<cd*>(a.in);
  const bool sharded = needs_exchange(a);
  // both copy streams start after whatever the caller queued before
  QMG_CUDA(cudaEventRecord(h.start, r.stream));
  QMG_CUDA(cudaStreamWaitEvent(h.up, h.start, 0));
  QMG_CUDA(cudaStreamWaitEvent(h.down, h.start, 0));
  // the two edge rows go first: row Y-1 is what row 0 reads across the periodic wrap, and on a y-slab both are what the
  // ring neighbours need, so the halo exchange runs while the bulk of rhs is still uploading
  rc = copy_rows(d_in, rhs_host, a.g, nc, 0, 1, 3, cudaMemcpyHostToDevice, h.up); if (rc) return rc;
  rc = copy_rows(d_in, rhs_host, a.g, nc, Y - 1, 1, 3, cudaMemcpyHostToDevice, h.up); if (rc) return rc;
  QMG_CUDA(cudaEventRecord(h.uploaded[HostPipe::kMaxChunks], h.up));
  if (a.accumulate) { rc = copy_rows(a.out, lhs_host, a.g, nc, 0, Y, out_mask, cudaMemcpyHostToDevice, h.up); if (rc) return rc; }
  for (int c = 0; c < nch; c++)
  {
    // rows 0 and Y-1 are already there; never rewrite a row a kernel may be reading
    int y0 = c * rows_per_chunk, y1 = (y0 + rows_per_chunk <= Y) ? y0 + rows_per_chunk : Y;
    if (c == 0) y0 = 1;
    if (y1 == Y) y1 = Y - 1;
    if (y1 > y0) { rc = copy_rows(d_in, rhs_host, a.g, nc, y0, y1 - y0, 3, cudaMemcpyHostToDevice, h.up); if (rc) return rc; }
    QMG_CUDA(cudaEventRecord(h.uploaded[c], h.up));
  }
  if (sharded)
  {
    HaloRows rows;
    QMG_CUDA(cudaStreamWaitEvent(r.stream, h.uploaded[HostPipe::kMaxChunks], 0));
    rc = halo_exchange_begin(a.in, 2 * a.g.xh, Y, nc, exchange_parities(a, n_par), &rows); if (rc) return rc;
    rc = halo_exchange_end(); if (rc) return rc;
    a.halo_ym = rows.ym; a.halo_yp = rows.yp;
  }
  for (int c = 0; c < nch; c++)
  {
    const int y0 = c * rows_per_chunk, cnt = (y0 + rows_per_chunk <= Y) ? rows_per_chunk : Y - y0;
    // chunk c reads rows y0-1 .. y0+cnt: everything up to chunk c+1 (the last chunk wraps to row 0, long since there)
    QMG_CUDA(cudaStreamWaitEvent(r.stream, h.uploaded[c + 1 < nch ? c + 1 : c], 0));
    a.y_off = y0; a.y_stride = 1; a.y_cnt = cnt;
    rc = dispatch_stencil(a, nc, n_par, false); if (rc) return rc;
    QMG_CUDA(cudaEventRecord(h.computed[c], r.stream));
    QMG_CUDA(cudaStreamWaitEvent(h.down, h.computed[c], 0));
    rc = copy_rows(lhs_host, a.out, a.g, nc, y0, cnt, out_mask, cudaMemcpyDeviceToHost, h.down); if (rc) return rc;
  }
  QMG_CUDA(cudaEventRecord(h.start, h.down));
  QMG_CUDA(cudaStreamWaitEvent(r.stream, h.start, 0));
  QMG_CUDA(cudaStreamSynchronize(r.stream));
  return 0;
}

} // namespace qmg

using namespace qmg;

extern "C" {

int qmg_stencil_apply_host(const qmg_stencil_desc* st, int pieces, int dir_mask, qmg_cplx* lhs_host, const qmg_cplx* rhs_host,
                           qmg_cplx* dev_lhs, qmg_cplx* dev_rhs, int rows_per_chunk)
{
  QMG_REQUIRE_INIT();
  if (st == nullptr) return fail_msg("qmg_stencil_apply_host: null stencil");
  if (lhs_host == nullptr || rhs_host == nullptr) return fail_msg("qmg_stencil_apply_host: null host vector");
  const size_t bytes = sizeof(cd) * (size_t)st->X * st->Y * st->nc;
  void* own_l = nullptr; void* own_r = nullptr;
  int rc = 0;
  if (dev_lhs == nullptr) { rc = qmg_malloc(&own_l, bytes); if (rc) return rc; dev_lhs = (qmg_cplx*)own_l; }
  if (dev_rhs == nullptr) { rc = qmg_malloc(&own_r, bytes); if (rc) { qmg_free(own_l); return rc; } dev_rhs = (qmg_cplx*)own_r; }
  StencilKArgs a; int n_par;
  rc = build_args(st, pieces, dir_mask, dev_lhs, dev_rhs, a, n_par);
  if (!rc)
  {
    const bool half_vectors = (pieces & (QMG_APPLY_EVEN_ROWS_ONLY | QMG_APPLY_ODD_ROWS_ONLY)) != 0;
    if (half_vectors || a.g.Y < 4 || st->halo_ym != nullptr || st->halo_yp != nullptr)
    {
      // partial applies take the plain route: whole vector up, apply, whole vector down
      rc = qmg_memcpy_h2d(dev_rhs, rhs_host, bytes);
      if (!rc && (pieces & QMG_APPLY_ACCUMULATE)) rc = qmg_memcpy_h2d(dev_lhs, lhs_host, bytes);
      if (!rc) rc = qmg_stencil_apply(st, pieces, dir_mask, dev_lhs, dev_rhs);
      if (!rc) rc = qmg_memcpy_d2h(lhs_host, dev_lhs, bytes);
    }
    else
    {
      if (rows_per_chunk <= 0)
      {
        // ~64 MB of rhs per chunk: long enough for PCIe to run at full rate, short enough that the exposed head and tail
        // (first upload, last download) stay a few per cent of the transfer
        const size_t row_bytes = sizeof(cd) * (size_t)st->X * st->nc;
        rows_per_chunk = (int)((64u << 20) / row_bytes);
        if (rows_per_chunk < 2) rows_per_chunk = 2;
      }
      rc = apply_host_pipelined(a, st->nc, n_par, reinterpret_cast<cd*>(lhs_host), reinterpret_cast<const cd*>(rhs_host), rows_per_chunk);
    }
  }
  qmg_free(own_l); qmg_free(own_r);
  return rc;
}

int qmg_stencil_apply(const qmg_stencil_desc* st, int pieces, int dir_mask, qmg_cplx* lhs, const qmg_cplx* rhs)
{
  QMG_REQUIRE_INIT();
  if (st != nullptr) prof_scope__.detail(st->nc, st->X, st->Y);
  StencilKArgs a; int n_par;
  int rc = build_args(st, pieces, dir_mask, lhs, rhs, a, n_par);
  if (rc) return rc;
  if ((const void*)lhs == (const void*)rhs && n_par == 2)
  {
    // in-place hopping into BOTH parities: one launch would read rows of parity 1 - p while other blocks rewrite them.
    // The reference is sequential -- apply_M_eo, then apply_M_oe on the updated even rows (stencil_2d.h:843-850) -- so
    // the in-place case is two ordered launches with exactly that meaning.
    rc = qmg_stencil_apply(st, (pieces & ~QMG_APPLY_ODD_ROWS_ONLY & ~QMG_APPLY_HOP_TO_ODD) | QMG_APPLY_EVEN_ROWS_ONLY, dir_mask, lhs, rhs);
    if (rc) return rc;
    return qmg_stencil_apply(st, (pieces & ~QMG_APPLY_EVEN_ROWS_ONLY & ~QMG_APPLY_HOP_TO_EVEN) | QMG_APPLY_ODD_ROWS_ONLY, dir_mask, lhs, rhs);
  }
  if (needs_exchange(a)) return apply_sharded(a, st->nc, n_par);
  return dispatch_stencil(a, st->nc, n_par, false);
}

int qmg_stencil_apply_residual(const qmg_stencil_desc* st, int pieces, int dir_mask, qmg_cplx* lhs, const qmg_cplx* rhs, const qmg_cplx* b)
{
  QMG_REQUIRE_INIT();
  if (st != nullptr) prof_scope__.detail(st->nc, st->X, st->Y);
  if (b == nullptr) return fail_msg("qmg_stencil_apply_residual: null right-hand side");
  if (pieces & QMG_APPLY_ACCUMULATE) return fail_msg("qmg_stencil_apply_residual: the residual epilogue is out-of-place (no QMG_APPLY_ACCUMULATE)");
  if ((const void*)lhs == (const void*)rhs) return fail_msg("qmg_stencil_apply_residual: lhs must not alias rhs (it may alias b)");
  StencilKArgs a; int n_par;
  int rc = build_args(st, pieces, dir_mask, lhs, rhs, a, n_par);
  if (rc) return rc;
  a.resid = reinterpret_cast<const cd*>(b);
  if (needs_exchange(a)) return apply_sharded(a, st->nc, n_par);
  return dispatch_stencil(a, st->nc, n_par, false);
}

int qmg_stencil_apply_dot(const qmg_stencil_desc* st, int pieces, qmg_cplx* lhs, const qmg_cplx* rhs, const qmg_cplx* dot_with, double* result3)
{
  QMG_REQUIRE_INIT();
  StencilKArgs a; int n_par;
  int rc = build_args(st, pieces, 15, lhs, rhs, a, n_par);
  if (rc) return rc;
  a.dotw = reinterpret_cast<const cd*>(dot_with);
  a.herm = 0;        // the fused-reduction flavour always reads the stored backward blocks
  if (needs_exchange(a))
  {
    // the fused reduction finishes in ONE launch (last-block tail), so the rows are fetched first, without overlap
    HaloRows rows;
    rc = halo_exchange_begin(a.in, st->X, st->Y, st->nc, exchange_parities(a, n_par), &rows); if (rc) return rc;
    rc = halo_exchange_end(); if (rc) return rc;
    a.halo_ym = rows.ym; a.halo_yp = rows.yp;
  }
  rc = dispatch_stencil(a, st->nc, n_par, true);
  if (rc) return rc;
  return fetch_result(result3, 3);
}

} // extern "C"
