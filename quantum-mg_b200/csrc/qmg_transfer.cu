// K4-K7: TransferMG prolong / restrict, per-aggregate block orthonormalisation and the
// Galerkin coarse-operator build (/root/reference/transfer/transfer.h:455-607,
// /root/reference/operators/coarse.h:90-471).
//
// The reference stores, per coarse site, a sorted list of the fine colour-vector indices of its
// aggregate (transfer.h:410-448).  For the regular non-overlapping blocking it supports that list
// is pure arithmetic on the even-odd layout, so no map is stored here: coarse site (xc, yc) owns
// fine sites x in [xc bx, (xc+1) bx), y in [yc by, (yc+1) by), and inside one fine row and one
// parity those sites are contiguous in memory ((bx/2) ncf complex numbers when bx is even).
//
// All four kernels are HBM-bound streaming passes over the null vectors
// (16 (N_f (nvec + 1) + N_c) algorithmic bytes per prolong / restrict).
#include "qmg_comm.cuh"

namespace qmg {

struct TGeom
{
  int Xf, Yf, ncf, Xc, Yc, ncc;
  int bx, by;          // block sizes
  int xhf;             // Xf/2
  int fspc;            // fine dof per aggregate = bx by ncf
  int seg;             // contiguous complex per (row, parity) of an aggregate = (bx/2) ncf   (bx even)
  int even_bx;
  long Vc;
};

static int make_geom(const qmg_transfer_desc* t, TGeom& g, const char* who)
{
  if (t == nullptr) return fail_msg("transfer: null descriptor");
  g.Xf = t->Xf; g.Yf = t->Yf; g.ncf = t->ncf; g.Xc = t->Xc; g.Yc = t->Yc; g.ncc = t->ncc;
  if (g.Xf < 2 || g.Yf < 2 || (g.Xf & 1) || (g.Yf & 1)) return fail_msg("transfer: fine lattice dimensions must be even and >= 2");
  if (g.Xc < 1 || g.Yc < 1 || g.Xf % g.Xc || g.Yf % g.Yc) return fail_msg("transfer: fine lattice dimension isn't divided evenly by coarse dimension");
  if (!((g.Xc == 1 && g.Yc == 1) || (!(g.Xc & 1) && !(g.Yc & 1)))) return fail_msg("transfer: coarse lattice must have even dimensions or be a single site");
  if (g.ncf < 1 || g.ncc < 1) return fail_msg("transfer: nc < 1");
  g.bx = g.Xf / g.Xc; g.by = g.Yf / g.Yc; g.xhf = g.Xf / 2;
  g.fspc = g.bx * g.by * g.ncf;
  g.even_bx = (g.bx & 1) ? 0 : 1;
  g.seg = (g.bx / 2) * g.ncf;
  g.Vc = (long)g.Xc * g.Yc;
  (void)who;
  return 0;
}

// coarse site index of coarse coordinates (lattice.h:75-81)
__host__ __device__ __forceinline__ long coarse_index(const TGeom& g, int xc, int yc)
{
  if (g.Vc == 1) return 0;
  const int par = (xc + yc) & 1;
  return (long)(yc + par * g.Yc) * (g.Xc / 2) + (xc / 2);
}

// fine colour-vector index of element e of aggregate (xc, yc).
// even bx: e = (row, parity, j) with j running over the contiguous segment; odd bx: e = (row, x, c).
__device__ __forceinline__ long agg_elem_index(const TGeom& g, int xc, int yc, int e)
{
  if (g.even_bx)
  {
    const int rowlen = 2 * g.seg;
    const int r = e / rowlen, rem = e - r * rowlen;
    const int p = rem / g.seg, j = rem - p * g.seg;
    const int y = yc * g.by + r;
    return ((long)(y + p * g.Yf) * g.xhf + xc * (g.bx / 2)) * g.ncf + j;
  }
  const int rowlen = g.bx * g.ncf;
  const int r = e / rowlen, rem = e - r * rowlen;
  const int xi = rem / g.ncf, c = rem - xi * g.ncf;
  const int x = xc * g.bx + xi, y = yc * g.by + r;
  const int p = (x + y) & 1;
  return ((long)(y + p * g.Yf) * g.xhf + (x >> 1)) * g.ncf + c;
}

template <int NV> struct VecPack { const cd* p[NV]; };
template <int NV> struct VecPackRW { cd* p[NV]; };

// ------------------------------------------------------------------ restrict --
// G lanes (a power of two <= 32) own one aggregate; each lane strides over the aggregate's
// elements accumulating NV partial dot products, then a transposing butterfly leaves every
// fully reduced value on exactly one lane: step s exchanges half of the still-live values, so
// NV values over 32 lanes cost NV - 1 + (5 - log2 NV) shuffles instead of 5 NV.
template <int NV>
__global__ void __launch_bounds__(256) restrict_kernel(const TGeom g, const VecPack<NV> nv, const int nv_count, const int v0,
                                                       const cd* __restrict__ fine, cd* __restrict__ coarse, const int G, const int logG,
                                                       const int overwrite)
{
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long agg = gt >> logG;             // aggregates are enumerated row-major in (yc, xc): neighbours in memory are neighbours in the grid
  const int lane_in = (int)(gt & (G - 1));
  const bool live = agg < g.Vc;
  const int yc = live ? (int)(agg / g.Xc) : 0, xc = live ? (int)(agg - (long)yc * g.Xc) : 0;

  cd acc[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) acc[v] = cmake(0.0, 0.0);
  if (live)
  {
    // two elements per trip: 2 (NV + 1) independent 16-byte loads in flight per lane (one element per trip left the
    // kernel latency-bound: 62 % of DRAM peak at 17 % issue utilisation, profiles/r03a_ncu_restrict_reduce_tile.txt)
    int e = lane_in;
    for (; e + G < g.fspc; e += 2 * G)
    {
      const long i0 = agg_elem_index(g, xc, yc, e), i1 = agg_elem_index(g, xc, yc, e + G);
      const cd f0 = __ldg(fine + i0), f1 = __ldg(fine + i1);
      cd n0[NV], n1[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) if (v < nv_count) { n0[v] = ld_stream(nv.p[v] + i0); n1[v] = ld_stream(nv.p[v] + i1); }
#pragma unroll
      for (int v = 0; v < NV; v++) if (v < nv_count) { cfma_conj(acc[v], n0[v], f0); cfma_conj(acc[v], n1[v], f1); }
    }
    for (; e < g.fspc; e += G)
    {
      const long idx = agg_elem_index(g, xc, yc, e);
      const cd f = __ldg(fine + idx);
#pragma unroll
      for (int v = 0; v < NV; v++)
        if (v < nv_count) cfma_conj(acc[v], ld_stream(nv.p[v] + idx), f);
    }
  }

  int vbase = 0;
#pragma unroll
  for (int s = 0; s < 5; s++)
  {
    const int off = G >> (s + 1);
    if (off >= 1)
    {
      const int HALF = NV >> (s + 1);    // compile-time after unrolling
      if (HALF >= 1)
      {
        const bool upper = (lane_in & off) != 0;
#pragma unroll
        for (int i = 0; i < (NV >> 1); i++)
          if (i < HALF)
          {
            const cd send = upper ? acc[i] : acc[i + HALF];
            const cd keep = upper ? acc[i + HALF] : acc[i];
            acc[i] = cadd(keep, shfl_xor_c(send, off));
          }
        if (upper) vbase += HALF;
      }
      else
        acc[0] = cadd(acc[0], shfl_xor_c(acc[0], off));
    }
  }
  // values left per lane: max(NV >> logG, 1); plain-reduced values are replicated over the low lanes
  int logNV = 0;
  while ((1 << logNV) < NV) logNV++;
  const int left = (logG >= logNV) ? 1 : (NV >> logG);
  const int repl_mask = (logG > logNV) ? ((G >> logNV) - 1) : 0;
  if (live && (lane_in & repl_mask) == 0)
  {
    const long ci = coarse_index(g, xc, yc);
#pragma unroll
    for (int i = 0; i < NV; i++)
      if (i < left && vbase + i < nv_count)
      {
        cd* dst = coarse + ci * g.ncc + v0 + vbase + i;
        *dst = overwrite ? acc[i] : cadd(*dst, acc[i]);
      }
  }
}

// ------------------------------------------------------------------- prolong --
// One thread per fine element in memory order: fully coalesced on the null vectors and the
// fine vector; the NV coarse values of the aggregate come through L1.
template <int NV>
__global__ void __launch_bounds__(256) prolong_kernel(const TGeom g, const VecPack<NV> nv, const int nv_count, const int v0,
                                                      const cd* __restrict__ coarse, cd* __restrict__ fine,
                                                      const cd* __restrict__ base, const int use_base)
{
  const int rowlen = g.xhf * g.ncf;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= rowlen) return;
  const int y = blockIdx.y, p = blockIdx.z;
  const int k = col / g.ncf;
  const int x = 2 * k + ((y + p) & 1);
  const long ci = coarse_index(g, x / g.bx, y / g.by);
  const long idx = ((long)(y + p * g.Yf) * g.xhf) * g.ncf + col;
  // use_base: fine = base + P coarse with the sum formed from zero first (what zero + prolong + cxpyz produce, in one
  // pass); base == nullptr then means a zero base
  cd acc = use_base ? cmake(0.0, 0.0) : fine[idx];
  const cd* cv = coarse + ci * g.ncc + v0;
#pragma unroll
  for (int v = 0; v < NV; v++)
    if (v < nv_count) cfma(acc, ld_stream(nv.p[v] + idx), __ldg(cv + v));
  if (use_base && base != nullptr) acc = cadd(base[idx], acc);
  fine[idx] = acc;
}

// ---- chirality-packed null vectors -----------------------------------------------------------------------------------------
// With QMG_DOUBLE_PROJECTION (every K-cycle of the reference: tests/n13 :389, n16 :409, n19 :269, n22 :308) null vector j is the
// upper-chirality projection and vector j + ncc/2 the lower one of the same solve: at a fine element of chirality h = (c >= ncf/2)
// only the ncc/2 vectors [h ncc/2, (h + 1) ncc/2) can be non-zero, half of every prolong / restrict read is zeros.  The packed
// copy keeps the non-zero half, interleaved -- packed[idx (ncc/2) + i] = nv[h ncc/2 + i][idx] -- so a fine element's coefficients
// are ONE contiguous 16 (ncc/2)-byte run: prolong 160 -> 96, restrict 144 -> 80 bytes per fine dof.  qmg_transfer_pack_chiral
// builds it and returns the sum of |nv|^2 over the entries it drops: the caller may use the packed kernels only if that is 0.
template <int NVH>
__global__ void __launch_bounds__(256) pack_chiral_kernel(long nf, int ncf, VecPack<2 * NVH> nv, cd* __restrict__ packed, double* partials, unsigned int* counter, double* result)
{
  __shared__ double smem[8];
  double acc[1] = {0.0};
  const long stride = (long)gridDim.x * blockDim.x;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < nf; idx += stride)
  {
    const int h = ((int)(idx % ncf) >= ncf / 2) ? 1 : 0;
#pragma unroll
    for (int i = 0; i < NVH; i++)
    {
      const cd keep = h ? nv.p[NVH + i][idx] : nv.p[i][idx];
      const cd drop = h ? nv.p[i][idx] : nv.p[NVH + i][idx];
      packed[idx * NVH + i] = keep;
      acc[0] += drop.x * drop.x + drop.y * drop.y;
    }
  }
  grid_reduce_finish<1>(acc, smem, partials, counter, result);
}

// G = seg lanes per aggregate, lane j walking down the 2 by (row, parity) segments of its aggregate (element j of each): its
// dof index c = j % ncf, hence its chirality, never changes, so it accumulates the NVH sums of ITS half; the lanes of one
// chirality then add up over every lane bit but the chirality bit, and lanes 0 / (ncf/2) write the two halves of the coarse site.
template <int NVH>
__global__ void __launch_bounds__(256) restrict_packed_kernel(const TGeom g, const cd* __restrict__ packed, const cd* __restrict__ fine,
                                                              cd* __restrict__ coarse, const int G, const int logG, const int overwrite)
{
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long agg = gt >> logG;
  const int lane_in = (int)(gt & (G - 1));
  const bool live = agg < g.Vc;
  const int yc = live ? (int)(agg / g.Xc) : 0, xc = live ? (int)(agg - (long)yc * g.Xc) : 0;
  cd acc[NVH];
#pragma unroll
  for (int i = 0; i < NVH; i++) acc[i] = cmake(0.0, 0.0);
  if (live)
  {
    int e = lane_in;
    for (; e + G < g.fspc; e += 2 * G)
    {
      const long i0 = agg_elem_index(g, xc, yc, e), i1 = agg_elem_index(g, xc, yc, e + G);
      const cd f0 = __ldg(fine + i0), f1 = __ldg(fine + i1);
      cd n0[NVH], n1[NVH];
#pragma unroll
      for (int i = 0; i < NVH; i++) { n0[i] = ld_stream(packed + i0 * NVH + i); n1[i] = ld_stream(packed + i1 * NVH + i); }
#pragma unroll
      for (int i = 0; i < NVH; i++) { cfma_conj(acc[i], n0[i], f0); cfma_conj(acc[i], n1[i], f1); }
    }
    for (; e < g.fspc; e += G)
    {
      const long idx = agg_elem_index(g, xc, yc, e);
      const cd f = __ldg(fine + idx);
#pragma unroll
      for (int i = 0; i < NVH; i++) cfma_conj(acc[i], ld_stream(packed + idx * NVH + i), f);
    }
  }
  const int cb = g.ncf / 2;       // the lane bit that tells the chirality
  for (int off = G >> 1; off >= 1; off >>= 1)
  {
    if (off == cb) continue;
#pragma unroll
    for (int i = 0; i < NVH; i++) acc[i] = cadd(acc[i], shfl_xor_c(acc[i], off));
  }
  if (live && (lane_in & ~cb) == 0)
  {
    const int h = (lane_in & cb) ? 1 : 0;
    cd* dst = coarse + coarse_index(g, xc, yc) * g.ncc + h * NVH;
#pragma unroll
    for (int i = 0; i < NVH; i++) dst[i] = overwrite ? acc[i] : cadd(dst[i], acc[i]);
  }
}

// fine element per thread: fine_out = (use_base ? base (or 0) : fine) + sum_i packed[idx][i] coarse[aggregate][h NVH + i]
template <int NVH>
__global__ void __launch_bounds__(256) prolong_packed_kernel(const TGeom g, const cd* __restrict__ packed, const cd* __restrict__ coarse,
                                                             cd* __restrict__ fine, const cd* __restrict__ base, const int use_base)
{
  const int rowlen = g.xhf * g.ncf;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= rowlen) return;
  const int y = blockIdx.y, p = blockIdx.z;
  const int k = col / g.ncf, c = col - k * g.ncf;
  const int x = 2 * k + ((y + p) & 1);
  const long ci = coarse_index(g, x / g.bx, y / g.by);
  const long idx = ((long)(y + p * g.Yf) * g.xhf) * g.ncf + col;
  const int h = (c >= g.ncf / 2) ? 1 : 0;
  cd acc = use_base ? cmake(0.0, 0.0) : fine[idx];
  const cd* cv = coarse + ci * g.ncc + h * NVH;
  cd n[NVH];
#pragma unroll
  for (int i = 0; i < NVH; i++) n[i] = ld_stream(packed + idx * NVH + i);
#pragma unroll
  for (int i = 0; i < NVH; i++) cfma(acc, n[i], __ldg(cv + i));
  if (use_base && base != nullptr) acc = cadd(base[idx], acc);
  fine[idx] = acc;
}

// shapes the packed kernels cover: symmetric chirality split (ncf, ncc even), NVH = ncc / 2 in {1, 2, 4, 8}, even blocks whose
// (row, parity) segments are a power of two <= 32 lanes and a multiple of ncf
static bool packed_shape_ok(const TGeom& g)
{
  const int nvh = g.ncc / 2;
  if ((g.ncf & 1) || (g.ncc & 1) || !(nvh == 1 || nvh == 2 || nvh == 4 || nvh == 8)) return false;
  if (!g.even_bx || g.seg > 32 || (g.seg & (g.seg - 1)) != 0 || g.seg % g.ncf != 0) return false;
  if ((g.ncf & (g.ncf - 1)) != 0) return false;
  return true;
}

template <int NV>
static int launch_restrict(const TGeom& g, const qmg_cplx* const* vecs, int count, int v0, const qmg_cplx* fine, qmg_cplx* coarse, int overwrite = 0)
{
  VecPack<NV> pk;
  for (int v = 0; v < NV; v++) pk.p[v] = reinterpret_cast<const cd*>(vecs[v < count ? v : 0]);
  // Lanes per aggregate.  With an even block width the aggregate's elements of one (row, parity) are `seg` contiguous
  // complex numbers and neighbouring aggregates continue the same memory row, so G = seg makes every warp-level load one
  // contiguous 512-byte span (32 / G aggregates side by side) and each lane walks down the 2 by (row, parity) segments of
  // its aggregate.  (G = 32 lanes on ONE aggregate reads eight 64-byte pieces of eight different DRAM pages per load:
  // 3.3 TB/s at 4096^2 -> 1024^2.)  Other shapes keep the element-strided assignment.
  int G = 1, logG = 0;
  if (g.even_bx && g.seg <= 32 && (g.seg & (g.seg - 1)) == 0) { while (G < g.seg) { G <<= 1; logG++; } }
  else { while (G < 32 && G < g.fspc) { G <<= 1; logG++; } }
  const long threads = g.Vc * G;
  const long blocks = (threads + 255) / 256;
  if (blocks > 0x7fffffffL) return fail_msg("restrict: lattice too large for the launch grid");
  restrict_kernel<NV><<<(unsigned)blocks, 256, 0, rt().stream>>>(g, pk, count, v0, reinterpret_cast<const cd*>(fine), reinterpret_cast<cd*>(coarse), G, logG, overwrite);
  QMG_LAUNCH_CHECK();
  return 0;
}

template <int NV>
static int launch_prolong(const TGeom& g, const qmg_cplx* const* vecs, int count, int v0, const qmg_cplx* coarse, qmg_cplx* fine,
                          const qmg_cplx* base = nullptr, int use_base = 0)
{
  VecPack<NV> pk;
  for (int v = 0; v < NV; v++) pk.p[v] = reinterpret_cast<const cd*>(vecs[v < count ? v : 0]);
  const int rowlen = g.xhf * g.ncf;
  if (g.Yf > 65535) return fail_msg("prolong: Y too large for the launch grid");
  dim3 grid((rowlen + 255) / 256, g.Yf, 2);
  prolong_kernel<NV><<<grid, 256, 0, rt().stream>>>(g, pk, count, v0, reinterpret_cast<const cd*>(coarse), reinterpret_cast<cd*>(fine),
                                                         reinterpret_cast<const cd*>(base), use_base);
  QMG_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------- block orthonormalise --
// One warp per aggregate; its (fspc x nvec) panel lives in shared memory while modified
// Gram-Schmidt runs (the reference expresses the same recurrences as nvec (nvec+1) / 2
// single-vector restrict / prolong sweeps over the whole lattice, transfer.h:540-602).
__global__ void __launch_bounds__(128) block_ortho_kernel(const TGeom g, cd* const* __restrict__ vecs, const int nvec, cd* __restrict__ chol)
{
  extern __shared__ cd panel_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  cd* panel = panel_all + (size_t)warp * g.fspc * nvec;
  for (long agg = (long)blockIdx.x * wpb + warp; agg < g.Vc; agg += (long)gridDim.x * wpb)
  {
    const int yc = (int)(agg / g.Xc), xc = (int)(agg - (long)yc * g.Xc);
    const long ci = coarse_index(g, xc, yc);
    for (int v = 0; v < nvec; v++)
      for (int e = lane; e < g.fspc; e += 32) panel[v * g.fspc + e] = vecs[v][agg_elem_index(g, xc, yc, e)];
    __syncwarp();
    for (int i = 0; i < nvec; i++)
    {
      cd* vi = panel + i * g.fspc;
      for (int j = 0; j < i; j++)
      {
        const cd* vj = panel + j * g.fspc;
        cd d = cmake(0.0, 0.0);
        for (int e = lane; e < g.fspc; e += 32) cfma_conj(d, vj[e], vi[e]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d = cadd(d, shfl_xor_c(d, o));
        if (chol != nullptr && lane == 0) chol[(ci * g.ncc + j) * g.ncc + i] = d;
        const cd md = cmake(-d.x, -d.y);
        for (int e = lane; e < g.fspc; e += 32) { cd t = vi[e]; cfma(t, md, vj[e]); vi[e] = t; }
        __syncwarp();
      }
      double nrm = 0.0;
      for (int e = lane; e < g.fspc; e += 32) nrm += vi[e].x * vi[e].x + vi[e].y * vi[e].y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nrm += shfl_xor_d(nrm, o);
      const double inv = 1.0 / sqrt(nrm);
      if (chol != nullptr && lane == 0) chol[(ci * g.ncc + i) * g.ncc + i] = cmake(1.0 / inv, 0.0);
      for (int e = lane; e < g.fspc; e += 32) vi[e] = cmake(vi[e].x * inv, vi[e].y * inv);
      __syncwarp();
    }
    for (int v = 0; v < nvec; v++)
      for (int e = lane; e < g.fspc; e += 32) vecs[v][agg_elem_index(g, xc, yc, e)] = panel[v * g.fspc + e];
    __syncwarp();
  }
}

// ------------------------------------------------------------- coarse build --
// Galerkin R A P one coarse site per CTA, computed directly instead of by 9 nc_c probe sweeps:
//   clover_c(I)[a][b]   = sum_{k in agg(I)} conj(R_a[k]) ( C(k) P_b[k] + sum_{d: k+d in agg(I)} H_d(k) P_b[k+d] )
//   hopping_c,d(I)[a][b] = sum_{k in agg(I), k+d outside} conj(R_a[k]) H_d(k) P_b[k+d]
// which is what the probing of coarse.h:140-444 evaluates (sources on one coarse parity so
// that neighbouring aggregates do not collide; 1-wide coarse dimensions fold into the clover).
struct CoarseBuildArgs
{
  TGeom g;
  const cd* clover;      // fine clover or nullptr
  const cd* hop;         // fine hopping or nullptr
  long size_cm_f;        // Vf ncf^2
  long size_cm_c;        // Vc ncc^2
  cd* const* P;          // device table of ncc prolong vectors
  cd* const* R;          // device table of ncc restrict vectors
  cd* clover_c;
  cd* hopping_c;
  // y-slab sharding: rows -1 / Y of the ncc prolong vectors from the ring neighbours (vector b at + b * prow), or nullptr
  const cd* Pym;
  const cd* Pyp;
  long prow;             // Xf ncf
};

__global__ void __launch_bounds__(256) coarse_build_kernel(const CoarseBuildArgs a, const int nchunk)
{
  extern __shared__ cd sm[];
  const TGeom& g = a.g;
  cd* T = sm;                                   // fspc x ncc
  cd* red = sm + (size_t)g.fspc * g.ncc;        // nchunk x ncc^2
  const int ncc2 = g.ncc * g.ncc;
  const long agg = blockIdx.x;
  const int yc = (int)(agg / g.Xc), xc = (int)(agg - (long)yc * g.Xc);
  const long ci = coarse_index(g, xc, yc);
  const int rowlen = g.bx * g.ncf;
  const long ncf2 = (long)g.ncf * g.ncf;

  for (int tgt = 0; tgt < 5; tgt++)
  {
    // phase A: T[k][b] = (A_tgt P_b)[k]
    for (int item = threadIdx.x; item < g.fspc * g.ncc; item += blockDim.x)
    {
      const int k = item / g.ncc, b = item - k * g.ncc;
      const int r = k / rowlen, rem = k - r * rowlen;
      const int xi = rem / g.ncf, c1 = rem - xi * g.ncf;
      const int x = xc * g.bx + xi, y = yc * g.by + r;
      const int p = (x + y) & 1;
      const long s = (long)(y + p * g.Yf) * g.xhf + (x >> 1);
      const cd* Pb = a.P[b];
      cd val = cmake(0.0, 0.0);
      if (tgt == 0 && a.clover != nullptr)
      {
        const cd* row = a.clover + s * ncf2 + (long)c1 * g.ncf;
        for (int c2 = 0; c2 < g.ncf; c2++) cfma(val, row[c2], Pb[s * g.ncf + c2]);
      }
      if (a.hop != nullptr)
        for (int mu = 0; mu < 4; mu++)
        {
          if (tgt != 0 && mu != tgt - 1) continue;
          int xn = x, yn = y;
          const cd* hrows = nullptr;     // set when the neighbour lives on the next slab
          if (mu == 0) xn = (x + 1 == g.Xf) ? 0 : x + 1;
          else if (mu == 2) xn = (x == 0) ? g.Xf - 1 : x - 1;
          else if (mu == 1) { if (y + 1 == g.Yf) { yn = 0; hrows = a.Pyp; } else yn = y + 1; }
          else { if (y == 0) { yn = g.Yf - 1; hrows = a.Pym; } else yn = y - 1; }
          const bool inside = (hrows == nullptr) && (xn / g.bx == xc) && (yn / g.by == yc);
          if ((tgt == 0) != inside) continue;
          const int pn = (xn + yn) & 1;   // Yf is even: row Y has the parity pattern of row 0, row -1 that of row Yf-1
          const cd* src = (hrows != nullptr) ? hrows + (long)b * a.prow + ((long)pn * g.xhf + (xn >> 1)) * g.ncf
                                            : Pb + ((long)(yn + pn * g.Yf) * g.xhf + (xn >> 1)) * g.ncf;
          const cd* row = a.hop + (long)mu * a.size_cm_f + s * ncf2 + (long)c1 * g.ncf;
          for (int c2 = 0; c2 < g.ncf; c2++) cfma(val, row[c2], src[c2]);
        }
      T[item] = val;
    }
    __syncthreads();
    // phase B: out[a][b] = sum_k conj(R_a[k]) T[k][b], k split into nchunk ranges
    const int per = (g.fspc + nchunk - 1) / nchunk;
    for (int item = threadIdx.x; item < ncc2 * nchunk; item += blockDim.x)
    {
      const int chunk = item / ncc2, pair = item - chunk * ncc2;
      const int ra = pair / g.ncc, b = pair - ra * g.ncc;
      const cd* Ra = a.R[ra];
      cd acc = cmake(0.0, 0.0);
      const int k1 = min(g.fspc, (chunk + 1) * per);
      for (int k = chunk * per; k < k1; k++)
      {
        const int r = k / rowlen, rem = k - r * rowlen;
        const int xi = rem / g.ncf, c1 = rem - xi * g.ncf;
        const int x = xc * g.bx + xi, y = yc * g.by + r;
        const int p = (x + y) & 1;
        const long idx = ((long)(y + p * g.Yf) * g.xhf + (x >> 1)) * g.ncf + c1;
        cfma_conj(acc, Ra[idx], T[k * g.ncc + b]);
      }
      red[item] = acc;
    }
    __syncthreads();
    for (int pair = threadIdx.x; pair < ncc2; pair += blockDim.x)
    {
      cd acc = red[pair];
      for (int c = 1; c < nchunk; c++) acc = cadd(acc, red[c * ncc2 + pair]);
      if (tgt == 0) a.clover_c[ci * ncc2 + pair] = acc;
      else a.hopping_c[(long)(tgt - 1) * a.size_cm_c + ci * ncc2 + pair] = acc;
    }
    __syncthreads();
  }
}

// device-resident copy of a host array of device pointers (kept alive by the caller until the kernel is enqueued;
// the copy itself is stream-ordered)
static int upload_ptrs(const qmg_cplx* const* host, int n, int slot, cd*** out)
{
  Runtime& r = rt();
  if (n * (slot + 1) > kMaxPtrs || n > kMaxPtrs / 2) return fail_msg("transfer: too many null vectors (max 128)");
  void** dst = r.d_ptrs + (size_t)slot * (kMaxPtrs / 2);
  QMG_CUDA(cudaMemcpyAsync(dst, host, sizeof(void*) * n, cudaMemcpyHostToDevice, r.stream));
  *out = reinterpret_cast<cd**>(dst);
  return 0;
}

} // namespace qmg

using namespace qmg;

extern "C" {

int qmg_restrict(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec, const qmg_cplx* fine, qmg_cplx* coarse)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_restrict"); if (rc) return rc;
  if (nvec < 1 || nvec > g.ncc) return fail_msg("qmg_restrict: nvec must be in [1, coarse nc]");
  int done = 0;
  while (done < nvec)
  {
    const int left = nvec - done;
    int take;
    if (left >= 8) { take = 8; rc = launch_restrict<8>(g, nullvecs_host + done, take, done, fine, coarse); }
    else if (left > 4) { take = left; rc = launch_restrict<8>(g, nullvecs_host + done, take, done, fine, coarse); }
    else if (left > 2) { take = left; rc = launch_restrict<4>(g, nullvecs_host + done, take, done, fine, coarse); }
    else if (left == 2) { take = 2; rc = launch_restrict<2>(g, nullvecs_host + done, take, done, fine, coarse); }
    else { take = 1; rc = launch_restrict<1>(g, nullvecs_host + done, take, done, fine, coarse); }
    if (rc) return rc;
    done += take;
  }
  return 0;
}

int qmg_prolong(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec, const qmg_cplx* coarse, qmg_cplx* fine)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_prolong"); if (rc) return rc;
  if (nvec < 1 || nvec > g.ncc) return fail_msg("qmg_prolong: nvec must be in [1, coarse nc]");
  int done = 0;
  while (done < nvec)
  {
    const int left = nvec - done;
    int take;
    if (left >= 8) { take = 8; rc = launch_prolong<8>(g, nullvecs_host + done, take, done, coarse, fine); }
    else if (left > 4) { take = left; rc = launch_prolong<8>(g, nullvecs_host + done, take, done, coarse, fine); }
    else if (left > 2) { take = left; rc = launch_prolong<4>(g, nullvecs_host + done, take, done, coarse, fine); }
    else if (left == 2) { take = 2; rc = launch_prolong<2>(g, nullvecs_host + done, take, done, coarse, fine); }
    else { take = 1; rc = launch_prolong<1>(g, nullvecs_host + done, take, done, coarse, fine); }
    if (rc) return rc;
    done += take;
  }
  return 0;
}

// coarse = P^dag fine, every coarse dof of the nvec vectors written outright (restrict_f2c after zero_vector, in one pass)
int qmg_restrict_overwrite(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec, const qmg_cplx* fine, qmg_cplx* coarse)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_restrict_overwrite"); if (rc) return rc;
  if (nvec != g.ncc) return fail_msg("qmg_restrict_overwrite: needs all coarse dof (nvec == coarse nc); the partial flavour is qmg_restrict");
  int done = 0;
  while (done < nvec)
  {
    const int left = nvec - done;
    int take;
    if (left >= 8) { take = 8; rc = launch_restrict<8>(g, nullvecs_host + done, take, done, fine, coarse, 1); }
    else if (left > 4) { take = left; rc = launch_restrict<8>(g, nullvecs_host + done, take, done, fine, coarse, 1); }
    else if (left > 2) { take = left; rc = launch_restrict<4>(g, nullvecs_host + done, take, done, fine, coarse, 1); }
    else if (left == 2) { take = 2; rc = launch_restrict<2>(g, nullvecs_host + done, take, done, fine, coarse, 1); }
    else { take = 1; rc = launch_restrict<1>(g, nullvecs_host + done, take, done, fine, coarse, 1); }
    if (rc) return rc;
    done += take;
  }
  return 0;
}

// fine_out = base + P coarse (base NULL: fine_out = P coarse): zero_vector + prolong_c2f + cxpyz of the K-cycle's
// correction step (stateful_multigrid.h:1005-1019) in one pass; fine_out may alias base
int qmg_prolong_add(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec, const qmg_cplx* coarse,
                    const qmg_cplx* base, qmg_cplx* fine_out)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_prolong_add"); if (rc) return rc;
  if (nvec < 1 || nvec > g.ncc) return fail_msg("qmg_prolong_add: nvec must be in [1, coarse nc]");
  // one pass holds 8 vectors; callers with more (none on the K-cycle path: coarse_dof = 8) use zero + qmg_prolong + cxpyz
  if (nvec > 8) return fail_msg("qmg_prolong_add: at most 8 vectors per fused pass");
  if (nvec > 4) return launch_prolong<8>(g, nullvecs_host, nvec, 0, coarse, fine_out, base, 1);
  if (nvec > 2) return launch_prolong<4>(g, nullvecs_host, nvec, 0, coarse, fine_out, base, 1);
  if (nvec == 2) return launch_prolong<2>(g, nullvecs_host, nvec, 0, coarse, fine_out, base, 1);
  return launch_prolong<1>(g, nullvecs_host, nvec, 0, coarse, fine_out, base, 1);
}

// 1 when qmg_transfer_pack_chiral / qmg_restrict_packed / qmg_prolong_packed cover this transfer's shape
int qmg_transfer_packed_supported(const qmg_transfer_desc* t)
{
  TGeom g;
  if (make_geom(t, g, "qmg_transfer_packed_supported")) return 0;
  return packed_shape_ok(g) ? 1 : 0;
}

int qmg_transfer_pack_chiral(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec, qmg_cplx* packed, double* dropped_norm2)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_transfer_pack_chiral"); if (rc) return rc;
  if (nvec != g.ncc || !packed_shape_ok(g)) return fail_msg("qmg_transfer_pack_chiral: shape not covered (see qmg_transfer_packed_supported)");
  const long nf = (long)g.Xf * g.Yf * g.ncf;
  Runtime& r = rt();
  long want = (nf + 255) / 256, cap = (long)r.sm_count * 8;
  const int grid = (int)(want < cap ? want : cap);
#define QMG_PACK(H) case H: { VecPack<2 * H> pk; for (int v = 0; v < 2 * H; v++) pk.p[v] = reinterpret_cast<const cd*>(nullvecs_host[v]); \
      pack_chiral_kernel<H><<<grid, 256, 0, r.stream>>>(nf, g.ncf, pk, reinterpret_cast<cd*>(packed), r.d_partials, r.d_counter, r.d_result); break; }
  switch (g.ncc / 2) { QMG_PACK(1) QMG_PACK(2) QMG_PACK(4) QMG_PACK(8) default: break; }
#undef QMG_PACK
  QMG_LAUNCH_CHECK();
  return fetch_result(dropped_norm2, 1);
}

// coarse (+)= P^dag fine from the packed copy; overwrite != 0: coarse = P^dag fine
int qmg_restrict_packed(const qmg_transfer_desc* t, const qmg_cplx* packed, const qmg_cplx* fine, qmg_cplx* coarse, int overwrite)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_restrict_packed"); if (rc) return rc;
  if (!packed_shape_ok(g)) return fail_msg("qmg_restrict_packed: shape not covered (see qmg_transfer_packed_supported)");
  int G = 1, logG = 0;
  while (G < g.seg) { G <<= 1; logG++; }
  const long threads = g.Vc * G;
  const long blocks = (threads + 255) / 256;
  if (blocks > 0x7fffffffL) return fail_msg("restrict: lattice too large for the launch grid");
#define QMG_RP(H) case H: restrict_packed_kernel<H><<<(unsigned)blocks, 256, 0, rt().stream>>>(g, reinterpret_cast<const cd*>(packed), \
      reinterpret_cast<const cd*>(fine), reinterpret_cast<cd*>(coarse), G, logG, overwrite); break;
  switch (g.ncc / 2) { QMG_RP(1) QMG_RP(2) QMG_RP(4) QMG_RP(8) default: break; }
#undef QMG_RP
  QMG_LAUNCH_CHECK();
  return 0;
}

// use_base == 0: fine_out += P coarse;  else fine_out = base + P coarse (base NULL: P coarse), the sum formed from zero first
int qmg_prolong_packed(const qmg_transfer_desc* t, const qmg_cplx* packed, const qmg_cplx* coarse, const qmg_cplx* base, qmg_cplx* fine_out, int use_base)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_prolong_packed"); if (rc) return rc;
  if (!packed_shape_ok(g)) return fail_msg("qmg_prolong_packed: shape not covered (see qmg_transfer_packed_supported)");
  const int rowlen = g.xhf * g.ncf;
  dim3 grid((rowlen + 255) / 256, g.Yf, 2);
  if (g.Yf > 65535) return fail_msg("prolong: Y too large for the launch grid");
#define QMG_PP(H) case H: prolong_packed_kernel<H><<<grid, 256, 0, rt().stream>>>(g, reinterpret_cast<const cd*>(packed), reinterpret_cast<const cd*>(coarse), \
      reinterpret_cast<cd*>(fine_out), reinterpret_cast<const cd*>(base), use_base); break;
  switch (g.ncc / 2) { QMG_PP(1) QMG_PP(2) QMG_PP(4) QMG_PP(8) default: break; }
#undef QMG_PP
  QMG_LAUNCH_CHECK();
  return 0;
}

int qmg_block_orthonormalize(const qmg_transfer_desc* t, qmg_cplx* const* nullvecs_host, int nvec, qmg_cplx* cholesky)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_block_orthonormalize"); if (rc) return rc;
  if (nvec < 1 || nvec > g.ncc) return fail_msg("qmg_block_orthonormalize: nvec must be in [1, coarse nc]");
  cd** dptrs;
  rc = upload_ptrs(nullvecs_host, nvec, 0, &dptrs); if (rc) return rc;
  const size_t per_warp = sizeof(cd) * (size_t)g.fspc * nvec;
  int wpb = 4;
  while (wpb > 1 && per_warp * wpb > 96 * 1024) wpb >>= 1;
  const size_t smem = per_warp * wpb;
  if (smem > 200 * 1024) return fail_msg("qmg_block_orthonormalize: aggregate panel does not fit in shared memory");
  if (smem > 48 * 1024) QMG_CUDA(cudaFuncSetAttribute(block_ortho_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long blocks = (g.Vc + wpb - 1) / wpb, cap = (long)rt().sm_count * 16;
  block_ortho_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 32 * wpb, smem, rt().stream>>>(g, dptrs, nvec, reinterpret_cast<cd*>(cholesky));
  QMG_LAUNCH_CHECK();
  // the pointer table is reused by the next call
  QMG_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}

int qmg_coarse_build(const qmg_transfer_desc* t, const qmg_stencil_desc* fine, const qmg_cplx* const* prolong_vecs_host,
                     const qmg_cplx* const* restrict_vecs_host, qmg_cplx* clover_c, qmg_cplx* hopping_c)
{
  QMG_REQUIRE_INIT();
  TGeom g; int rc = make_geom(t, g, "qmg_coarse_build"); if (rc) return rc;
  if (fine == nullptr || fine->X != g.Xf || fine->Y != g.Yf || fine->nc != g.ncf) return fail_msg("qmg_coarse_build: fine stencil does not match the transfer's fine lattice");
  CoarseBuildArgs a;
  a.g = g;
  a.clover = reinterpret_cast<const cd*>(fine->clover);
  a.hop = reinterpret_cast<const cd*>(fine->hopping);
  a.size_cm_f = (long)g.Xf * g.Yf * g.ncf * g.ncf;
  a.size_cm_c = g.Vc * g.ncc * g.ncc;
  cd** dP; cd** dR;
  rc = upload_ptrs(prolong_vecs_host, g.ncc, 0, &dP); if (rc) return rc;
  rc = upload_ptrs(restrict_vecs_host != nullptr ? restrict_vecs_host : prolong_vecs_host, g.ncc, 1, &dR); if (rc) return rc;
  a.P = dP; a.R = dR;
  a.clover_c = reinterpret_cast<cd*>(clover_c);
  a.hopping_c = reinterpret_cast<cd*>(hopping_c);
  a.Pym = nullptr; a.Pyp = nullptr; a.prow = (long)g.Xf * g.ncf;
  HaloTemp halo;
  if (comm().nranks > 1 && g.Yc < 2) return fail_msg("qmg_coarse_build: a slab must keep at least two coarse rows");
  if (comm().active && fine->hopping != nullptr && g.Yc > 1)   // (one coarse row: +-y hops fold into the clover, coarse.h:226-229)
  {
    // the coarse +-y links of the aggregates on the slab edges see the neighbouring slab's prolong vectors
    rc = halo.alloc_rows(g.Xf, g.ncf, g.ncc); if (rc) return rc;
    for (int b = 0; b < g.ncc; b++)
    {
      rc = halo_exchange_sync(reinterpret_cast<const cd*>(prolong_vecs_host[b]), g.Xf, g.Yf, g.ncf, halo.ym + b * a.prow, halo.yp + b * a.prow, 3);
      if (rc) return rc;
    }
    a.Pym = halo.ym; a.Pyp = halo.yp;
  }
  const int ncc2 = g.ncc * g.ncc;
  int nchunk = 256 / ncc2; if (nchunk < 1) nchunk = 1; if (nchunk > g.fspc) nchunk = g.fspc;
  const size_t smem = sizeof(cd) * ((size_t)g.fspc * g.ncc + (size_t)nchunk * ncc2);
  if (smem > 200 * 1024) return fail_msg("qmg_coarse_build: aggregate does not fit in shared memory");
  if (smem > 48 * 1024) QMG_CUDA(cudaFuncSetAttribute(coarse_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (g.Vc > 0x7fffffffL) return fail_msg("qmg_coarse_build: coarse lattice too large for the launch grid");
  coarse_build_kernel<<<(unsigned)g.Vc, 256, smem, rt().stream>>>(a, nchunk);
  QMG_LAUNCH_CHECK();
  QMG_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}

} // extern "C"
