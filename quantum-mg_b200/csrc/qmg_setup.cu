// Setup-path kernels: operator fills from U(1) links (K10), stencil-variant
// builders (K9), cshift, and the batched nc x nc site-matrix routines (K3).
// None of these is on the per-iteration path; they are written to be
// coalesced (one thread per OUTPUT element) and correct, not tuned.
#include "qmg_comm.cuh"
#include "qmg_launch.cuh"

namespace qmg {

constexpr int kBlock = 256;

template <class F>
__global__ void __launch_bounds__(kBlock) grid_stride_kernel(long n, F f)
{
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

template <class F> static int launch_n(long n, F f)
{
  if (n <= 0) return 0;
  long want = (n + kBlock - 1) / kBlock, cap = (long)rt().sm_count * 8;
  grid_stride_kernel<<<(int)(want < cap ? want : cap), kBlock, 0, rt().stream>>>(n, f);
  QMG_LAUNCH_CHECK();
  return 0;
}

// decode a full site index into (p, y, k)
__device__ __forceinline__ void site_decode(const Geom& g, long site, int& p, int& y, int& k)
{
  p = site >= (long)g.half ? 1 : 0;
  const unsigned h = (unsigned)(site - (long)p * g.half);
  y = h / g.xh; k = h % g.xh;
}

// full site index of the neighbour of `site` in direction mu
__device__ __forceinline__ long site_nbr(const Geom& g, long site, int mu)
{
  int p, y, k; site_decode(g, site, p, y, k);
  return (long)(1 - p) * g.half + nbr_h(g, p, y, k, mu);
}

// first dof of `field` (dof complex per site) at the neighbour of `site` in direction mu; across the slab edges the
// value comes from the rows received from the ring neighbours (y-slab sharding) when they are present
__device__ __forceinline__ const cd* field_nbr(const Geom& g, const cd* field, const HaloRows& h, long site, int mu, int dof)
{
  int p, y, k; site_decode(g, site, p, y, k);
  return nbr_site_ptr(field, h, g, p, y, k, mu, dof);
}

static inline int check_dims(int X, int Y, const char* who)
{
  if (X < 2 || Y < 2 || (X & 1) || (Y & 1)) { fail_msg(who); return 1; }
  return 0;
}

// Link that multiplies the neighbour in direction mu, as seen from `site`:
// U_mu(x) for forward hops, conj(U_mu(x - mu)) for backward hops
// (wilson.h:178-209: cshift FROM_XM1/YM1 then conj_vector).
// hy: rows -1 / Y of the U_y field when the lattice is a slab of a sharded one.
__device__ __forceinline__ cd link_for(const Geom& g, const cd* gauge, const HaloRows& hy, long V, long site, int mu)
{
  if (mu < 2) return gauge[(long)mu * V + site];
  if (mu == 3) return cconj(*field_nbr(g, gauge + V, hy, site, 3, 1));
  const long back = site_nbr(g, site, mu);
  return cconj(gauge[(long)(mu - 2) * V + back]);
}

// rows -1 / Y of U_y for the link fills
static int gauge_halo(HaloTemp& t, const cd* gauge, int X, int Y, HaloRows& hy)
{
  int rc = t.fetch(gauge + (long)X * Y, 0, 1, X, Y, 1);
  if (rc) return rc;
  hy = t.rows(0, X, 1);
  return 0;
}

} // namespace qmg

using namespace qmg;
#define CD(p) reinterpret_cast<cd*>(p)
#define CCD(p) reinterpret_cast<const cd*>(p)

extern "C" {

// operators/wilson.h:153-209.  Spin structure per direction:
//  +x: 1/2 [[-w, 1],[ 1,-w]]   +y: 1/2 [[-w,-i],[ i,-w]]
//  -x: 1/2 [[-w,-1],[-1,-w]]   -y: 1/2 [[-w, i],[-i,-w]]   clover = 2w * 1
int qmg_fill_wilson(int X, int Y, double w, const qmg_cplx* gauge_, qmg_cplx* clover_, qmg_cplx* hopping_)
{
  QMG_REQUIRE_INIT();
  if (check_dims(X, Y, "qmg_fill_wilson: X and Y must be even and >= 2")) return 2;
  const cd* gauge = CCD(gauge_); cd* clover = CD(clover_); cd* hop = CD(hopping_);
  Geom g; g.xh = X / 2; g.Y = Y; g.half = (unsigned)(X / 2) * Y;
  const long V = (long)X * Y;
  HaloTemp halo; HaloRows hy;
  { int hrc = gauge_halo(halo, gauge, X, Y, hy); if (hrc) return hrc; }
  return launch_n(V * 4 * 5, [=] __device__(long e) {
    const long per = V * 4;
    const int which = (int)(e / per);        // 0 clover, 1..4 hopping mu = which-1
    const long r = e - (long)which * per;
    const long site = r >> 2; const int c = (int)(r & 3);
    const bool diagel = (c == 0 || c == 3);
    if (which == 0) { clover[r] = diagel ? cmake(2.0 * w, 0.0) : cmake(0.0, 0.0); return; }
    const int mu = which - 1;
    const cd u = link_for(g, gauge, hy, V, site, mu);
    cd coef;
    if (diagel) coef = cmake(-0.5 * w, 0.0);
    else if (mu == 0) coef = cmake(0.5, 0.0);
    else if (mu == 2) coef = cmake(-0.5, 0.0);
    else if (mu == 1) coef = (c == 1) ? cmake(0.0, -0.5) : cmake(0.0, 0.5);
    else coef = (c == 1) ? cmake(0.0, 0.5) : cmake(0.0, -0.5);
    hop[(long)mu * per + r] = cmul(coef, u);
  });
}

// How far the stored clover / hopping blocks of an nc = 2 set are from the Wilson blocks of the gauge field the descriptor
// carries (wilson_gauge, wilson_w): result2 = { sum |stored - regenerated|^2, sum |stored|^2 } over all five blocks.  The
// matrix-free apply (csrc/qmg_stencil.cu wilson_mf_kernel) may be switched on only when the first number is EXACTLY zero.
int qmg_wilson_mf_deviation(const qmg_stencil_desc* st, double* result2)
{
  QMG_REQUIRE_INIT();
  if (st == nullptr || st->clover == nullptr || st->hopping == nullptr || st->wilson_gauge == nullptr)
    return fail_msg("qmg_wilson_mf_deviation: needs clover, hopping and wilson_gauge");
  if (st->nc != 2) return fail_msg("qmg_wilson_mf_deviation: nc must be 2");
  if (check_dims(st->X, st->Y, "qmg_wilson_mf_deviation: X and Y must be even and >= 2")) return 2;
  const cd* gauge = CCD(st->wilson_gauge); const cd* clover = CCD(st->clover); const cd* hop = CCD(st->hopping);
  const double w = st->wilson_w;
  Geom g; g.xh = st->X / 2; g.Y = st->Y; g.half = (unsigned)(st->X / 2) * st->Y;
  const long V = (long)st->X * st->Y;
  HaloTemp halo; HaloRows hy;
  { int hrc = gauge_halo(halo, gauge, st->X, st->Y, hy); if (hrc) return hrc; }
  return launch_reduce<2>(V * 4 * 5, [=] __device__(long e, double (&acc)[2]) {
    const long per = V * 4;
    const int which = (int)(e / per);
    const long r = e - (long)which * per;
    const long site = r >> 2; const int c = (int)(r & 3);
    const bool diagel = (c == 0 || c == 3);
    cd want, have;
    if (which == 0) { want = diagel ? cmake(2.0 * w, 0.0) : cmake(0.0, 0.0); have = clover[r]; }
    else
    {
      const int mu = which - 1;
      const cd u = link_for(g, gauge, hy, V, site, mu);
      cd coef;
      if (diagel) coef = cmake(-0.5 * w, 0.0);
      else if (mu == 0) coef = cmake(0.5, 0.0);
      else if (mu == 2) coef = cmake(-0.5, 0.0);
      else if (mu == 1) coef = (c == 1) ? cmake(0.0, -0.5) : cmake(0.0, 0.5);
      else coef = (c == 1) ? cmake(0.0, 0.5) : cmake(0.0, -0.5);
      want = cmul(coef, u);
      have = hop[(long)mu * per + r];
    }
    const double dx = have.x - want.x, dy = have.y - want.y;
    acc[0] += dx * dx + dy * dy;
    acc[1] += have.x * have.x + have.y * have.y;
  }, result2);
}

// operators/staggered.h:50-72: +x -1/2 U, +y -1/2 eta U, -x +1/2 U*, -y +1/2 eta U*, eta = 1 - 2 (x mod 2)
int qmg_fill_staggered(int X, int Y, const qmg_cplx* gauge_, qmg_cplx* hopping_)
{
  QMG_REQUIRE_INIT();
  if (check_dims(X, Y, "qmg_fill_staggered: X and Y must be even and >= 2")) return 2;
  const cd* gauge = CCD(gauge_); cd* hop = CD(hopping_);
  Geom g; g.xh = X / 2; g.Y = Y; g.half = (unsigned)(X / 2) * Y;
  const long V = (long)X * Y;
  HaloTemp halo; HaloRows hy;
  { int hrc = gauge_halo(halo, gauge, X, Y, hy); if (hrc) return hrc; }
  return launch_n(V * 4, [=] __device__(long e) {
    const int mu = (int)(e / V);
    const long site = e - (long)mu * V;
    int p, y, k; site_decode(g, site, p, y, k);
    const double eta = ((y + p) & 1) ? -1.0 : 1.0;      // x mod 2 = (y+p)&1
    double s = (mu < 2) ? -0.5 : 0.5;
    if (mu & 1) s *= eta;
    const cd u = link_for(g, gauge, hy, V, site, mu);
    hop[e] = cmake(s * u.x, s * u.y);
  });
}

// operators/gaugedlaplace.h:45-68: clover = 4, hopping = -U (conjugated, shifted for backward hops)
int qmg_fill_laplace(int X, int Y, const qmg_cplx* gauge_, qmg_cplx* clover_, qmg_cplx* hopping_)
{
  QMG_REQUIRE_INIT();
  if (check_dims(X, Y, "qmg_fill_laplace: X and Y must be even and >= 2")) return 2;
  const cd* gauge = CCD(gauge_); cd* clover = CD(clover_); cd* hop = CD(hopping_);
  Geom g; g.xh = X / 2; g.Y = Y; g.half = (unsigned)(X / 2) * Y;
  const long V = (long)X * Y;
  HaloTemp halo; HaloRows hy;
  { int hrc = gauge_halo(halo, gauge, X, Y, hy); if (hrc) return hrc; }
  return launch_n(V * 5, [=] __device__(long e) {
    const int which = (int)(e / V);
    const long site = e - (long)which * V;
    if (which == 0) { clover[site] = cmake(4.0, 0.0); return; }
    const int mu = which - 1;
    const cd u = link_for(g, gauge, hy, V, site, mu);
    hop[(long)mu * V + site] = cmake(-u.x, -u.y);
  });
}

// operators/dwf.h:154-237 (Shamir, nc = 2 Ls): Ls Wilson copies (clover 3w) on the
// 2x2 block diagonal, -P_+ / -P_- between adjacent s slices, +m P_-/P_+ wrap.
int qmg_fill_dwf(int X, int Y, int Ls, double w, double mass_re, double mass_im, const qmg_cplx* gauge_, qmg_cplx* clover_, qmg_cplx* hopping_)
{
  QMG_REQUIRE_INIT();
  if (check_dims(X, Y, "qmg_fill_dwf: X and Y must be even and >= 2")) return 2;
  if (Ls < 2) return fail_msg("qmg_fill_dwf: Ls must be >= 2");
  const cd* gauge = CCD(gauge_); cd* clover = CD(clover_); cd* hop = CD(hopping_);
  Geom g; g.xh = X / 2; g.Y = Y; g.half = (unsigned)(X / 2) * Y;
  const long V = (long)X * Y;
  const int nc = 2 * Ls; const long nc2 = (long)nc * nc;
  const cd mass = cmake(mass_re, mass_im);
  HaloTemp halo; HaloRows hy;
  { int hrc = gauge_halo(halo, gauge, X, Y, hy); if (hrc) return hrc; }
  return launch_n(V * nc2 * 5, [=] __device__(long e) {
    const long per = V * nc2;
    const int which = (int)(e / per);
    const long r = e - (long)which * per;
    const long site = r / nc2; const int c = (int)(r - site * nc2);
    const int row = c / nc, col = c % nc;
    if (which == 0)
    {
      cd v = cmake(0.0, 0.0);
      if (row == col) v = cmake(3.0 * w, 0.0);
      else if ((col & 1) == 0 && row == col + 2) v = cmake(-1.0, 0.0);          // -P_+ : (2j+2, 2j)
      else if ((row & 1) == 1 && col == row + 2) v = cmake(-1.0, 0.0);          // -P_- : (2j+1, 2j+3)
      if (row == nc - 1 && col == 1) v = mass;                                 // m P_-
      if (row == 0 && col == nc - 2) v = mass;                                 // m P_+
      clover[r] = v;
      return;
    }
    const int mu = which - 1;
    cd out = cmake(0.0, 0.0);
    if ((row >> 1) == (col >> 1))
    {
      const int cc = ((row & 1) << 1) | (col & 1);     // position inside the 2x2 Wilson block
      const bool diagel = (cc == 0 || cc == 3);
      cd coef;
      if (diagel) coef = cmake(-0.5 * w, 0.0);
      else if (mu == 0) coef = cmake(0.5, 0.0);
      else if (mu == 2) coef = cmake(-0.5, 0.0);
      else if (mu == 1) coef = (cc == 1) ? cmake(0.0, -0.5) : cmake(0.0, 0.5);
      else coef = (cc == 1) ? cmake(0.0, 0.5) : cmake(0.0, -0.5);
      out = cmul(coef, link_for(g, gauge, hy, V, site, mu));
    }
    hop[(long)mu * per + r] = out;
  });
}

// stencil/stencil_2d.h:1080-1139: dagger_clover = clover^dag,
// dagger_hopping_mu(x) = [hopping_{-mu}(x+mu)]^dag.
int qmg_build_dagger(int X, int Y, int nc, const qmg_cplx* clover_, const qmg_cplx* hopping_, qmg_cplx* dclover_, qmg_cplx* dhopping_)
{
  QMG_REQUIRE_INIT();
  if (check_dims(X, Y, "qmg_build_dagger: X and Y must be even and >= 2")) return 2;
  const cd* clover = CCD(clover_); const cd* hop = CCD(hopping_); cd* dclover = CD(dclover_); cd* dhop = CD(dhopping_);
  Geom g; g.xh = X / 2; g.Y = Y; g.half = (unsigned)(X / 2) * Y;
  const long V = (long)X * Y; const long nc2 = (long)nc * nc; const long per = V * nc2;
  int rc = 0;
  if (clover != nullptr && dclover != nullptr)
    rc = launch_n(per, [=] __device__(long e) {
      const long site = e / nc2; const int c = (int)(e - site * nc2);
      const int row = c / nc, col = c % nc;
      dclover[e] = cconj(clover[site * nc2 + (long)col * nc + row]);
    });
  if (rc) return rc;
  if (hop != nullptr && dhop != nullptr)
  {
    // sharded: row Y of hopping_{-y} (for mu = +y) and row -1 of hopping_{+y} (for mu = -y) come from the ring neighbours
    HaloTemp halo;
    rc = halo.fetch(hop, per, 4, X, Y, (int)nc2); if (rc) return rc;
    const cd* hym = halo.ym; const cd* hyp = halo.yp;
    const long hrow = (long)X * nc2;
    rc = launch_n(per * 4, [=] __device__(long e) {
      const int mu = (int)(e / per);
      const long r = e - (long)mu * per;
      const long site = r / nc2; const int c = (int)(r - site * nc2);
      const int row = c / nc, col = c % nc;
      const int om = opposite_dir(mu);
      HaloRows h; if (hym != nullptr) { h.ym = hym + om * hrow; h.yp = hyp + om * hrow; }
      const cd* nb = field_nbr(g, hop + (long)om * per, h, site, mu, (int)nc2);
      dhop[e] = cconj(nb[(long)col * nc + row]);
    });
  }
  return rc;
}

// distance of the stored backward blocks from  H_{-mu}(x)[a][b] = s_a s_b conj(H_{+mu}(x - mu)[b][a])
int qmg_stencil_gamma5_deviation(const qmg_stencil_desc* st, double* result2)
{
  QMG_REQUIRE_INIT();
  if (st == nullptr || st->hopping == nullptr) return fail_msg("qmg_stencil_gamma5_deviation: needs a hopping term");
  if (check_dims(st->X, st->Y, "qmg_stencil_gamma5_deviation: X and Y must be even and >= 2")) return 2;
  const int nc = st->nc;
  if (nc % 2) return fail_msg("qmg_stencil_gamma5_deviation: nc must be even");
  const cd* hop = CCD(st->hopping);
  Geom g; g.xh = st->X / 2; g.Y = st->Y; g.half = (unsigned)(st->X / 2) * st->Y;
  const long V = (long)st->X * st->Y; const long nc2 = (long)nc * nc; const long per = V * nc2;
  HaloTemp halo;
  int rc = halo.fetch(hop + per, 0, 1, st->X, st->Y, (int)nc2); if (rc) return rc;     // rows -1 / Y of the +y blocks
  const HaloRows hr = halo.rows(0, st->X, (int)nc2);
  return launch_reduce<2>(per * 2, [=] __device__(long e, double (&acc)[2]) {
    const int mu = 2 + (int)(e / per);                 // backward direction
    const long r = e % per;
    const long site = r / nc2; const int c = (int)(r - site * nc2);
    const int row = c / nc, col = c % nc;
    const cd back = hop[(long)mu * per + r];
    const cd* fwd = (mu == 3) ? field_nbr(g, hop + per, hr, site, 3, (int)nc2) : hop + site_nbr(g, site, 2) * nc2;
    const cd f = fwd[(long)col * nc + row];
    const double sg = ((2 * row < nc) == (2 * col < nc)) ? 1.0 : -1.0;
    const double dx = back.x - sg * f.x, dy = back.y + sg * f.y;
    acc[0] += dx * dx + dy * dy;
    acc[1] += back.x * back.x + back.y * back.y;
  }, result2);
}

// cshift/cshift_2d.h:225: lhs(x) = rhs(x + dir), written on the parity OPPOSITE to each source parity in eo.
int qmg_cshift(qmg_cplx* lhs_, const qmg_cplx* rhs_, int cdir, int eo, int dof, int X, int Y)
{
  QMG_REQUIRE_INIT();
  if (check_dims(X, Y, "qmg_cshift: X and Y must be even and >= 2")) return 2;
  if (cdir < 2 || cdir > 5) return fail_msg("qmg_cshift: only distance-one shifts exist (cshift_2d.h:120-129)");
  cd* lhs = CD(lhs_); const cd* rhs = CCD(rhs_);
  Geom g; g.xh = X / 2; g.Y = Y; g.half = (unsigned)(X / 2) * Y;
  const int mu = cdir - 2;    // QMG_CSHIFT_FROM_XP1=2 .. YM1=5  ->  +x,+y,-x,-y
  const long V = (long)X * Y;
  HaloTemp halo; HaloRows h;
  if (comm().active && (mu & 1))
  {
    // the periodic-boundary loops the reference marks "Becomes MPI" (cshift_2d.h:101,114)
    if (halo.alloc_rows(X, dof)) return 1;
    int rc = halo_exchange_sync(rhs, X, Y, dof, halo.ym, halo.yp, eo & 3); if (rc) return rc;
    h = halo.rows(0, X, dof);
  }
  return launch_n(V * dof, [=] __device__(long e) {
    const long site = e / dof; const int d = (int)(e - site * dof);
    const int p = site >= (long)g.half ? 1 : 0;
    // destination parity p receives from source parity 1-p: FROM_EVEN(1) writes odd, FROM_ODD(2) writes even
    const int need = p ? 1 : 2;
    if (!(eo & need)) return;
    lhs[e] = field_nbr(g, rhs, h, site, mu, dof)[d];
  });
}

// ------------------------------------------------------ batched site matrices --

int qmg_cmat_xy(const qmg_cplx* M_, const qmg_cplx* x_, qmg_cplx* y_, long nsites, int nc, int accumulate)
{
  QMG_REQUIRE_INIT();
  const cd* M = CCD(M_); const cd* x = CCD(x_); cd* y = CD(y_);
  return launch_n(nsites * nc, [=] __device__(long e) {
    const long s = e / nc; const int r = (int)(e - s * nc);
    cd acc = accumulate ? y[e] : cmake(0.0, 0.0);
    const cd* row = M + (s * nc + r) * nc;
    for (int c = 0; c < nc; c++) cfma(acc, row[c], x[s * nc + c]);
    y[e] = acc;
  });
}

int qmg_cmat_single_xy(const qmg_cplx* M_, const qmg_cplx* x_, qmg_cplx* y_, long nsites, int nc)
{
  QMG_REQUIRE_INIT();
  const cd* M = CCD(M_); const cd* x = CCD(x_); cd* y = CD(y_);
  return launch_n(nsites * nc, [=] __device__(long e) {
    const long s = e / nc; const int r = (int)(e - s * nc);
    cd acc = cmake(0.0, 0.0);
    for (int c = 0; c < nc; c++) cfma(acc, M[(long)r * nc + c], x[s * nc + c]);
    y[e] = acc;
  });
}

int qmg_cmat_conjtrans(const qmg_cplx* in_, qmg_cplx* out_, long nsites, int nc)
{
  QMG_REQUIRE_INIT();
  const cd* in = CCD(in_); cd* out = CD(out_);
  const long nc2 = (long)nc * nc;
  if ((const void*)in == (const void*)out)
    return launch_n(nsites * nc2, [=] __device__(long e) {
      const long s = e / nc2; const int c = (int)(e - s * nc2);
      const int row = c / nc, col = c % nc;
      if (row > col) return;
      const long a = s * nc2 + (long)row * nc + col, b = s * nc2 + (long)col * nc + row;
      const cd va = out[a], vb = out[b];
      out[a] = cconj(vb); out[b] = cconj(va);
    });
  return launch_n(nsites * nc2, [=] __device__(long e) {
    const long s = e / nc2; const int c = (int)(e - s * nc2);
    const int row = c / nc, col = c % nc;
    out[e] = cconj(in[s * nc2 + (long)col * nc + row]);
  });
}

int qmg_cmat_mul(const qmg_cplx* X_, const qmg_cplx* Y_, qmg_cplx* Z_, long nsites, int nc)
{
  QMG_REQUIRE_INIT();
  const cd* Xm = CCD(X_); const cd* Ym = CCD(Y_); cd* Zm = CD(Z_);
  const long nc2 = (long)nc * nc;
  return launch_n(nsites * nc2, [=] __device__(long e) {
    const long s = e / nc2; const int c = (int)(e - s * nc2);
    const int row = c / nc, col = c % nc;
    cd acc = cmake(0.0, 0.0);
    for (int k = 0; k < nc; k++) cfma(acc, Xm[s * nc2 + (long)row * nc + k], Ym[s * nc2 + (long)k * nc + col]);
    Zm[e] = acc;
  });
}

int qmg_cmat_add_pattern(const double* pattern_host, int len, qmg_cplx* v_, long nrepeat)
{
  QMG_REQUIRE_INIT();
  if (len > kMaxPtrs) return fail_msg("qmg_cmat_add_pattern: pattern longer than 256 elements");
  Runtime& r = rt();
  QMG_CUDA(cudaMemcpyAsync(r.d_scalars, pattern_host, sizeof(double) * 2 * len, cudaMemcpyHostToDevice, r.stream));
  const cd* pat = reinterpret_cast<const cd*>(r.d_scalars);
  cd* v = CD(v_);
  int rc = launch_n(nrepeat * len, [=] __device__(long e) {
    const cd pv = pat[e % len];
    cd t = v[e]; t.x += pv.x; t.y += pv.y; v[e] = t;
  });
  // the staging table may be reused by the next call: order it after this kernel
  if (!rc) QMG_CUDA(cudaStreamSynchronize(r.stream));
  return rc;
}

} // extern "C"

namespace qmg {

// One warp inverts one nc x nc matrix by Gauss-Jordan elimination with partial
// pivoting on [A | 1] held in shared memory.  (The reference gets the same
// inverse from a batched QR, stencil_2d.h:1536-1537; both are backward stable.)
__global__ void __launch_bounds__(32) cmat_inverse_kernel(const cd* M, cd* Minv, long nsites, int nc)
{
  extern __shared__ cd aug[];            // nc rows x 2nc columns
  const int lane = threadIdx.x;
  const int w = 2 * nc;
  for (long s = blockIdx.x; s < nsites; s += gridDim.x)
  {
    const cd* m = M + s * (long)nc * nc;
    for (int e = lane; e < nc * w; e += 32)
    {
      const int r = e / w, c = e % w;
      aug[e] = (c < nc) ? m[r * nc + c] : cmake(c - nc == r ? 1.0 : 0.0, 0.0);
    }
    __syncwarp();
    for (int j = 0; j < nc; j++)
    {
      // pivot search over rows j..nc-1 of column j
      double best = -1.0; int brow = j;
      for (int r = j + lane; r < nc; r += 32)
      {
        const cd v = aug[r * w + j];
        const double mag = v.x * v.x + v.y * v.y;
        if (mag > best) { best = mag; brow = r; }
      }
      for (int o = 16; o > 0; o >>= 1)
      {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
        if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
      }
      if (brow != j)
        for (int c = lane; c < w; c += 32) { const cd t = aug[j * w + c]; aug[j * w + c] = aug[brow * w + c]; aug[brow * w + c] = t; }
      __syncwarp();
      const cd piv = aug[j * w + j];
      __syncwarp();
      for (int c = lane; c < w; c += 32) aug[j * w + c] = cdiv(aug[j * w + c], piv);
      __syncwarp();
      for (int r = 0; r < nc; r++)
      {
        if (r == j) continue;
        const cd f = aug[r * w + j];
        __syncwarp();
        for (int c = lane; c < w; c += 32)
        {
          cd t = aug[r * w + c];
          cfma(t, cmake(-f.x, -f.y), aug[j * w + c]);
          aug[r * w + c] = t;
        }
        __syncwarp();
      }
    }
    cd* out = Minv + s * (long)nc * nc;
    for (int e = lane; e < nc * nc; e += 32) out[e] = aug[(e / nc) * w + nc + (e % nc)];
    __syncwarp();
  }
}

// One warp factorises one nc x nc matrix, M = Q R, by modified Gram-Schmidt over the columns (Q unitary, R upper
// triangular with a real positive diagonal) -- the factorisation quantum-linalg's cMATx_do_qr_square hands to
// cMATqr_do_xinv_square (stencil/stencil_2d.h:1536-1537).  Q and R live in shared memory while the warp works.
__global__ void __launch_bounds__(32) cmat_qr_kernel(const cd* M, cd* Q, cd* R, long nsites, int nc)
{
  extern __shared__ cd qr_smem[];        // q: nc x nc, r: nc x nc
  cd* q = qr_smem; cd* r = qr_smem + nc * nc;
  const int lane = threadIdx.x;
  for (long s = blockIdx.x; s < nsites; s += gridDim.x)
  {
    const cd* m = M + s * (long)nc * nc;
    for (int e = lane; e < nc * nc; e += 32) { q[e] = m[e]; r[e] = cmake(0.0, 0.0); }
    __syncwarp();
    for (int j = 0; j < nc; j++)
    {
      for (int i = 0; i < j; i++)
      {
        cd d = cmake(0.0, 0.0);
        for (int k = lane; k < nc; k += 32) cfma_conj(d, q[k * nc + i], q[k * nc + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d = cadd(d, shfl_xor_c(d, o));
        if (lane == 0) r[i * nc + j] = d;
        const cd md = cmake(-d.x, -d.y);
        for (int k = lane; k < nc; k += 32) { cd t = q[k * nc + j]; cfma(t, md, q[k * nc + i]); q[k * nc + j] = t; }
        __syncwarp();
      }
      double nrm = 0.0;
      for (int k = lane; k < nc; k += 32) { const cd v = q[k * nc + j]; nrm += v.x * v.x + v.y * v.y; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nrm += shfl_xor_d(nrm, o);
      nrm = sqrt(nrm);
      if (lane == 0) r[j * nc + j] = cmake(nrm, 0.0);
      for (int k = lane; k < nc; k += 32) { const cd v = q[k * nc + j]; q[k * nc + j] = cmake(v.x / nrm, v.y / nrm); }
      __syncwarp();
    }
    cd* qo = Q + s * (long)nc * nc; cd* ro = R + s * (long)nc * nc;
    for (int e = lane; e < nc * nc; e += 32) { qo[e] = q[e]; ro[e] = r[e]; }
    __syncwarp();
  }
}

// Minv = R^-1 Q^dag: lane c owns column c of the result and back-substitutes R x = (Q^dag)[:, c]
__global__ void __launch_bounds__(32) cmat_qr_inverse_kernel(const cd* Q, const cd* R, cd* Minv, long nsites, int nc)
{
  extern __shared__ cd qr_smem[];
  cd* q = qr_smem; cd* r = qr_smem + nc * nc; cd* x = r + nc * nc;      // x: nc x nc result
  const int lane = threadIdx.x;
  for (long s = blockIdx.x; s < nsites; s += gridDim.x)
  {
    const cd* qi = Q + s * (long)nc * nc; const cd* ri = R + s * (long)nc * nc;
    for (int e = lane; e < nc * nc; e += 32) { q[e] = qi[e]; r[e] = ri[e]; }
    __syncwarp();
    for (int c = lane; c < nc; c += 32)
      for (int i = nc - 1; i >= 0; i--)
      {
        cd t = cconj(q[c * nc + i]);                      // (Q^dag)[i][c]
        for (int k = i + 1; k < nc; k++) cfma(t, cmake(-r[i * nc + k].x, -r[i * nc + k].y), x[k * nc + c]);
        x[i * nc + c] = cdiv(t, r[i * nc + i]);
      }
    __syncwarp();
    cd* out = Minv + s * (long)nc * nc;
    for (int e = lane; e < nc * nc; e += 32) out[e] = x[e];
    __syncwarp();
  }
}

} // namespace qmg

extern "C" {

int qmg_cmat_inverse(const qmg_cplx* M_, qmg_cplx* Minv_, long nsites, int nc)
{
  QMG_REQUIRE_INIT();
  if (nsites <= 0) return 0;
  if (nc > 64) return fail_msg("qmg_cmat_inverse: nc > 64 unsupported");
  const size_t smem = sizeof(cd) * (size_t)nc * 2 * nc;
  if (smem > 48 * 1024) QMG_CUDA(cudaFuncSetAttribute(cmat_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long cap = (long)rt().sm_count * 32;
  cmat_inverse_kernel<<<(int)(nsites < cap ? nsites : cap), 32, smem, rt().stream>>>(CCD(M_), CD(Minv_), nsites, nc);
  QMG_LAUNCH_CHECK();
  return 0;
}

// M = Q R per site (modified Gram-Schmidt) and Minv = R^-1 Q^dag: the pair quantum-linalg's cMATx_do_qr_square /
// cMATqr_do_xinv_square name (stencil/stencil_2d.h:1536-1537), for callers that read Q or R
int qmg_cmat_qr(const qmg_cplx* M_, qmg_cplx* Q_, qmg_cplx* R_, long nsites, int nc)
{
  QMG_REQUIRE_INIT();
  if (nsites <= 0) return 0;
  if (nc > 48) return fail_msg("qmg_cmat_qr: nc > 48 unsupported");
  const size_t smem = sizeof(cd) * (size_t)nc * nc * 2;
  if (smem > 48 * 1024) QMG_CUDA(cudaFuncSetAttribute(cmat_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long cap = (long)rt().sm_count * 32;
  cmat_qr_kernel<<<(int)(nsites < cap ? nsites : cap), 32, smem, rt().stream>>>(CCD(M_), CD(Q_), CD(R_), nsites, nc);
  QMG_LAUNCH_CHECK();
  return 0;
}
int qmg_cmat_qr_inverse(const qmg_cplx* Q_, const qmg_cplx* R_, qmg_cplx* Minv_, long nsites, int nc)
{
  QMG_REQUIRE_INIT();
  if (nsites <= 0) return 0;
  if (nc > 48) return fail_msg("qmg_cmat_qr_inverse: nc > 48 unsupported");
  const size_t smem = sizeof(cd) * (size_t)nc * nc * 3;
  if (smem > 48 * 1024) QMG_CUDA(cudaFuncSetAttribute(cmat_qr_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long cap = (long)rt().sm_count * 32;
  cmat_qr_inverse_kernel<<<(int)(nsites < cap ? nsites : cap), 32, smem, rt().stream>>>(CCD(Q_), CCD(R_), CD(Minv_), nsites, nc);
  QMG_LAUNCH_CHECK();
  return 0;
}

// stencil/stencil_2d.h:1452-1601: cinv = (clover + diag shifts)^-1, identity clover,
// rbj_hopping_mu(x) = hopping_mu(x) * cinv(x+mu).
int qmg_build_rbjacobi(const qmg_stencil_desc* st, qmg_cplx* cinv_, qmg_cplx* rbj_clover_, qmg_cplx* rbj_hopping_)
{
  QMG_REQUIRE_INIT();
  if (st == nullptr) return fail_msg("qmg_build_rbjacobi: null stencil");
  if (check_dims(st->X, st->Y, "qmg_build_rbjacobi: X and Y must be even and >= 2")) return 2;
  const int nc = st->nc;
  const long V = (long)st->X * st->Y; const long nc2 = (long)nc * nc; const long per = V * nc2;
  Geom g; g.xh = st->X / 2; g.Y = st->Y; g.half = (unsigned)(st->X / 2) * st->Y;
  const cd* clover = CCD(st->clover); const cd* hop = CCD(st->hopping);
  cd* cinv = CD(cinv_); cd* rclover = CD(rbj_clover_); cd* rhop = CD(rbj_hopping_);
  const bool dof_ok = (nc % 2 == 0);
  const cd sh = cmake(st->shift[0], st->shift[1]), eo = cmake(st->eo_shift[0], st->eo_shift[1]);
  const cd df = dof_ok ? cmake(st->dof_shift[0], st->dof_shift[1]) : cmake(0.0, 0.0);
  // B = clover + diag(shift +- eo_shift +- dof_shift), staged in rbj_clover (overwritten with 1 afterwards)
  int rc = launch_n(per, [=] __device__(long e) {
    const long site = e / nc2; const int c = (int)(e - site * nc2);
    const int row = c / nc, col = c % nc;
    cd v = clover != nullptr ? clover[e] : cmake(0.0, 0.0);
    if (row == col)
    {
      const double se = site >= (long)g.half ? -1.0 : 1.0;
      const double sd = (2 * row >= nc && nc > 1) ? -1.0 : 1.0;
      v.x += sh.x + se * eo.x + sd * df.x;
      v.y += sh.y + se * eo.y + sd * df.y;
    }
    rclover[e] = v;
  });
  if (rc) return rc;
  rc = qmg_cmat_inverse(rbj_clover_, cinv_, V, nc);
  if (rc) return rc;
  rc = launch_n(per, [=] __device__(long e) {
    const int c = (int)(e % nc2);
    rclover[e] = cmake((c / nc == c % nc) ? 1.0 : 0.0, 0.0);
  });
  if (rc) return rc;
  if (hop != nullptr && rhop != nullptr)
  {
    // sharded: B^-1 on rows -1 / Y comes from the ring neighbours
    HaloTemp halo;
    rc = halo.fetch(cinv, 0, 1, st->X, st->Y, (int)nc2); if (rc) return rc;
    const HaloRows hr = halo.rows(0, st->X, (int)nc2);
    rc = launch_n(per * 4, [=] __device__(long e) {
      const int mu = (int)(e / per);
      const long r = e - (long)mu * per;
      const long site = r / nc2; const int c = (int)(r - site * nc2);
      const int row = c / nc, col = c % nc;
      const cd* h = hop + (long)mu * per + site * nc2 + (long)row * nc;
      const cd* b = field_nbr(g, cinv, hr, site, mu, (int)nc2) + col;
      cd acc = cmake(0.0, 0.0);
      for (int k = 0; k < nc; k++) cfma(acc, h[k], b[(long)k * nc]);
      rhop[e] = acc;
    });
  }
  return rc;
}

} // extern "C"
