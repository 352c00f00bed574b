// U(1) gauge-field side of the path: plaquette / topological charge / non-compact action,
// gauge transformation, APE smearing and the non-compact heatbath, on the nc = 1 even-odd layout
// (gauge[mu * V + site], mu in {x, y}; /root/reference/u1/u1_utils.h).
//
// The reference builds each of these from full-lattice cshift + BLAS passes (u1_utils.h:241-508) and sweeps the
// heatbath serially, site after site (:607-667, "This algorithm can't be parallelized as is... We would need subsets").
// Here every observable is ONE pass that reads each link once and gathers its 2-3 neighbours through L1/L2, and the
// heatbath is the same conditional gaussian update applied to four independent subsets in turn (x links on even rows,
// x links on odd rows, y links on even columns, y links on odd columns): an x link only sees x links of the rows
// above and below, a y link only y links of the columns left and right, so each subset is updated in parallel from a
// counter-based generator.  It is a different sweep order of the same Markov kernel: same stationary distribution
// (<cos plaq> = exp(-1/(2 beta))), not the same random stream as the serial sweep.
//
// y-slab sharding: rows -1 / Y of the fields come from the ring neighbours (one exchange per pass / per subset).
#include "qmg_comm.cuh"
#include "qmg_launch.cuh"

namespace qmg {

// A scalar (dof = 1) site field seen by coordinates, with optional halo rows for y = -1 and y = Y.
template <typename T>
struct FieldView
{
  const T* base; const T* ym; const T* yp;
  int X, Y, xh;
  __device__ __forceinline__ T at(int x, int y) const
  {
    if (x == X) x = 0; else if (x < 0) x = X - 1;
    const int par = (x + y) & 1;            // Y even: rows -1 and Y carry the parity pattern of rows Y-1 and 0
    if (y < 0) { if (ym != nullptr) return ym[par * xh + (x >> 1)]; y = Y - 1; }
    else if (y >= Y) { if (yp != nullptr) return yp[par * xh + (x >> 1)]; y = 0; }
    return base[(size_t)(y + par * Y) * xh + (x >> 1)];
  }
};

template <typename T>
static FieldView<T> make_view(const T* base, const T* ym, const T* yp, int X, int Y)
{
  FieldView<T> v; v.base = base; v.ym = ym; v.yp = yp; v.X = X; v.Y = Y; v.xh = X / 2; return v;
}

// coordinates of site index s (all even sites, then all odd)
__device__ __forceinline__ void site_xy(int X, int Y, long s, int& x, int& y)
{
  const int xh = X / 2; const long half = (long)xh * Y;
  const int p = s >= half ? 1 : 0;
  const long h = s - p * half;
  y = (int)(h / xh);
  x = 2 * (int)(h - (long)y * xh) + ((y + p) & 1);
}

static int check_gauge_dims(int X, int Y, const char* who)
{
  if (X < 2 || Y < 2 || (X & 1) || (Y & 1)) return fail_msg(who);
  return 0;
}

// Halo rows of the two link fields (mu = x, y) of a complex gauge field or of a real phase field.
// A real field of X/2 doubles per parity row is exchanged as a complex field of X/4 elements per row.
struct GaugeHalo
{
  HaloTemp t;
  const void* ym[2] = { nullptr, nullptr };
  const void* yp[2] = { nullptr, nullptr };
  int fetch_complex(const cd* gauge, int X, int Y)
  {
    if (!comm().active) return 0;
    int rc = t.fetch(gauge, (long)X * Y, 2, X, Y, 1); if (rc) return rc;
    for (int mu = 0; mu < 2; mu++) { ym[mu] = t.ym + (size_t)mu * X; yp[mu] = t.yp + (size_t)mu * X; }
    return 0;
  }
  int fetch_real(const double* phases, int X, int Y)
  {
    if (!comm().active) return 0;
    if (X % 4 != 0) return fail_msg("U(1) phase fields on y-slabs need X divisible by 4");
    int rc = t.fetch(reinterpret_cast<const cd*>(phases), (long)X * Y / 2, 2, X / 2, Y, 1); if (rc) return rc;
    for (int mu = 0; mu < 2; mu++)
    {
      ym[mu] = reinterpret_cast<const double*>(t.ym) + (size_t)mu * X;
      yp[mu] = reinterpret_cast<const double*>(t.yp) + (size_t)mu * X;
    }
    return 0;
  }
};

} // namespace qmg

using namespace qmg;

extern "C" {

// phases -> compact links (quantum-linalg polar_vector, tests/n13_wilson_kcycle/wilson_kcycle.cpp:212)
int qmg_polar_vector(const double* phases, qmg_cplx* out_, long n)
{
  QMG_REQUIRE_INIT();
  cd* out = reinterpret_cast<cd*>(out_);
  return launch_ew(n, [=] __device__(long i) { double s, c; sincos(phases[i], &s, &c); out[i] = cmake(c, s); });
}

// result4 = { Re <plaq>, Im <plaq>, topological charge, 0 }: the average of U_x(x) U_y(x+x^) U_x*(x+y^) U_y*(x) over the
// (global) volume and sum_x arg(plaq) / 2 pi   (u1_utils.h:424-508), one pass for both.
int qmg_u1_plaquette(const qmg_cplx* gauge_, int X, int Y, double* result4)
{
  QMG_REQUIRE_INIT();
  int rc = check_gauge_dims(X, Y, "qmg_u1_plaquette: X and Y must be even and >= 2"); if (rc) return rc;
  const cd* gauge = reinterpret_cast<const cd*>(gauge_);
  const long V = (long)X * Y;
  GaugeHalo halo; rc = halo.fetch_complex(gauge, X, Y); if (rc) return rc;
  const FieldView<cd> ux = make_view(gauge, (const cd*)halo.ym[0], (const cd*)halo.yp[0], X, Y);
  const FieldView<cd> uy = make_view(gauge + V, (const cd*)halo.ym[1], (const cd*)halo.yp[1], X, Y);
  double out[3];
  rc = launch_reduce<3>(V, [=] __device__(long s, double (&acc)[3]) {
    int x, y; site_xy(X, Y, s, x, y);
    cd p = cmul(gauge[s], uy.at(x + 1, y));
    p = cmul(p, cconj(ux.at(x, y + 1)));
    p = cmul(p, cconj(gauge[V + s]));
    acc[0] += p.x; acc[1] += p.y; acc[2] += atan2(p.y, p.x);
  }, out);
  if (rc) return rc;
  const double Vg = (double)V * qmg_comm_size();
  result4[0] = out[0] / Vg; result4[1] = out[1] / Vg; result4[2] = out[2] * 0.5 / 3.14159265358979323846; result4[3] = 0.0;
  return 0;
}

// beta/2 sum_x (A_x(x) + A_y(x+x^) - A_x(x+y^) - A_y(x))^2   (u1_utils.h:386-421)
int qmg_u1_noncompact_action(const double* phases, int X, int Y, double beta, double* result)
{
  QMG_REQUIRE_INIT();
  int rc = check_gauge_dims(X, Y, "qmg_u1_noncompact_action: X and Y must be even and >= 2"); if (rc) return rc;
  const long V = (long)X * Y;
  GaugeHalo halo; rc = halo.fetch_real(phases, X, Y); if (rc) return rc;
  const FieldView<double> ax = make_view(phases, (const double*)halo.ym[0], (const double*)halo.yp[0], X, Y);
  const FieldView<double> ay = make_view(phases + V, (const double*)halo.ym[1], (const double*)halo.yp[1], X, Y);
  double out[1];
  rc = launch_reduce<1>(V, [=] __device__(long s, double (&acc)[1]) {
    int x, y; site_xy(X, Y, s, x, y);
    const double f = phases[s] + ay.at(x + 1, y) - ax.at(x, y + 1) - phases[V + s];
    acc[0] += f * f;
  }, out);
  if (rc) return rc;
  *result = 0.5 * beta * out[0];
  return 0;
}

// u_mu(x) <- g(x) u_mu(x) g*(x + mu)   (u1_utils.h:241-272); trans: V complex on the same lattice
int qmg_u1_gauge_transform(qmg_cplx* gauge_, const qmg_cplx* trans_, int X, int Y)
{
  QMG_REQUIRE_INIT();
  int rc = check_gauge_dims(X, Y, "qmg_u1_gauge_transform: X and Y must be even and >= 2"); if (rc) return rc;
  cd* gauge = reinterpret_cast<cd*>(gauge_); const cd* trans = reinterpret_cast<const cd*>(trans_);
  const long V = (long)X * Y;
  HaloTemp halo; rc = halo.fetch(trans, 0, 1, X, Y, 1); if (rc) return rc;
  const FieldView<cd> g = make_view(trans, (const cd*)halo.ym, (const cd*)halo.yp, X, Y);
  return launch_ew(2 * V, [=] __device__(long e) {
    const int mu = e >= V ? 1 : 0;
    const long s = e - (long)mu * V;
    int x, y; site_xy(X, Y, s, x, y);
    const cd fwd = mu == 0 ? g.at(x + 1, y) : g.at(x, y + 1);
    gauge[e] = cmul(cmul(trans[s], gauge[e]), cconj(fwd));
  });
}

// n_iter sweeps of  U_mu(x) <- proj_U(1) [ U_mu(x) + alpha (upper staple + lower staple) ]   (u1_utils.h:276-383).
// Every sweep reads the previous field only (Jacobi), so `out` and a scratch copy alternate.
// textbook = 0 reproduces what the reference COMPUTES: its y-link section adds both y staples to the x link of the same
// site (`caxpy(alpha, link_vec, smeared_field, size_cm)` at u1_utils.h:352 and :372 lacks the `+ size_cm`), so an x link
// receives four staples and a y link is only re-projected.  textbook = 1 is the smearing the comments describe.
int qmg_u1_ape_smear(qmg_cplx* out_, const qmg_cplx* in_, int X, int Y, double alpha, int n_iter, int textbook)
{
  QMG_REQUIRE_INIT();
  int rc = check_gauge_dims(X, Y, "qmg_u1_ape_smear: X and Y must be even and >= 2"); if (rc) return rc;
  const long V = (long)X * Y;
  cd* out = reinterpret_cast<cd*>(out_);
  void* scratch_v = nullptr;
  rc = qmg_malloc(&scratch_v, sizeof(cd) * 2 * V); if (rc) return rc;
  cd* scratch = reinterpret_cast<cd*>(scratch_v);
  // arrange the ping-pong so that the last sweep lands in `out`
  cd* bufs[2] = { (n_iter % 2 == 0) ? out : scratch, (n_iter % 2 == 0) ? scratch : out };
  rc = qmg_copy(reinterpret_cast<qmg_cplx*>(bufs[0]), in_, 2 * V);
  for (int it = 0; it < n_iter && !rc; it++)
  {
    const cd* src = bufs[it & 1]; cd* dst = bufs[(it + 1) & 1];
    GaugeHalo halo; rc = halo.fetch_complex(src, X, Y); if (rc) break;
    const FieldView<cd> ux = make_view(src, (const cd*)halo.ym[0], (const cd*)halo.yp[0], X, Y);
    const FieldView<cd> uy = make_view(src + V, (const cd*)halo.ym[1], (const cd*)halo.yp[1], X, Y);
    rc = launch_ew(2 * V, [=] __device__(long e) {
      const int mu = e >= V ? 1 : 0;
      const long s = e - (long)mu * V;
      int x, y; site_xy(X, Y, s, x, y);
      cd st = cmake(0.0, 0.0);
      const bool x_staples = (mu == 0), y_staples = textbook ? (mu == 1) : (mu == 0);
      if (x_staples)
      {
        st = cadd(st, cmul(cmul(uy.at(x, y), ux.at(x, y + 1)), cconj(uy.at(x + 1, y))));
        st = cadd(st, cmul(cmul(cconj(uy.at(x, y - 1)), ux.at(x, y - 1)), uy.at(x + 1, y - 1)));
      }
      if (y_staples)
      {
        st = cadd(st, cmul(cmul(ux.at(x, y), uy.at(x + 1, y)), cconj(ux.at(x, y + 1))));
        st = cadd(st, cmul(cmul(cconj(ux.at(x - 1, y)), uy.at(x - 1, y)), ux.at(x - 1, y + 1)));
      }
      const cd u = src[e];
      const double re = u.x + alpha * st.x, im = u.y + alpha * st.y;
      double sn, cs; sincos(atan2(im, re), &sn, &cs);     // arg_vector then polar, as the reference projects
      dst[e] = cmake(cs, sn);
    });
  }
  qmg_free(scratch_v);
  return rc;
}

// n_update non-compact heatbath updates of the phase field at coupling beta (u1_utils.h:607-667): every link is redrawn
// from N(-staple / 2, 1 / (2 beta)), four independent subsets per update (see the header of this file).
// update0: number of updates this field has already received (offsets the random counter so successive calls continue
// one stream); returns through it nothing -- the caller keeps the count.
int qmg_u1_heatbath(double* phases, int X, int Y, double beta, int n_update, uint64_t seed, uint64_t update0)
{
  QMG_REQUIRE_INIT();
  int rc = check_gauge_dims(X, Y, "qmg_u1_heatbath: X and Y must be even and >= 2"); if (rc) return rc;
  if (!(beta > 0.0)) return fail_msg("qmg_u1_heatbath: beta must be positive");
  const long V = (long)X * Y;
  const int xh = X / 2;
  const double width = sqrt(0.5 / beta);
  const long y0_global = (long)qmg_comm_rank() * Y;       // the random counter is keyed by GLOBAL coordinates
  for (int u = 0; u < n_update; u++)
    for (int sub = 0; sub < 4; sub++)
    {
      GaugeHalo halo; rc = halo.fetch_real(phases, X, Y); if (rc) return rc;
      const FieldView<double> ax = make_view((const double*)phases, (const double*)halo.ym[0], (const double*)halo.yp[0], X, Y);
      const FieldView<double> ay = make_view((const double*)phases + V, (const double*)halo.ym[1], (const double*)halo.yp[1], X, Y);
      const int mu = sub >> 1, cls = sub & 1;
      const uint64_t step = (update0 + (uint64_t)u) * 2 + (uint64_t)mu;
      // one thread per link of the subset: (row or column class, position); V/2 links
      rc = launch_ew(V / 2, [=] __device__(long t) {
        int x, y;
        if (mu == 0)
        {
          // x links on rows y = cls (mod 2): t -> (row index, x)
          y = 2 * (int)(t / X) + cls;
          x = (int)(t % X);
        }
        else
        {
          // y links on columns x = cls (mod 2): t -> (y, column index); consecutive threads walk one memory row
          y = (int)(t / xh);
          x = 2 * (int)(t % xh) + cls;
        }
        double staple;
        if (mu == 0)
          staple = ay.at(x + 1, y) - ax.at(x, y + 1) - ay.at(x, y) - ay.at(x + 1, y - 1) - ax.at(x, y - 1) + ay.at(x, y - 1);
        else
          staple = ax.at(x, y + 1) - ay.at(x + 1, y) - ax.at(x, y) - ax.at(x - 1, y + 1) - ay.at(x - 1, y) + ax.at(x - 1, y);
        double n0, n1;
        philox_normal2(seed, (uint64_t)(y0_global + y) * (uint64_t)X + (uint64_t)x, step, n0, n1);
        const int par = (x + y) & 1;
        phases[(size_t)mu * V + (size_t)(y + par * Y) * xh + (x >> 1)] = width * n0 - 0.5 * staple;
      });
      if (rc) return rc;
    }
  return 0;
}

} // extern "C"
