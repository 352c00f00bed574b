// quantum-mg on B200 -- Lattice2D: the even-odd (eo, y, x, c) layout contract
// (/root/reference/lattice/lattice.h:12-396).  Host-side integer geometry only; the kernels carry
// the same arithmetic in quantum-mg_b200/csrc/qmg_lattice.cuh.
//
//   site   i = (y + parity Y) X/2 + x/2,  parity = (x + y) mod 2   (all even sites, then all odd)
//   field  element (site i, local index l of `per_site` values, block b of size volume*per_site):
//          b * volume * per_site + i * per_site + l
// ColorVector: per_site = nc; ColorMatrix: nc^2 with l = c1 nc + c2; Gauge / Hopping / Corner:
// ColorMatrix blocks b = mu (x, y | +x, +y, -x, -y | the four diagonals).
#ifndef QMG_B200_LATTICE
#define QMG_B200_LATTICE

class Lattice2D
{
private:
  int dims[2];
  int nc;
  long volume;

  inline long block(long i, long l, long per_site, long b) const { return (b * volume + i) * per_site + l; }
  inline void unblock(long idx, long per_site, long& i, long& l, long& b) const
  {
    b = idx / (volume * per_site); idx -= b * volume * per_site;
    i = idx / per_site; l = idx - i * per_site;
  }

public:
  Lattice2D(int xlen, int ylen, int my_nc) : nc(my_nc) { dims[0] = xlen; dims[1] = ylen; volume = (long)xlen * ylen; }
  Lattice2D(const Lattice2D& o) : nc(o.nc) { dims[0] = o.dims[0]; dims[1] = o.dims[1]; volume = o.volume; }
  ~Lattice2D() { }

  // lattice.h:60: change the dof per site in place
  void update_nc(int my_nc) { nc = my_nc; }

  // ---- coordinates -> indices (lattice.h:75-182)
  inline int coord_to_index(int x, int y) const
  {
    if (volume == 1) return 0;
    const int parity = (x + y) & 1;
    return (int)((long)(y + parity * dims[1]) * dims[0] / 2 + (x / 2) % (dims[0] / 2));
  }
  inline int coord_to_index(int* c) const { return coord_to_index(c[0], c[1]); }
  inline int dof_coord_to_index(int total_dof, int x, int y, int dof) const { return (int)block(coord_to_index(x, y), dof, total_dof, 0); }
  inline int dof_coord_to_index(int total_dof, int* c, int dof) const { return dof_coord_to_index(total_dof, c[0], c[1], dof); }
  inline int dof_coord_to_index(int total_dof, int i, int dof) const { return (int)block(i, dof, total_dof, 0); }
  inline int cv_coord_to_index(int x, int y, int c) const { return (int)block(coord_to_index(x, y), c, nc, 0); }
  inline int cv_coord_to_index(int* xy, int c) const { return cv_coord_to_index(xy[0], xy[1], c); }
  inline int cv_coord_to_index(int i, int c) const { return (int)block(i, c, nc, 0); }
  inline int vol_index_dof_to_cv_index(int i, int c) const { return (int)block(i, c, nc, 0); }
  inline int cm_coord_to_index(int x, int y, int c1, int c2) const { return (int)block(coord_to_index(x, y), c1 * nc + c2, nc * nc, 0); }
  inline int cm_coord_to_index(int* xy, int c1, int c2) const { return cm_coord_to_index(xy[0], xy[1], c1, c2); }
  inline int cm_coord_to_index(int i, int c1, int c2) const { return (int)block(i, c1 * nc + c2, nc * nc, 0); }
  inline int gauge_coord_to_index(int x, int y, int c1, int c2, int mu) const { return (int)block(coord_to_index(x, y), c1 * nc + c2, nc * nc, mu); }
  inline int gauge_coord_to_index(int* xy, int c1, int c2, int mu) const { return gauge_coord_to_index(xy[0], xy[1], c1, c2, mu); }
  inline int gauge_coord_to_index(int i, int c1, int c2, int mu) const { return (int)block(i, c1 * nc + c2, nc * nc, mu); }
  inline int hopping_coord_to_index(int x, int y, int c1, int c2, int mu) const { return gauge_coord_to_index(x, y, c1, c2, mu); }
  inline int hopping_coord_to_index(int* xy, int c1, int c2, int mu) const { return gauge_coord_to_index(xy[0], xy[1], c1, c2, mu); }
  inline int hopping_coord_to_index(int i, int c1, int c2, int mu) const { return gauge_coord_to_index(i, c1, c2, mu); }
  inline int corner_coord_to_index(int x, int y, int c1, int c2, int munu) const { return gauge_coord_to_index(x, y, c1, c2, munu); }
  inline int corner_coord_to_index(int* xy, int c1, int c2, int munu) const { return gauge_coord_to_index(xy[0], xy[1], c1, c2, munu); }
  inline int corner_coord_to_index(int i, int c1, int c2, int munu) const { return gauge_coord_to_index(i, c1, c2, munu); }

  // ---- indices -> coordinates (lattice.h:199-283)
  inline void index_to_coord(int i, int& x, int& y) const
  {
    if (volume == 1) { x = y = 0; return; }
    const int xh = dims[0] / 2;
    const int parity = (int)(i / (volume / 2));
    y = i / xh - parity * dims[1];
    x = 2 * (i % xh) + ((y + parity) & 1);
  }
  inline void index_to_coord(int i, int* xy) const { index_to_coord(i, xy[0], xy[1]); }
  inline void dof_index_to_coord(int idx, int total_dof, int& x, int& y, int& dof) const
  { long i, l, b; unblock(idx, total_dof, i, l, b); index_to_coord((int)i, x, y); dof = (int)l; }
  inline void dof_index_to_coord(int idx, int total_dof, int* xy, int& dof) const { dof_index_to_coord(idx, total_dof, xy[0], xy[1], dof); }
  inline void cv_index_to_coord(int idx, int& x, int& y, int& c) const { dof_index_to_coord(idx, nc, x, y, c); }
  inline void cv_index_to_coord(int idx, int* xy, int& c) const { dof_index_to_coord(idx, nc, xy[0], xy[1], c); }
  inline void cm_index_to_coord(int idx, int& x, int& y, int& c1, int& c2) const
  { long i, l, b; unblock(idx, (long)nc * nc, i, l, b); index_to_coord((int)i, x, y); c1 = (int)(l / nc); c2 = (int)(l % nc); }
  inline void cm_index_to_coord(int idx, int* xy, int& c1, int& c2) const { cm_index_to_coord(idx, xy[0], xy[1], c1, c2); }
  inline void gauge_index_to_coord(int idx, int& x, int& y, int& c1, int& c2, int& mu) const
  { long i, l, b; unblock(idx, (long)nc * nc, i, l, b); index_to_coord((int)i, x, y); c1 = (int)(l / nc); c2 = (int)(l % nc); mu = (int)b; }
  inline void gauge_index_to_coord(int idx, int* xy, int& c1, int& c2, int& mu) const { gauge_index_to_coord(idx, xy[0], xy[1], c1, c2, mu); }
  inline void hopping_index_to_coord(int idx, int& x, int& y, int& c1, int& c2, int& mu) const { gauge_index_to_coord(idx, x, y, c1, c2, mu); }
  inline void hopping_index_to_coord(int idx, int* xy, int& c1, int& c2, int& mu) const { gauge_index_to_coord(idx, xy[0], xy[1], c1, c2, mu); }
  inline void corner_index_to_coord(int idx, int& x, int& y, int& c1, int& c2, int& munu) const { gauge_index_to_coord(idx, x, y, c1, c2, munu); }
  inline void corner_index_to_coord(int idx, int* xy, int& c1, int& c2, int& munu) const { gauge_index_to_coord(idx, xy[0], xy[1], c1, c2, munu); }

  // ---- parity queries.  The reference's index predicates test "i > size/2", i.e. they report the ODD half and
  // are off by one at the boundary (lattice.h:288-316); kept so that a driver relying on them sees the same answers.
  inline bool index_is_even(int i) const { return i > volume / 2; }
  inline bool cv_index_is_even(int i) const { return i > get_size_cv() / 2; }
  inline bool cm_index_is_even(int i) const { return i > get_size_cm() / 2; }
  inline bool gauge_index_is_even(int i) const { return i > get_size_gauge() / 2; }
  inline bool hopping_index_is_even(int i) const { return i > get_size_hopping() / 2; }
  inline bool corner_index_is_even(int i) const { return i > get_size_corner() / 2; }
  inline bool coord_is_even(int x, int y) const { return ((x + y) & 1) == 0; }

  // ---- sizes (lattice.h:327-394)
  inline void get_dim(int* out) const { out[0] = dims[0]; out[1] = dims[1]; }
  inline int get_dim_mu(int mu) const { return (mu >= 0 && mu < 2) ? dims[mu] : -1; }
  inline int get_nd() const { return 2; }
  inline int get_nc() const { return nc; }
  inline int get_nc_nc() const { return nc * nc; }
  inline long get_volume() const { return volume; }
  inline long get_size_dof(int total_dof) const { return volume * total_dof; }
  inline long get_size_cv() const { return volume * nc; }
  inline long get_size_cm() const { return volume * nc * nc; }
  inline long get_size_gauge() const { return volume * nc * nc * 2; }
  inline long get_size_hopping() const { return volume * nc * nc * 4; }
  inline long get_size_corner() const { return volume * nc * nc * 4; }
};

#endif
