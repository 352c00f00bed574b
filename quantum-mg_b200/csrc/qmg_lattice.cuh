// Even-odd index arithmetic shared by the kernels
// (the layout contract of /root/reference/lattice/lattice.h:75-81,199-205 and
// the neighbour semantics of /root/reference/cshift/cshift_2d.h:45-222).
#pragma once
#include "qmg_common.cuh"

namespace qmg {

// A site is (parity p, row y, column-in-parity k) with x = 2k + ((y+p)&1).
// Its index inside its parity half is h = y*xh + k; the full site index is p*half + h.
struct Geom
{
  int xh;        // X/2
  int Y;
  unsigned half; // xh*Y
};

// In-parity index (inside parity 1-p) of the neighbour of (p,y,k) in direction mu (+x,+y,-x,-y).
__host__ __device__ __forceinline__ unsigned nbr_h(const Geom& g, int p, int y, int k, int mu)
{
  const int sft = (y + p) & 1;
  int yy = y, kk = k;
  if (mu == 0) { kk = k + sft; if (kk == g.xh) kk = 0; }
  else if (mu == 2) { kk = k - 1 + sft; if (kk < 0) kk = g.xh - 1; }
  else if (mu == 1) { yy = (y + 1 == g.Y) ? 0 : y + 1; }
  else { yy = (y == 0) ? g.Y - 1 : y - 1; }
  return (unsigned)yy * g.xh + kk;
}

__host__ __device__ __forceinline__ int opposite_dir(int mu) { return (mu + 2) & 3; }

} // namespace qmg
