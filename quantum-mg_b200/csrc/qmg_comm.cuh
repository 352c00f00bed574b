// y-slab sharding over the GPUs of one node: one process per GPU, every rank owns rows
// [rank Yl, (rank+1) Yl) of the global lattice as a self-contained even-odd (X, Yl) lattice.
// The only data-path exchange is ONE boundary row per field per neighbour-dependent operation (the "Becomes MPI"
// periodic-boundary loops of /root/reference/cshift/cshift_2d.h:94-119), plus all-reduces of the 1..k scalars every
// dot / norm produces.
#pragma once
#include "qmg_lattice.cuh"

namespace qmg {

struct Comm
{
  bool active = false;             // halo rows come from the ring neighbours (or from this slab itself in loopback mode)
  bool loopback = false;           // one rank exchanging with itself through the same code path (single-GPU testing)
  int nranks = 1, rank = 0;
  void* nccl = nullptr;            // ncclComm_t
  cudaStream_t stream = nullptr;   // exchange stream, runs beside the interior stencil kernel
  cudaEvent_t ev_main = nullptr, ev_done = nullptr;
  // staging for the halo rows (grown on demand): 2 rows out, 2 rows in
  cd* send_lo = nullptr; cd* send_hi = nullptr; cd* recv_ym = nullptr; cd* recv_yp = nullptr;
  size_t row_cap = 0;              // capacity of each buffer in complex elements
  long halo_exchanges = 0;
  long allreduces = 0;
  // peer mailboxes of the in-kernel all-reduce (RedState, qmg_common.cuh): own block + the peers' blocks, IPC-mapped
  bool p2p = false;
  unsigned long long halo_seq = 0;      // p2p halo exchanges so far (same on every rank: exchanges are collective)
  unsigned int* d_halo_counter = nullptr;
  long p2p_halo_exchanges = 0;
  double* mail_local = nullptr;
  double* mail[kMaxRanks] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
};
int upload_red_state(int nranks, int rank, int p2p, double* const* mail);   // qmg_runtime.cu
long long p2p_watchdog_cycles();                                            // qmg_runtime.cu
Comm& comm();

// rows -1 and Y of a field as received from the ring neighbours, layout (parity, x/2, dof)
struct HaloRows { const cd* ym = nullptr; const cd* yp = nullptr; };

// Exchange the boundary rows of `field` (dof complex per site on an X x Y slab) with the ring neighbours on the MAIN
// stream; out_ym / out_yp (each X*dof complex, layout (parity, x/2, dof)) receive rows -1 and Y.
// parity_mask: bit p set = rows of parity-half p are exchanged (a half-length vector only has p = 0).
int halo_exchange_sync(const cd* field, int X, int Y, int dof, cd* out_ym, cd* out_yp, int parity_mask = 3);
// Overlapped flavour for the stencil: pack + exchange run on the exchange stream, ordered after everything queued so far
// on the main stream; the caller launches the interior rows, then calls halo_exchange_end() before the boundary rows.
int halo_exchange_begin(const cd* field, int X, int Y, int dof, int parity_mask, HaloRows* out);
int halo_exchange_end();
int allreduce_result(double* d_buf, int count, int op_max);

// Temporary halo rows of a set-up field (link fills, variant builders, coarse build): fetched once, released on scope exit.
struct HaloTemp
{
  cd* ym = nullptr; cd* yp = nullptr;
  HaloTemp() {}
  ~HaloTemp();
  // nfields fields of `dof` complex per site, field f at base + f * field_stride; rows of field f at ym/yp + f * X * dof
  int fetch(const cd* base, long field_stride, int nfields, int X, int Y, int dof);
  int alloc_rows(int X, int dof, int nfields = 1);   // just the storage (the caller runs the exchange)
  HaloRows rows(int f, int X, int dof) const { HaloRows h; if (ym) { h.ym = ym + (size_t)f * X * dof; h.yp = yp + (size_t)f * X * dof; } return h; }
private:
  HaloTemp(const HaloTemp&); HaloTemp& operator=(const HaloTemp&);
};

// pointer to the first dof of the neighbour of (p, y, k) in direction mu, looking into the halo rows at the slab edges
__device__ __forceinline__ const cd* nbr_site_ptr(const cd* base, const HaloRows& h, const Geom& g, int p, int y, int k, int mu, int dof)
{
  const int q = 1 - p;
  if (mu == 1 && h.yp != nullptr && y == g.Y - 1) return h.yp + ((size_t)q * g.xh + k) * dof;
  if (mu == 3 && h.ym != nullptr && y == 0) return h.ym + ((size_t)q * g.xh + k) * dof;
  return base + ((size_t)q * g.half + nbr_h(g, p, y, k, mu)) * dof;
}

} // namespace qmg
