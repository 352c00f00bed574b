// NCCL plumbing of the y-slab sharding.  NCCL is resolved at run time with dlopen: when the process already holds
// a libnccl.so.2 (torch's bundled copy under torchrun) that one is used, so there are never two NCCLs in one process,
// and a single-GPU or CPU-only process never needs the library at all.
#include "qmg_comm.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace qmg {

Comm& comm() { static Comm c; return c; }

namespace {

typedef struct { char internal[128]; } NcclUniqueId;   // NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;
enum { kNcclSum = 0, kNcclMax = 2, kNcclChar = 0, kNcclDouble = 8 };

struct NcclApi
{
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi& api() { static NcclApi a; return a; }

int load_nccl()
{
  NcclApi& a = api();
  if (a.handle != nullptr) return 0;
  a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (a.handle == nullptr) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (a.handle == nullptr) return fail_msg("qmg_comm: libnccl.so.2 not found");
#define QMG_SYM(field, name) *(void**)(&a.field) = dlsym(a.handle, name); if (a.field == nullptr) return fail_msg("qmg_comm: NCCL symbol missing: " name)
  QMG_SYM(GetUniqueId, "ncclGetUniqueId"); QMG_SYM(CommInitRank, "ncclCommInitRank"); QMG_SYM(CommDestroy, "ncclCommDestroy");
  QMG_SYM(AllReduce, "ncclAllReduce"); QMG_SYM(AllGather, "ncclAllGather"); QMG_SYM(Send, "ncclSend"); QMG_SYM(Recv, "ncclRecv");
  QMG_SYM(GroupStart, "ncclGroupStart"); QMG_SYM(GroupEnd, "ncclGroupEnd"); QMG_SYM(GetErrorString, "ncclGetErrorString");
#undef QMG_SYM
  return 0;
}

int nccl_fail(const char* what, int rc)
{
  char buf[256];
  snprintf(buf, sizeof(buf), "NCCL failure in %s: %s", what, api().GetErrorString ? api().GetErrorString(rc) : "?");
  return fail_msg(buf);
}
#define QMG_NCCL(call) do { int rc__ = (call); if (rc__ != 0) return nccl_fail(#call, rc__); } while (0)

// rows y = 0 and y = Y-1 of a (parity, y, x/2, dof) field, packed (parity, x/2, dof); only the parity halves in the mask
__global__ void __launch_bounds__(256) pack_rows_kernel(const cd* __restrict__ field, cd* __restrict__ lo, cd* __restrict__ hi, int rowlen, long half_elems, int Y, int parity_mask)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * rowlen) return;
  const int p = i / rowlen, j = i - p * rowlen;
  if (!((parity_mask >> p) & 1)) return;
  lo[i] = field[(long)p * half_elems + j];
  hi[i] = field[(long)p * half_elems + (long)(Y - 1) * rowlen + j];
}

int ensure_rows(size_t elems)
{
  Comm& c = comm();
  if (elems <= c.row_cap) return 0;
  QMG_CUDA(cudaDeviceSynchronize());
  cd** bufs[] = { &c.send_lo, &c.send_hi, &c.recv_ym, &c.recv_yp };
  for (int i = 0; i < 4; i++)
  {
    if (*bufs[i] != nullptr) cudaFree(*bufs[i]);
    QMG_CUDA(cudaMalloc(bufs[i], sizeof(cd) * elems));
  }
  c.row_cap = elems;
  return 0;
}

// ring exchange of packed rows on stream s: my top row goes up and arrives as the upper rank's row -1, my bottom row
// goes down and arrives as the lower rank's row Y.  `first` / `count`: the span of the packed rows that is live.
int ring_exchange(const cd* lo, const cd* hi, cd* ym, cd* yp, size_t first, size_t count, cudaStream_t s)
{
  Comm& c = comm();
  c.halo_exchanges++;
  if (c.loopback)
  {
    QMG_CUDA(cudaMemcpyAsync(ym + first, hi + first, sizeof(cd) * count, cudaMemcpyDeviceToDevice, s));
    QMG_CUDA(cudaMemcpyAsync(yp + first, lo + first, sizeof(cd) * count, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  NcclApi& a = api();
  const int up = (c.rank + 1) % c.nranks, down = (c.rank + c.nranks - 1) % c.nranks;
  const size_t n = 2 * count;   // doubles
  QMG_NCCL(a.GroupStart());
  QMG_NCCL(a.Send(hi + first, n, kNcclDouble, up, c.nccl, s));
  QMG_NCCL(a.Recv(ym + first, n, kNcclDouble, down, c.nccl, s));
  QMG_NCCL(a.Send(lo + first, n, kNcclDouble, down, c.nccl, s));
  QMG_NCCL(a.Recv(yp + first, n, kNcclDouble, up, c.nccl, s));
  QMG_NCCL(a.GroupEnd());
  return 0;
}

// The boundary rows written STRAIGHT into the ring neighbours' halo mailboxes (NVLink peer stores), no send / receive
// kernels: every thread stores its elements of row Y-1 into the upper neighbour's "row -1" slot and of row 0 into the
// lower neighbour's "row Y" slot and fences; the last block to finish raises the sequence number in both neighbours'
// flags and then waits for the two flags of its OWN mailbox, so when the kernel retires the neighbours' rows are here.
// Two slots alternate: a neighbour can only be one exchange ahead (it waits for my flag of every exchange).
__global__ void __launch_bounds__(256) halo_p2p_kernel(const cd* __restrict__ field, int rowlen, long half_elems, int Y, int parity_mask,
                                                       cd* up_ym, cd* down_yp, unsigned long long* up_flag, unsigned long long* down_flag,
                                                       const unsigned long long* my_flags, unsigned long long seq, unsigned int* counter, int rank,
                                                       long long watchdog_cycles, unsigned long long* host_err)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * rowlen)
  {
    const int p = i / rowlen, j = i - p * rowlen;
    if ((parity_mask >> p) & 1)
    {
      down_yp[i] = field[(long)p * half_elems + j];
      up_ym[i] = field[(long)p * half_elems + (long)(Y - 1) * rowlen + j];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0)
  {
    *counter = 0u;
    __threadfence_system();
    st_sys_u64(up_flag, seq);
    st_sys_u64(down_flag, seq);
  }
  if (threadIdx.x < 2)
  {
    const long long t0 = clock64();
    while (ld_sys_u64(my_flags + threadIdx.x) < seq)
      if (clock64() - t0 > watchdog_cycles) { st_sys_u64(host_err, 257ull + threadIdx.x); break; }   // no trap: the host turns this into an error code (peer_timeout_error)
  }
}

static inline unsigned long long* halo_flags(double* block, int slot) { return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(block) + kHaloOffset) + 2 * slot; }
static inline cd* halo_slot(double* block, int slot, int side) { return reinterpret_cast<cd*>(reinterpret_cast<char*>(block) + kHaloDataOffset) + ((size_t)slot * 2 + side) * kHaloCap; }

// Exchange the boundary rows of `field` on stream s.  out_ym / out_yp: where the caller wants rows -1 / Y (or nullptr:
// leave them where they arrive).  landed: where they are when s reaches this point.
int exchange_on(cudaStream_t s, const cd* field, int X, int Y, int dof, int parity_mask, cd* out_ym, cd* out_yp, HaloRows* landed)
{
  Comm& c = comm();
  parity_mask &= 3;
  const int rowlen = (X / 2) * dof;
  const size_t first = (parity_mask == 2) ? (size_t)rowlen : 0;
  const size_t count = (parity_mask == 3) ? (size_t)2 * rowlen : (size_t)rowlen;
  if (c.p2p && !c.loopback && (size_t)2 * rowlen <= kHaloCap)
  {
    const unsigned long long seq = ++c.halo_seq;
    const int slot = (int)(seq & 1);
    const int up = (c.rank + 1) % c.nranks, down = (c.rank + c.nranks - 1) % c.nranks;
    cd* my_ym = halo_slot(c.mail[c.rank], slot, 0); cd* my_yp = halo_slot(c.mail[c.rank], slot, 1);
    if (parity_mask != 0)
    {
      halo_p2p_kernel<<<(2 * rowlen + 255) / 256, 256, 0, s>>>(field, rowlen, (long)rowlen * Y, Y, parity_mask,
          halo_slot(c.mail[up], slot, 0), halo_slot(c.mail[down], slot, 1), halo_flags(c.mail[up], slot) + 0, halo_flags(c.mail[down], slot) + 1,
          halo_flags(c.mail[c.rank], slot), seq, c.d_halo_counter, c.rank, p2p_watchdog_cycles(), rt().h_err);
      QMG_LAUNCH_CHECK();
      c.halo_exchanges++; c.p2p_halo_exchanges++;
      if (out_ym != nullptr) QMG_CUDA(cudaMemcpyAsync(out_ym + first, my_ym + first, sizeof(cd) * count, cudaMemcpyDeviceToDevice, s));
      if (out_yp != nullptr) QMG_CUDA(cudaMemcpyAsync(out_yp + first, my_yp + first, sizeof(cd) * count, cudaMemcpyDeviceToDevice, s));
    }
    landed->ym = out_ym != nullptr ? out_ym : my_ym;
    landed->yp = out_yp != nullptr ? out_yp : my_yp;
    return 0;
  }
  int rc = ensure_rows((size_t)X * dof); if (rc) return rc;
  cd* ym = out_ym != nullptr ? out_ym : c.recv_ym;
  cd* yp = out_yp != nullptr ? out_yp : c.recv_yp;
  landed->ym = ym; landed->yp = yp;
  if (parity_mask == 0) return 0;
  pack_rows_kernel<<<(2 * rowlen + 255) / 256, 256, 0, s>>>(field, c.send_lo, c.send_hi, rowlen, (long)rowlen * Y, Y, parity_mask);
  QMG_LAUNCH_CHECK();
  return ring_exchange(c.send_lo, c.send_hi, ym, yp, first, count, s);
}

} // namespace

int halo_exchange_sync(const cd* field, int X, int Y, int dof, cd* out_ym, cd* out_yp, int parity_mask)
{
  Comm& c = comm();
  if (!c.active) return fail_msg("halo exchange requested without an active communicator");
  HaloRows landed;
  return exchange_on(rt().stream, field, X, Y, dof, parity_mask, out_ym, out_yp, &landed);
}

int halo_exchange_begin(const cd* field, int X, int Y, int dof, int parity_mask, HaloRows* out)
{
  Comm& c = comm();
  if (!c.active) return fail_msg("halo exchange requested without an active communicator");
  QMG_CUDA(cudaEventRecord(c.ev_main, rt().stream));
  QMG_CUDA(cudaStreamWaitEvent(c.stream, c.ev_main, 0));
  int rc = exchange_on(c.stream, field, X, Y, dof, parity_mask, nullptr, nullptr, out); if (rc) return rc;
  QMG_CUDA(cudaEventRecord(c.ev_done, c.stream));
  return 0;
}
int halo_exchange_end()
{
  QMG_CUDA(cudaStreamWaitEvent(rt().stream, comm().ev_done, 0));
  return 0;
}

int allreduce_result(double* d_buf, int count, int op_max)
{
  Comm& c = comm();
  if (!c.active || c.loopback || c.p2p) return 0;      // p2p: the reducing kernel already summed over the ranks
  c.allreduces++;
  QMG_NCCL(api().AllReduce(d_buf, d_buf, (size_t)count, kNcclDouble, op_max ? kNcclMax : kNcclSum, c.nccl, rt().stream));
  return 0;
}

// Peer mailboxes for the in-kernel all-reduce: every rank allocates one block, the IPC handles go round with one
// ncclAllGather, and each rank maps the other ranks' blocks (NVLink peer memory).  All ranks switch together: if any
// rank cannot map a peer, everybody stays on ncclAllReduce.
static int setup_mailboxes(bool want)
{
  // Every rank runs the SAME collectives whatever happens locally (QMG_P2P=0 on this rank only, an allocation or a
  // mapping that fails): a local problem votes "no" in the unanimity all-reduce instead of leaving the peers waiting.
  Comm& c = comm();
  NcclApi& a = api();
  cudaStream_t s = rt().stream;
  int ok = want ? 1 : 0;
  if (ok && cudaMalloc(&c.mail_local, kPeerBlockBytes) != cudaSuccess) { cudaGetLastError(); c.mail_local = nullptr; ok = 0; }
  if (ok && cudaMalloc(&c.d_halo_counter, sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); c.d_halo_counter = nullptr; ok = 0; }
  if (ok && cudaMemset(c.d_halo_counter, 0, sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMemset(c.mail_local, 0, kHaloDataOffset) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  if (ok && cudaIpcGetMemHandle(&mine, c.mail_local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  // the handles and the vote travel in device scratch; without it this rank cannot take part in any collective at all
  char* d_handles = nullptr;
  QMG_CUDA(cudaMalloc(&d_handles, sizeof(mine) * (c.nranks + 1) + sizeof(double)));
  double* d_ok = reinterpret_cast<double*>(d_handles + sizeof(mine) * (c.nranks + 1));
  std::vector<cudaIpcMemHandle_t> all(c.nranks);
  bool comm_failed = false;
  if (cudaMemcpy(d_handles + sizeof(mine) * c.nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  if (a.AllGather(d_handles + sizeof(mine) * c.nranks, d_handles, sizeof(mine), kNcclChar, c.nccl, s) != 0) comm_failed = true;
  if (!comm_failed && cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); comm_failed = true; }
  if (!comm_failed && cudaMemcpy(all.data(), d_handles, sizeof(mine) * c.nranks, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  for (int r = 0; r < c.nranks && ok && !comm_failed; r++)
  {
    if (r == c.rank) { c.mail[r] = c.mail_local; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    c.mail[r] = reinterpret_cast<double*>(p);
  }
  // unanimous?
  const double mine_ok = ok;
  double total = 0.0;
  if (!comm_failed && cudaMemcpy(d_ok, &mine_ok, sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); comm_failed = true; }
  if (!comm_failed && a.AllReduce(d_ok, d_ok, 1, kNcclDouble, kNcclSum, c.nccl, s) != 0) comm_failed = true;
  if (!comm_failed && cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); comm_failed = true; }
  if (!comm_failed && cudaMemcpy(&total, d_ok, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); comm_failed = true; }
  cudaFree(d_handles);
  if (comm_failed) return fail_msg("qmg_comm_init: the NCCL rendezvous of the peer mailboxes failed");
  c.p2p = ((int)(total + 0.5) == c.nranks);
  return 0;
}

static void release_mailboxes()
{
  Comm& c = comm();
  for (int r = 0; r < kMaxRanks; r++)
  {
    if (c.mail[r] != nullptr && c.mail[r] != c.mail_local) cudaIpcCloseMemHandle(c.mail[r]);
    c.mail[r] = nullptr;
  }
  if (c.mail_local != nullptr) cudaFree(c.mail_local);
  if (c.d_halo_counter != nullptr) cudaFree(c.d_halo_counter);
  c.mail_local = nullptr; c.d_halo_counter = nullptr; c.p2p = false; c.halo_seq = 0;
  cudaGetLastError();
}

HaloTemp::~HaloTemp() { qmg_free(ym); qmg_free(yp); }

int HaloTemp::alloc_rows(int X, int dof, int nfields)
{
  const size_t row = (size_t)X * dof;
  int rc = qmg_malloc((void**)&ym, sizeof(cd) * row * nfields); if (rc) return rc;
  return qmg_malloc((void**)&yp, sizeof(cd) * row * nfields);
}

int HaloTemp::fetch(const cd* base, long field_stride, int nfields, int X, int Y, int dof)
{
  if (!comm().active) return 0;          // periodic inside this lattice: rows() stays empty
  const size_t row = (size_t)X * dof;
  int rc = alloc_rows(X, dof, nfields); if (rc) return rc;
  for (int f = 0; f < nfields; f++)
  {
    rc = halo_exchange_sync(base + (size_t)f * field_stride, X, Y, dof, ym + f * row, yp + f * row, 3);
    if (rc) return rc;
  }
  return 0;
}

} // namespace qmg

using namespace qmg;

extern "C" {

int qmg_comm_unique_id(void* out128)
{
  int rc = load_nccl(); if (rc) return rc;
  NcclUniqueId id;
  QMG_NCCL(api().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int qmg_comm_init(int nranks, int rank, const void* unique_id128)
{
  QMG_REQUIRE_INIT();
  Comm& c = comm();
  if (c.active && c.loopback) { int frc = qmg_comm_finalize(); if (frc) return frc; }    // a real communicator replaces the loopback
  if (c.active) return fail_msg("qmg_comm_init: already initialised");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail_msg("qmg_comm_init: bad rank / size");
  c.nranks = nranks; c.rank = rank;
  if (nranks == 1) return 0;       // a single slab is the plain periodic lattice (see qmg_comm_set_loopback)
  int rc = load_nccl(); if (rc) return rc;
  NcclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  QMG_NCCL(api().CommInitRank(&c.nccl, nranks, id, rank));
  QMG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  QMG_CUDA(cudaEventCreateWithFlags(&c.ev_main, cudaEventDisableTiming));
  QMG_CUDA(cudaEventCreateWithFlags(&c.ev_done, cudaEventDisableTiming));
  c.active = true;
  // in-kernel all-reduce over NVLink peer memory unless QMG_P2P=0 (then every reduction calls ncclAllReduce)
  const char* env = getenv("QMG_P2P");
  rc = setup_mailboxes(!(env != nullptr && env[0] == '0') && nranks <= kMaxRanks); if (rc) return rc;
  if (!c.p2p) release_mailboxes();
  return upload_red_state(nranks, rank, c.p2p ? 1 : 0, c.mail);
}

// One rank exchanging with itself: the slab IS the periodic lattice, but every neighbour access across y = 0 / Y-1
// goes through the pack / exchange / halo-row code path of the sharded build (device copies instead of NCCL).
int qmg_comm_set_loopback(int on)
{
  QMG_REQUIRE_INIT();
  Comm& c = comm();
  if (c.active && !c.loopback) return fail_msg("qmg_comm_set_loopback: a multi-rank communicator is active");
  if (on && !c.active)
  {
    QMG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    QMG_CUDA(cudaEventCreateWithFlags(&c.ev_main, cudaEventDisableTiming));
    QMG_CUDA(cudaEventCreateWithFlags(&c.ev_done, cudaEventDisableTiming));
    c.nranks = 1; c.rank = 0; c.loopback = true; c.active = true;
    return 0;
  }
  if (!on && c.active) return qmg_comm_finalize();
  return 0;
}

int qmg_comm_finalize(void)
{
  Comm& c = comm();
  if (!c.active) { c.nranks = 1; c.rank = 0; return 0; }
  cudaDeviceSynchronize();
  if (!c.loopback)
  {
    // nobody may still be writing into a mailbox that is about to be unmapped: one last collective as a barrier
    if (c.p2p) { double* d = nullptr; if (cudaMalloc(&d, sizeof(double)) == cudaSuccess) { cudaMemset(d, 0, sizeof(double)); api().AllReduce(d, d, 1, kNcclDouble, kNcclSum, c.nccl, rt().stream); cudaStreamSynchronize(rt().stream); cudaFree(d); } }
    release_mailboxes();
    upload_red_state(1, 0, 0, nullptr);
    api().CommDestroy(c.nccl);
  }
  cudaStreamDestroy(c.stream); cudaEventDestroy(c.ev_main); cudaEventDestroy(c.ev_done);
  cd* bufs[] = { c.send_lo, c.send_hi, c.recv_ym, c.recv_yp };
  for (int i = 0; i < 4; i++) if (bufs[i] != nullptr) cudaFree(bufs[i]);
  c = Comm();
  return 0;
}

int qmg_comm_size(void) { return comm().nranks; }
int qmg_comm_rank(void) { return comm().rank; }
int qmg_comm_active(void) { return comm().active ? 1 : 0; }
long qmg_comm_halo_exchanges(void) { return comm().halo_exchanges; }
long qmg_comm_allreduces(void) { return comm().allreduces; }
int qmg_comm_p2p(void) { return comm().p2p ? 1 : 0; }
long qmg_comm_p2p_halo_exchanges(void) { return comm().p2p_halo_exchanges; }

// rows -1 and Y of a field of `dof` complex per site on this rank's X x Y slab (setup helpers, tests)
int qmg_halo_exchange(const qmg_cplx* field, int X, int Y, int dof, qmg_cplx* out_ym, qmg_cplx* out_yp)
{
  QMG_REQUIRE_INIT();
  if (!comm().active) return fail_msg("qmg_halo_exchange: not sharded (qmg_comm_init with more than one rank first)");
  return halo_exchange_sync(reinterpret_cast<const cd*>(field), X, Y, dof, reinterpret_cast<cd*>(out_ym), reinterpret_cast<cd*>(out_yp), 3);
}

} // extern "C"
