// Runtime plumbing of libqmg_b200: device selection, stream, memory, error text.
#include "qmg_common.cuh"
#include <string.h>
#include <stdlib.h>
#include <time.h>
#include <sched.h>
#include <algorithm>
#include <map>
#include <unordered_map>
#include <vector>

namespace qmg {

Runtime& rt() { static Runtime r; return r; }

struct ProfEntry { double seconds = 0.0; long calls = 0; };
static std::map<std::string, ProfEntry>& prof_table() { static std::map<std::string, ProfEntry> t; return t; }
static thread_local int prof_depth = 0;
static double now_seconds() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

void ProfScope::detail(int nc, int X, int Y) { if (on) snprintf(tag, sizeof(tag), " nc%d %dx%d", nc, X, Y); }

ProfScope::ProfScope(const char* n) : name(n), t0(0.0), on(false)
{
  tag[0] = 0;
  if (!rt().profile || !rt().ready) return;
  on = (prof_depth++ == 0);          // nested entries (one C-ABI call using another) are charged to the outer one
  if (on) { cudaStreamSynchronize(rt().stream); t0 = now_seconds(); }
}
ProfScope::~ProfScope()
{
  if (!rt().profile || !rt().ready) return;
  prof_depth--;
  if (!on) return;
  cudaStreamSynchronize(rt().stream);
  ProfEntry& e = prof_table()[std::string(name) + tag];
  e.seconds += now_seconds() - t0; e.calls++;
}

// Size-keyed cache of device blocks.  The Krylov solvers allocate and release their work vectors on every call
// (like the reference's allocate_vector / deallocate_vector); cudaMalloc / cudaFree would serialise the device each
// time, so released blocks are parked here and handed back to the next request of the same size.  All work is
// issued on one stream, so reuse is ordered after the last kernel that touched the block.
struct BlockCache
{
  std::map<size_t, std::vector<void*> > idle;
  std::unordered_map<void*, size_t> size_of;
  size_t idle_bytes = 0;
};
static BlockCache& cache() { static BlockCache c; return c; }

static int cache_trim()
{
  BlockCache& c = cache();
  cudaStreamSynchronize(rt().stream);
  for (auto& kv : c.idle)
    for (void* p : kv.second) { c.size_of.erase(p); cudaFree(p); }
  c.idle.clear();
  c.idle_bytes = 0;
  return 0;
}

int fail(const char* what, cudaError_t e, const char* file, int line)
{
  char buf[512];
  snprintf(buf, sizeof(buf), "[QMG-ERROR]: CUDA failure %s (%s) at %s:%d", cudaGetErrorString(e), what, file, line);
  rt().error = buf;
  fprintf(stderr, "%s\n", buf);
  return 1;
}

int fail_msg(const char* msg)
{
  rt().error = std::string("[QMG-ERROR]: ") + msg;
  fprintf(stderr, "%s\n", rt().error.c_str());
  return 2;
}

int allreduce_result(double* d_buf, int count, int op_max);   // qmg_comm.cu

// A kernel that waits for a peer (in-kernel all-reduce, halo rows stored by the neighbours) gives up after the watchdog
// time and records who was missing in mapped host memory; the host reports it as an ordinary error code.
int peer_timeout_error()
{
  const unsigned long long e = *rt().h_err;
  char buf[160];
  if (e == 1000) snprintf(buf, sizeof(buf), "a pipeline barrier of the ring stencil kernel did not complete (bulk copy lost?)");
  else if (e > 256) snprintf(buf, sizeof(buf), "rank %d gave up waiting for a halo row from its %s neighbour (peer lost?)", qmg_comm_rank(), (e - 257) ? "upper" : "lower");
  else snprintf(buf, sizeof(buf), "rank %d gave up waiting for rank %d in an all-reduce (peer lost?)", qmg_comm_rank(), (int)(e - 1));
  return fail_msg(buf);
}

int fetch_result(double* host_out, int count, int op_max)
{
  Runtime& r = rt();
  r.red_seq++;
  if (r.publish_now)
  {
    // the last block of the reducing kernel wrote the (all-reduced) values and then the sequence number into mapped
    // pinned memory: poll for it instead of a device-to-host copy and a stream synchronisation
    volatile unsigned long long* flag = r.h_flag;
    unsigned long spins = 0;
    while (*flag < r.red_seq)
    {
      if ((++spins & 0xfff) == 0)
      {
        cudaError_t e = cudaStreamQuery(r.stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) return fail("reduction kernel", e, __FILE__, __LINE__);
        if (e == cudaSuccess && *flag < r.red_seq) return fail_msg("reduction finished without publishing its result");
      }
    }
    __sync_synchronize();
    if (*r.h_err != 0) return peer_timeout_error();
    for (int i = 0; i < count; i++) host_out[i] = r.h_result[i];
    return 0;
  }
  int rc = allreduce_result(r.d_result, count, op_max);
  if (rc) return rc;
  QMG_CUDA(cudaMemcpyAsync(r.h_result, r.d_result, sizeof(double) * count, cudaMemcpyDeviceToHost, r.stream));
  QMG_CUDA(cudaStreamSynchronize(r.stream));
  for (int i = 0; i < count; i++) host_out[i] = r.h_result[i];
  return 0;
}

// Ranks can be far apart in HOST time when they meet in a collective (one still generating its gauge field, say), so the
// in-kernel waits are generous; a peer that died still ends the wait with an error instead of a hang.
long long p2p_watchdog_cycles()
{
  double seconds = 120.0;
  const char* e = getenv("QMG_P2P_TIMEOUT_S");
  if (e != nullptr && atof(e) > 0.0) seconds = atof(e);
  return (long long)(seconds * 2.0e9);
}

int skip_result(double* result_dev, int count)
{
  rt().red_seq++;
  return allreduce_result(result_dev, count, 0);     // no-op unless sharded without peer mailboxes
}

// (re)write the device-resident reduction state: called at init and whenever the communicator changes
int upload_red_state(int nranks, int rank, int p2p, double* const* mail)
{
  Runtime& r = rt();
  QMG_CUDA(cudaStreamSynchronize(r.stream));
  RedState st;
  memset(&st, 0, sizeof(st));
  st.seq = r.red_seq;
  // with several ranks the kernels only hold final values when they all-reduce through the peer mailboxes themselves
  r.publish_now = (r.publish && (nranks == 1 || p2p)) ? 1 : 0;
  st.nranks = nranks; st.rank = rank; st.p2p = p2p; st.publish = r.publish_now;
  for (int i = 0; i < kMaxRanks; i++) st.mail[i] = (mail != nullptr && i < nranks) ? mail[i] : nullptr;
  st.host_out = r.h_result; st.host_flag = r.h_flag; st.host_err = r.h_err;
  st.coll_seq = 0;     // a new communicator starts with fresh (zeroed) mailboxes on every rank
  st.watchdog_cycles = p2p_watchdog_cycles();
  QMG_CUDA(cudaMemcpy(r.d_counter, &st, sizeof(st), cudaMemcpyHostToDevice));
  return 0;
}

double* ensure_partials(size_t ndoubles)
{
  Runtime& r = rt();
  if (ndoubles <= r.partials_cap) return r.d_partials;
  cudaStreamSynchronize(r.stream);
  if (r.d_partials) cudaFree(r.d_partials);
  r.d_partials = nullptr; r.partials_cap = 0;
  if (cudaMalloc(&r.d_partials, sizeof(double) * ndoubles) != cudaSuccess) { fail_msg("out of device memory for reduction scratch"); return nullptr; }
  r.partials_cap = ndoubles;
  return r.d_partials;
}

int reduction_grid(long n_items, int items_per_block)
{
  long want = (n_items + items_per_block - 1) / items_per_block;
  long cap = (long)rt().sm_count * 8;
  if (cap > kMaxRedBlocks) cap = kMaxRedBlocks;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

} // namespace qmg

using namespace qmg;

extern "C" {

int qmg_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int qmg_init(int device)
{
  Runtime& r = rt();
  if (r.ready && (device < 0 || device == r.device)) return 0;
  if (r.ready) qmg_finalize();
  int n = qmg_device_count();
  if (n <= 0) return fail_msg("no CUDA device visible: libqmg_b200 has no CPU fallback");
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  QMG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  QMG_CUDA(cudaGetDeviceProperties(&prop, device));
  r.device = device;
  r.sm_count = prop.multiProcessorCount;
  if (ensure_partials((size_t)kMaxRedBlocks * kMaxRedWidth) == nullptr) return 1;
  QMG_CUDA(cudaMalloc(&r.d_counter, sizeof(RedState)));
  QMG_CUDA(cudaMalloc(&r.d_result, sizeof(double) * 2 * kMaxPtrs));
  QMG_CUDA(cudaHostAlloc(&r.h_result, sizeof(double) * 2 * kMaxPtrs + 2 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
  r.h_flag = reinterpret_cast<unsigned long long*>(r.h_result + 2 * kMaxPtrs);
  r.h_err = r.h_flag + 1;
  *r.h_flag = 0; *r.h_err = 0; r.red_seq = 0;
  {
    const char* pe = getenv("QMG_PUBLISH");
    r.publish = (pe != nullptr && pe[0] == '0') ? 0 : 1;
  }
  if (upload_red_state(1, 0, 0, nullptr)) return 1;
  QMG_CUDA(cudaMalloc(&r.d_ptrs, sizeof(void*) * kMaxPtrs));
  QMG_CUDA(cudaMalloc(&r.d_scalars, sizeof(double) * 2 * kMaxPtrs));
  QMG_CUDA(cudaDeviceSynchronize());
  const char* env = getenv("QMG_MANAGED");
  if (env != nullptr && env[0] == '1') r.managed = 1;
  env = getenv("QMG_TILE");
  r.tile_kernel = (env != nullptr && env[0] >= '0' && env[0] <= '9') ? atoi(env) : 1;
  { const char* eb = getenv("QMG_BICGSTAB_FUSED"); r.bicgstab_fused = (eb != nullptr && eb[0] == '0') ? 0 : 1; }
  env = getenv("QMG_PROFILE");
  if (env != nullptr && env[0] == '1') r.profile = 1;
  r.ready = true;
  env = getenv("QMG_LOOPBACK");   // single-GPU exercise of the sharded code path (see qmg_comm_set_loopback)
  if (env != nullptr && env[0] == '1') return qmg_comm_set_loopback(1);
  return 0;
}

int qmg_finalize(void)
{
  Runtime& r = rt();
  if (!r.ready) return 0;
  qmg_comm_finalize();
  cache_trim();
  cudaFree(r.d_partials); cudaFree(r.d_counter); cudaFree(r.d_result); cudaFree(r.d_ptrs); cudaFree(r.d_scalars);
  cudaFreeHost(r.h_result);
  r.d_partials = nullptr; r.partials_cap = 0; r.d_counter = nullptr; r.d_result = nullptr; r.h_result = nullptr; r.d_ptrs = nullptr; r.d_scalars = nullptr;
  r.ready = false;
  return 0;
}

int qmg_set_stream(void* cuda_stream) { QMG_REQUIRE_INIT(); rt().stream = (cudaStream_t)cuda_stream; return 0; }
void* qmg_get_stream(void) { return (void*)rt().stream; }
int qmg_sync(void) { QMG_REQUIRE_INIT(); QMG_CUDA(cudaStreamSynchronize(rt().stream)); if (*rt().h_err != 0) return peer_timeout_error(); return 0; }
const char* qmg_last_error(void) { return rt().error.c_str(); }
int qmg_sm_count(void) { return rt().sm_count; }
long qmg_kernel_launches(void) { return rt().launches; }

int qmg_malloc(void** dptr, size_t bytes)
{
  QMG_REQUIRE_INIT();
  if (bytes == 0) bytes = 16;
  bytes = (bytes + 255) & ~(size_t)255;
  BlockCache& c = cache();
  auto it = c.idle.find(bytes);
  if (it != c.idle.end() && !it->second.empty())
  {
    *dptr = it->second.back();
    it->second.pop_back();
    c.idle_bytes -= bytes;
    return 0;
  }
  cudaError_t e;
  for (int attempt = 0; attempt < 2; attempt++)
  {
    if (rt().managed) e = cudaMallocManaged(dptr, bytes, cudaMemAttachGlobal);
    else e = cudaMalloc(dptr, bytes);
    if (e == cudaSuccess) break;
    cudaGetLastError();
    if (attempt == 0) cache_trim();   // give parked blocks back to the driver and try once more
  }
  if (e != cudaSuccess) return qmg::fail("device allocation", e, __FILE__, __LINE__);
  if (rt().managed)
  {
    cudaMemAdvise(*dptr, bytes, cudaMemAdviseSetPreferredLocation, rt().device);
    cudaGetLastError();   // advice is best effort
  }
  c.size_of[*dptr] = bytes;
  return 0;
}
int qmg_free(void* dptr)
{
  if (dptr == nullptr) return 0;
  BlockCache& c = cache();
  auto it = c.size_of.find(dptr);
  if (it == c.size_of.end()) { QMG_CUDA(cudaFree(dptr)); return 0; }   // not ours (allocated before a finalize)
  c.idle[it->second].push_back(dptr);
  c.idle_bytes += it->second;
  return 0;
}
int qmg_trim(void) { return cache_trim(); }
size_t qmg_cached_bytes(void) { return cache().idle_bytes; }
int qmg_memcpy_h2d(void* dst, const void* src, size_t bytes)
{
  QMG_REQUIRE_INIT();
  QMG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, rt().stream));
  QMG_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}
int qmg_memcpy_d2h(void* dst, const void* src, size_t bytes)
{
  QMG_REQUIRE_INIT();
  QMG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, rt().stream));
  QMG_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}
int qmg_memcpy_d2d(void* dst, const void* src, size_t bytes)
{
  QMG_REQUIRE_INIT();
  QMG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, rt().stream));
  return 0;
}
int qmg_zero_bytes(void* dptr, size_t bytes)
{
  QMG_REQUIRE_INIT();
  if (bytes == 0) return 0;
  QMG_CUDA(cudaMemsetAsync(dptr, 0, bytes, rt().stream));
  return 0;
}
int qmg_set_tile_kernel(int mode) { QMG_REQUIRE_INIT(); if (mode < 0 || mode > 15) return fail_msg("qmg_set_tile_kernel: mode must be 0 .. 15"); rt().tile_kernel = mode; return 0; }
int qmg_get_tile_kernel(void) { return rt().tile_kernel; }
int qmg_set_bicgstab_fused(int on) { QMG_REQUIRE_INIT(); rt().bicgstab_fused = on ? 1 : 0; return 0; }
int qmg_get_bicgstab_fused(void) { return rt().bicgstab_fused; }
int qmg_profile_enable(int on) { rt().profile = on ? 1 : 0; return 0; }
int qmg_profile_reset(void) { prof_table().clear(); return 0; }
// prints "name calls seconds" rows sorted by time to stdout and returns the total seconds
double qmg_profile_report(void)
{
  std::vector<std::pair<double, std::string> > rows;
  double total = 0.0;
  for (auto& kv : prof_table()) { rows.push_back(std::make_pair(kv.second.seconds, kv.first)); total += kv.second.seconds; }
  std::sort(rows.begin(), rows.end());
  printf("[QMG-PROFILE] %-40s %10s %12s %7s\n", "entry point", "calls", "seconds", "share");
  for (size_t i = rows.size(); i-- > 0;)
    printf("[QMG-PROFILE] %-40s %10ld %12.6f %6.1f%%\n", rows[i].second.c_str(), prof_table()[rows[i].second].calls, rows[i].first, 100.0 * rows[i].first / (total > 0 ? total : 1.0));
  printf("[QMG-PROFILE] %-40s %10s %12.6f\n", "total", "", total);
  fflush(stdout);
  return total;
}
int qmg_set_alloc_mode(int managed)
{
  QMG_REQUIRE_INIT();
  const int want = managed ? 1 : 0;
  // parked blocks are keyed by size only: hand them back to the driver so that the next qmg_malloc really is of the new kind
  if (want != rt().managed) { int rc = cache_trim(); if (rc) return rc; }
  rt().managed = want;
  return 0;
}
int qmg_get_alloc_mode(void) { return rt().managed; }
// NUMA node the active GPU hangs off (sysfs), or -1 when the platform does not say
int qmg_device_numa_node(void)
{
  if (!rt().ready) return -1;
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), rt().device) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char* c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (f == nullptr) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

// Pinned host staging on the GPU's own NUMA node: with one process per GPU, every rank's host vectors otherwise land on
// whichever node the launcher started it on (node 0 for all eight ranks of a torchrun job), and eight PCIe streams then
// cross the socket interconnect into one memory controller.  The calling thread is moved onto the GPU's node for the
// allocation and the first touch (default "local" policy), then gets its affinity back.
int qmg_malloc_host(void** hptr, size_t bytes)
{
  QMG_REQUIRE_INIT();
  if (bytes == 0) bytes = 16;
  cpu_set_t old_set, node_set;
  bool moved = false;
  const int node = qmg_device_numa_node();
  const char* off = getenv("QMG_NUMA");
  if (node >= 0 && !(off != nullptr && off[0] == '0') && sched_getaffinity(0, sizeof(old_set), &old_set) == 0)
  {
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (f != nullptr)
    {
      CPU_ZERO(&node_set);
      int a, b, any = 0;
      // "0-31,64-95": ranges separated by commas
      while (fscanf(f, "%d", &a) == 1)
      {
        b = a;
        int ch = fgetc(f);
        if (ch == '-') { if (fscanf(f, "%d", &b) != 1) b = a; ch = fgetc(f); }
        for (int cpu = a; cpu <= b && cpu < CPU_SETSIZE; cpu++) if (CPU_ISSET(cpu, &old_set)) { CPU_SET(cpu, &node_set); any = 1; }   // never leave the cpuset we were given
        if (ch != ',') break;
      }
      fclose(f);
      if (any && sched_setaffinity(0, sizeof(node_set), &node_set) == 0) moved = true;
    }
  }
  cudaError_t e = cudaHostAlloc(hptr, bytes, cudaHostAllocPortable);
  if (e == cudaSuccess) memset(*hptr, 0, bytes);        // first touch from the GPU's node
  if (moved) sched_setaffinity(0, sizeof(old_set), &old_set);
  if (e != cudaSuccess) return qmg::fail("pinned host allocation", e, __FILE__, __LINE__);
  return 0;
}
int qmg_free_host(void* hptr) { if (hptr) { QMG_CUDA(cudaFreeHost(hptr)); } return 0; }

} // extern "C"
