"""Latency of the small-level building blocks of the coarsest GCR, one process per GPU (torchrun) or a single GPU:
a reduction (kernel + in-kernel all-reduce + host poll), a Krylov step (2 kernels, 2 all-reduces, 1 host wait), a
device-side GCR orthogonalisation (k = 8) and an nc = 8 apply on this rank's slab of a 512 x 512 lattice.
  python -m torch.distributed.run --nproc-per-node 8 tools/latency_probe.py      |      python tools/latency_probe.py"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import qmg  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
qmg.init(local)
if world > 1:
    qmg.comm_init()
lib = qmg.lib()
X, Y, nc = 512, 512 // world, 8
V = X * Y
n = V * nc


def rnd(m, s):
    t = qmg.cvec(m, zero=False)
    qmg.check(lib.qmg_gaussian(qmg.ptr(t), C.c_long(m), C.c_uint64(s), C.c_uint64(rank), C.c_double(1.0)))
    return t


def wall(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


x, y, p, q, r = (rnd(n, s) for s in range(5))
out = (C.c_double * 8)()
res = {}
res["norm2sq (1 reduction, host wait)"] = wall(lambda: lib.qmg_norm2sq(qmg.ptr(x), C.c_long(n), out))
res["krylov_step (2 kernels, 1 host wait)"] = wall(lambda: lib.qmg_krylov_step(C.c_double(1.0), qmg.ptr(p), qmg.ptr(q), qmg.ptr(x), qmg.ptr(x), qmg.ptr(r), qmg.ptr(r), None, C.c_long(n), 0, out, None))
k = 8
Ap, P = [rnd(n, 10 + j) for j in range(k)], [rnd(n, 30 + j) for j in range(k)]
arrA, arrP = (C.c_void_p * k)(*[t.data_ptr() for t in Ap]), (C.c_void_p * k)(*[t.data_ptr() for t in P])
dots = torch.zeros(2 * k, dtype=torch.float64, device="cuda")
apn = torch.ones(k + 1, dtype=torch.float64, device="cuda")
res["gcr_orthogonalize k=8 (no host wait)"] = wall(lambda: lib.qmg_gcr_orthogonalize(arrA, arrP, k, qmg.ptr(y), qmg.ptr(q), qmg.ptr(p), qmg.ptr(r), C.c_long(n), qmg.ptr(dots), qmg.ptr(apn)))
for kk in (1, 2, 4, 8, 12, 16, 24):
    ApK, PK = [rnd(n, 10 + j) for j in range(kk)], [rnd(n, 30 + j) for j in range(kk)]
    aA, aP = (C.c_void_p * kk)(*[t.data_ptr() for t in ApK]), (C.c_void_p * kk)(*[t.data_ptr() for t in PK])
    dk = torch.zeros(2 * kk, dtype=torch.float64, device="cuda")
    ak = torch.ones(kk + 1, dtype=torch.float64, device="cuda")
    us = wall(lambda: lib.qmg_gcr_orthogonalize(aA, aP, kk, qmg.ptr(y), qmg.ptr(q), qmg.ptr(p), qmg.ptr(r), C.c_long(n), qmg.ptr(dk), qmg.ptr(ak)), reps=100)
    gb = 16.0 * n * ((kk + 1) + (2 * kk + 3) + 2) / 1e9
    res["gcr_orthogonalize k=%d: %.0f MB -> %.0f GB/s" % (kk, gb * 1e3, gb / (us * 1e-6))] = us
    del ApK, PK
res["caxpy (1 kernel, no wait)"] = wall(lambda: lib.qmg_caxpy(C.c_double(0.0), C.c_double(0.0), qmg.ptr(x), qmg.ptr(y), C.c_long(n)))
cl, hp = rnd(V * nc * nc, 50), rnd(4 * V * nc * nc, 51)
d = qmg.stencil_desc(X, Y, nc, cl, hp, shift=0.1)
res["stencil apply nc=8 stored blocks"] = wall(lambda: qmg.stencil_apply(d, y, x))
res["  + norm2sq after it (apply, wait)"] = wall(lambda: (qmg.stencil_apply(d, y, x), lib.qmg_norm2sq(qmg.ptr(y), C.c_long(n), out)))
if rank == 0:
    print("ranks %d, slab %d x %d nc %d (%d elements per rank)" % (world, X, Y, nc, n))
    for kx, v in res.items():
        print("  %-44s %8.1f us" % (kx, v))
if world > 1:
    qmg.comm_finalize()
    dist.destroy_process_group()
