"""e2e (host rhs -> device -> apply -> host lhs through qmg_stencil_apply_host) against the chunk size of its pipeline.
  python tools/e2e_probe.py [L]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import qmg  # noqa: E402

qmg.init(0)
lib = qmg.lib()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
V, n = L * L, L * L * 2
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
ph = torch.randn(2 * V, generator=gen, device="cuda", dtype=torch.float64) / 6.0 ** 0.5
gauge = torch.polar(torch.ones_like(ph), ph)
cl, hp = qmg.fill_wilson(L, L, gauge)
del ph, gauge
rhs, lhs = qmg.cvec(n), qmg.cvec(n)
qmg.check(lib.qmg_gaussian(qmg.ptr(rhs), C.c_long(n), C.c_uint64(7), C.c_uint64(0), C.c_double(1.0)))
desc = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.075)
hin, hout = C.c_void_p(), C.c_void_p()
qmg.check(lib.qmg_malloc_host(C.byref(hin), C.c_size_t(16 * n)))
qmg.check(lib.qmg_malloc_host(C.byref(hout), C.c_size_t(16 * n)))
qmg.check(lib.qmg_memcpy_d2h(hin, qmg.ptr(rhs), C.c_size_t(16 * n)))
for rows in (0, 32, 64, 128, 256, 512, 1024, 2048, 0):
    qmg.stencil_apply_host(desc, hout, hin, dev_lhs=lhs, dev_rhs=rhs, rows_per_chunk=rows)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        qmg.stencil_apply_host(desc, hout, hin, dev_lhs=lhs, dev_rhs=rhs, rows_per_chunk=rows)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print("rows_per_chunk %5d (%.0f MB): %.2f ms per apply, %.0f GB/s" % (rows, rows * L * 2 * 16 / 1e6, ms, 384.0 * V / ms / 1e6), flush=True)
