"""Wilson apply on L x L: stored blocks against the matrix-free flavour (gauge links), bursts and isolated launches.  python tools/mf_probe.py [L]"""
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
del ph
cl, hp = qmg.fill_wilson(L, L, gauge)
x, y, b = qmg.cvec(n, zero=False), qmg.cvec(n), qmg.cvec(n, zero=False)
for v in (x, b):
    qmg.check(lib.qmg_gaussian(qmg.ptr(v), C.c_long(n), C.c_uint64(7), C.c_uint64(0), C.c_double(1.0)))
stored = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.05)
free = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.05, wilson_gauge=gauge, wilson_w=1.0)
print("deviation", qmg.wilson_mf_deviation(free))
for name, d, nbytes in (("stored blocks", stored, 384.0 * V), ("matrix-free", free, 96.0 * V)):
    for resid in (False, True):
        def fn():
            if resid:
                qmg.check(lib.qmg_stencil_apply_residual(C.byref(d), C.c_int(15), C.c_int(15), qmg.ptr(y), qmg.ptr(x), qmg.ptr(b)))
            else:
                qmg.stencil_apply(d, y, x)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%-14s %-9s %dx%d: %.4f ms  (%.0f GB/s on its %d B/site; %.0f GB/s counted as the contract's 384 B/site)"
              % (name, "residual" if resid else "apply", L, L, ms, (nbytes + (32.0 * V if resid else 0)) / ms / 1e6, int(nbytes / V), 384.0 * V / ms / 1e6), flush=True)
