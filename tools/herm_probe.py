"""Stored-block vs link-compressed apply, a few launches each (for ncu).  python tools/herm_probe.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import qmg  # noqa: E402

qmg.init(0)
lib = qmg.lib()


def rnd(n):
    t = qmg.cvec(n, zero=False)
    qmg.check(lib.qmg_gaussian(qmg.ptr(t), C.c_long(n), C.c_uint64(1), C.c_uint64(n % 97), C.c_double(1.0)))
    return t


for nc, L in ((8, 2048), (2, 8192)):
    V = L * L
    cl, hp = rnd(V * nc * nc), rnd(4 * V * nc * nc)
    x, y = rnd(V * nc), qmg.cvec(V * nc)
    for herm in (False, True):
        d = qmg.stencil_desc(L, L, nc, cl, hp, shift=0.1, gamma5_hermitian=herm)
        for _ in range(3):
            qmg.stencil_apply(d, y, x)
    torch.cuda.synchronize()
    del cl, hp, x, y
    torch.cuda.empty_cache()
print("ok")
