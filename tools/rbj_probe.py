"""Plain apply against the right-block-Jacobi flavour (identity clover: 4 of 5 blocks) at one size.  python tools/rbj_probe.py [nc L]"""
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
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
reps = int(os.environ.get("REPS", "20"))


def rnd(n):
    t = qmg.cvec(n, zero=False)
    qmg.check(lib.qmg_gaussian(qmg.ptr(t), C.c_long(n), C.c_uint64(1), C.c_uint64(n % 97), C.c_double(1.0)))
    return t


V = L * L
cl, hp = rnd(V * nc * nc), rnd(4 * V * nc * nc)
x, y = rnd(V * nc), qmg.cvec(V * nc)
d = qmg.stencil_desc(L, L, nc, cl, hp, shift=0.1)
d2 = qmg.stencil_desc(L, L, nc, None, hp)
HOP = qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD
cases = (("plain (5 blocks)", lambda: qmg.stencil_apply(d, y, x), 5),
         ("identity clover + hopping (4 blocks)", lambda: qmg.stencil_apply(d2, y, x, HOP | qmg.APPLY_IDENTITY_CLOVER), 4),
         ("hopping only (4 blocks)", lambda: qmg.stencil_apply(d2, y, x, HOP), 4),
         ("clover + shift only (1 block)", lambda: qmg.stencil_apply(d, y, x, qmg.APPLY_CLOVER | qmg.APPLY_SHIFT), 1))
for name, fn, nb in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 16.0 * V * (nc * nc * nb + 2 * nc)
    print("nc=%d %dx%d %-40s %.4f ms  %.0f GB/s" % (nc, L, L, name, ms, nbytes / ms / 1e6), flush=True)
