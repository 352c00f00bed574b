"""Tile-kernel timing at nc = 8 for the patch shape selected by QMG_TILE.  python tools/tile_probe.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import qmg  # noqa: E402

if os.environ.get("QMG_LIB_OVERRIDE"):      # A / B timing against an older build of the kernel library
    qmg.LIB_PATH = os.environ["QMG_LIB_OVERRIDE"]
qmg.init(0)
lib = qmg.lib()


def rnd(n):
    t = qmg.cvec(n, zero=False)
    qmg.check(lib.qmg_gaussian(qmg.ptr(t), C.c_long(n), C.c_uint64(1), C.c_uint64(n % 97), C.c_double(1.0)))
    return t


for nc, L in ((8, 2048), (8, 512)) if not os.environ.get('TILE_PROBE_SMALL') else ((8, 2048),):
    V = L * L
    cl, hp = rnd(V * nc * nc), rnd(4 * V * nc * nc)
    x, y = rnd(V * nc), qmg.cvec(V * nc)
    for herm in (False, True):
        d = qmg.stencil_desc(L, L, nc, cl, hp, shift=0.1, gamma5_hermitian=herm)
        for _ in range(3):
            qmg.stencil_apply(d, y, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            qmg.stencil_apply(d, y, x)
        e1.record()
        torch.cuda.synchronize()
        print("QMG_TILE=%s nc=%d %dx%d herm=%d: %.4f ms" % (os.environ.get("QMG_TILE", "1"), nc, L, L, herm, e0.elapsed_time(e1) / 20), flush=True)
        if os.environ.get("TILE_PROBE_ISOLATED"):
            # one launch at a time with the GPU idle in between (what ncu times) against the burst above (what a solver sees)
            import time
            ts = []
            for _ in range(8):
                time.sleep(0.05)
                e0.record(); qmg.stencil_apply(d, y, x); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print("   isolated launches (50 ms idle before each): " + " ".join("%.4f" % v for v in ts), flush=True)
            e0.record()
            for _ in range(400):
                qmg.stencil_apply(d, y, x)
            e1.record(); torch.cuda.synchronize()
            print("   burst of 400: %.4f ms per apply" % (e0.elapsed_time(e1) / 400), flush=True)
    del cl, hp, x, y
    torch.cuda.empty_cache()
