"""Device-time GB/s of the hot kernels through the C ABI (CUDA events on the launching stream).
  python tools/kernel_probe.py [--reps 10]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import qmg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
args = ap.parse_args()
if os.environ.get("QMG_LIB_OVERRIDE"):      # A / B timing against an older build of the kernel library
    qmg.LIB_PATH = os.environ["QMG_LIB_OVERRIDE"]
qmg.init(0)
lib = qmg.lib()


def timed(fn, reps=args.reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def rnd(n):
    t = qmg.cvec(n, zero=False)
    qmg.check(lib.qmg_gaussian(qmg.ptr(t), C.c_long(n), C.c_uint64(1), C.c_uint64(n % 97), C.c_double(1.0)))
    return t


print("%-44s %10s %10s %8s" % ("kernel", "ms", "GB/s", "of 6548"))


def report(name, sec, nbytes):
    print("%-44s %10.4f %10.1f %8.3f" % (name, sec * 1e3, nbytes / sec / 1e9, nbytes / sec / 1e9 / 6547.8), flush=True)


if not args.only or "stencil" in args.only:
    # nc = 12, 24, 48: domain-wall operators with Ls = 6, 12, 24 (the any-nc kernel)
    cases = ((12, 1024), (24, 512), (48, 256)) if os.environ.get("PROBE_GENERIC_ONLY") else None
    for nc, L in cases or ((2, 8192), (2, 4096), (2, 1024), (8, 2048), (8, 1024), (8, 256), (1, 8192), (4, 2048), (16, 512), (12, 1024), (24, 512), (48, 256)):
        V = L * L
        cl, hp = rnd(V * nc * nc), rnd(4 * V * nc * nc)
        x, y = rnd(V * nc), qmg.cvec(V * nc)
        d = qmg.stencil_desc(L, L, nc, cl, hp, shift=0.1)
        sec = timed(lambda: qmg.stencil_apply(d, y, x))
        report("stencil apply nc=%d %dx%d" % (nc, L, L), sec, 16.0 * V * (nc * nc * 5 + 2 * nc))
        if nc in (2, 4, 8, 16, 32):
            dh = qmg.stencil_desc(L, L, nc, cl, hp, shift=0.1, gamma5_hermitian=True)
            sec = timed(lambda: qmg.stencil_apply(dh, y, x))
            report("  link-compressed (3 of 5 blocks) nc=%d %dx%d" % (nc, L, L), sec, 16.0 * V * (nc * nc * 5 + 2 * nc))
        d2 = qmg.stencil_desc(L, L, nc, None, hp)
        sec = timed(lambda: qmg.stencil_apply(d2, y, x, qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD | qmg.APPLY_IDENTITY_CLOVER))
        report("  rbjacobi (identity clover) nc=%d %dx%d" % (nc, L, L), sec, 16.0 * V * (nc * nc * 4 + 2 * nc))
        del cl, hp, x, y
        torch.cuda.empty_cache()

if not args.only or "blas" in args.only:
    for n in (2 * 4096 * 4096, 8 * 1024 * 1024, 8 * 256 * 256):
        x, y, p, q = rnd(n), rnd(n), rnd(n), rnd(n)
        out = (C.c_double * 4)()
        cd = C.c_double
        report("dot_norm n=%d" % n, timed(lambda: lib.qmg_dot_norm(qmg.ptr(x), qmg.ptr(y), C.c_long(n), out)), 32.0 * n)
        report("norm2sq n=%d" % n, timed(lambda: lib.qmg_norm2sq(qmg.ptr(x), C.c_long(n), out)), 16.0 * n)
        report("update_xr_norm n=%d" % n, timed(lambda: lib.qmg_update_xr_norm(cd(0.1), cd(0.2), qmg.ptr(p), qmg.ptr(q), qmg.ptr(x), qmg.ptr(y), C.c_long(n), out)), 96.0 * n)
        report("caxpbyz n=%d" % n, timed(lambda: lib.qmg_caxpbyz(cd(1.0), cd(0.0), qmg.ptr(x), cd(-1.0), cd(0.0), qmg.ptr(y), qmg.ptr(p), C.c_long(n))), 48.0 * n)
        vs = [rnd(n) for _ in range(8)]
        arr = (C.c_void_p * 8)(*[t.data_ptr() for t in vs])
        res = (C.c_double * 16)()
        report("multi_dot k=8 n=%d" % n, timed(lambda: lib.qmg_multi_dot(arr, 8, qmg.ptr(x), C.c_long(n), res)), 16.0 * 9 * n)
        co = (C.c_double * 16)(*([0.01] * 16))
        report("multi_axpyz k=8 n=%d" % n, timed(lambda: lib.qmg_multi_axpyz(co, arr, 8, qmg.ptr(x), qmg.ptr(y), C.c_long(n))), 16.0 * 10 * n)
        del vs, x, y, p, q
        torch.cuda.empty_cache()

if not args.only or "transfer" in args.only:
    for (Lf, ncf, Lc, ncc) in ((4096, 2, 1024, 8), (1024, 8, 256, 8), (256, 8, 64, 8)):
        nf, ncv = Lf * Lf * ncf, Lc * Lc * ncc
        nv = [rnd(nf) for _ in range(ncc)]
        f, c = rnd(nf), rnd(ncv)
        td = qmg.transfer_desc(Lf, Lf, ncf, Lc, Lc, ncc)
        nbytes = 16.0 * (nf * (ncc + 1) + ncv)
        report("restrict %dx%d nc%d -> %dx%d nc%d" % (Lf, Lf, ncf, Lc, Lc, ncc), timed(lambda: qmg.restrict(td, nv, f, c)), nbytes)
        report("prolong  %dx%d nc%d <- %dx%d nc%d" % (Lf, Lf, ncf, Lc, Lc, ncc), timed(lambda: qmg.prolong(td, nv, c, f)), nbytes + 16.0 * nf)
        del nv, f, c
        torch.cuda.empty_cache()
