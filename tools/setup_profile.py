"""Per-entry-point profile of the K-cycle SET-UP (null vectors, block orthonormalisation, Galerkin builds) at one size.
  python tools/setup_profile.py 4096"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import capi  # noqa: E402
import latutil  # noqa: E402
import qmg  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
qmg.init(0)
be = capi.Backend("gpu")
if os.environ.get("SETUP_LINK_COMPRESSED", "1") != "0":      # as bench.py builds its hierarchies
    be.fn("kcycle_setup_link_compressed")(1)
g = latutil.synthetic_gauge(L, L, 6.0, 1337, slab=True)
kc = capi.KCycle(be, L, -0.05, g, n_refine=2, inner_iters=100, coarsest_iters=400)      # warm allocator
print("warm set-up seconds", kc.solve(max_iter=1)["setup_seconds"], flush=True)
kc.free()
lib = qmg.lib()
lib.qmg_profile_reset()
lib.qmg_profile_enable(1)
t0 = time.perf_counter()
kc = capi.KCycle(be, L, -0.05, g, n_refine=2, inner_iters=100, coarsest_iters=400)
print("profiled set-up wall seconds", time.perf_counter() - t0, flush=True)
lib.qmg_profile_enable(0)
lib.qmg_profile_report.restype = C.c_double
sys.stdout.flush()
lib.qmg_profile_report()
kc.free()
