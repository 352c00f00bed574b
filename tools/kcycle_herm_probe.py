"""K-cycle solve time with stored-block applies vs link-compressed (gamma5-hermitian) applies.  python tools/kcycle_herm_probe.py 4096"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import capi  # noqa: E402
import latutil  # noqa: E402
import qmg  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
qmg.init(0)
be = capi.Backend("gpu")
g = latutil.synthetic_gauge(L, L, 6.0, 1337, slab=True)
b = latutil.gaussian_cv(L * L * 2, 3)
kc = capi.KCycle(be, L, -0.05, g, n_refine=2, inner_iters=100, coarsest_iters=400)
del g
kc.solve(b, tol=1e-10, restart=8, max_iter=100)
for mode in ("stored", "compressed", "stored", "compressed"):
    n = kc.gamma5_hermitian(mode == "compressed")
    out = kc.solve(b, tol=1e-10, restart=8, max_iter=100)
    print(json.dumps(dict(mode=mode, levels_switched=n, iter=out["iter"], seconds=out["seconds"], relres=out["check_relres"],
                          ops=[kc.tracker(l)["total"] for l in range(3)], precond_s=kc.time_precond(1, 2))), flush=True)
kc.free()
