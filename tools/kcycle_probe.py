"""Time the native K-cycle driver (kcycle_new / kcycle_solve) on the GPU (or the oracle) at a list of sizes.
  python tools/kcycle_probe.py gpu 256 1024 4096 [--mass -0.05] [--verbosity 0]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import capi  # noqa: E402
import latutil  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("backend")
ap.add_argument("sizes", type=int, nargs="+")
ap.add_argument("--mass", type=float, default=-0.05)
ap.add_argument("--verbosity", type=int, default=0)
ap.add_argument("--n-refine", type=int, default=2)
ap.add_argument("--inner-iters", type=int, default=60)
ap.add_argument("--coarsest-iters", type=int, default=200)
ap.add_argument("--max-iter", type=int, default=60)
ap.add_argument("--null-iters", type=int, default=500)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--restart", type=int, default=32)
ap.add_argument("--seed", type=int, default=1337)
ap.add_argument("--hermitian", action="store_true", help="link-compressed (gamma5-hermitian) applies on every level that qualifies")
ap.add_argument("--hermitian-setup", action="store_true", dest="hermitian_setup", help="link-compressed applies already during the set-up (null-vector solves)")
ap.add_argument("--matrix-free", action="store_true", dest="matrix_free", help="Wilson fine operator applies matrix-free (gauge links instead of stored blocks)")
ap.add_argument("--unfused", action="store_true", help="the reference's sweep-for-sweep K-cycle (round-1 behaviour) instead of the fused one")
args = ap.parse_args()
if args.backend == "gpu":
    import qmg
    qmg.init(0)
be = capi.Backend(args.backend)
if args.backend == "gpu" and args.hermitian_setup:
    be.fn("kcycle_setup_link_compressed")(1)
if args.backend == "gpu" and args.matrix_free:
    be.fn("kcycle_setup_matrix_free")(1)
for L in args.sizes:
    t0 = time.perf_counter()
    g = latutil.synthetic_gauge(L, L, 6.0, args.seed, slab=True)
    t1 = time.perf_counter()
    kc = capi.KCycle(be, L, args.mass, g, n_refine=args.n_refine, inner_iters=args.inner_iters, coarsest_iters=args.coarsest_iters,
                     null_max_iter=args.null_iters, verbosity=args.verbosity)
    t2 = time.perf_counter()
    if args.backend == "gpu":
        if args.hermitian:
            print("link-compressed levels:", kc.gamma5_hermitian(True, tile_levels_only=True), flush=True)
        if args.unfused:
            kc.set_fused(False)
    out = kc.solve(max_iter=args.max_iter, verbosity=args.verbosity, restart=args.restart)
    t3 = time.perf_counter()
    if args.profile and args.backend == "gpu":
        qmg.lib().qmg_profile_reset(); qmg.lib().qmg_profile_enable(1)
    t4 = time.perf_counter()
    out2 = kc.solve(max_iter=args.max_iter, verbosity=0, restart=args.restart)
    if args.profile and args.backend == "gpu":
        qmg.lib().qmg_profile_enable(0)
        qmg.lib().qmg_profile_report.restype = __import__("ctypes").c_double
        sys.stdout.flush(); qmg.lib().qmg_profile_report()
    out["second_solve_s"] = out2["seconds"]; out["second_solve_iter"] = out2["iter"]; out["second_solve_wall_s"] = time.perf_counter() - t4
    out.update(L=L, gauge_s=t1 - t0, setup_wall_s=t2 - t1, solve_wall_s=t3 - t2, per_level=[kc.tracker(l) for l in range(args.n_refine + 1)],
               executed=[kc.executed(l) for l in range(args.n_refine + 1)],
               precond_s=kc.time_precond(1, 2))
    print(json.dumps(out), flush=True)
    kc.free()
