"""Set-up and solve times of BASELINE configs 3 (staggered 1024^2, 3 levels) and 4 (n22 adaptive set-up + 3-level Wilson K-cycle
on 4096^2) on the GPU, as tests/test_host_gpu.py runs them at full size.  python tools/baseline_configs_probe.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import capi  # noqa: E402
import latutil  # noqa: E402
import qmg  # noqa: E402

qmg.init(0)
gpu = capi.Backend("gpu")


def run(name, make, solve_kw):
    for rep in range(2):            # second pass: warm allocator
        t0 = time.perf_counter()
        kc = make()
        t1 = time.perf_counter()
        out = kc.solve(**solve_kw)
        out2 = kc.solve(**solve_kw)
        ops = [kc.tracker(l)["total"] for l in range(3)]
        ex = [kc.executed(l) for l in range(3)]
        kc.free()
        qmg.check(qmg.lib().qmg_trim())
        print(json.dumps({"config": name, "pass": rep, "setup_wall_s": round(t1 - t0, 3), "setup_seconds": out.get("setup_seconds"), "iter": out["iter"],
                          "solve_seconds_cold": out["seconds"], "solve_seconds_warm": out2["seconds"], "check_relres": out["check_relres"],
                          "per_level_ops": ops, "per_level_ops_executed": ex}), flush=True)


g = latutil.synthetic_gauge(1024, 1024, 6.0, 21)
run("staggered 1024^2, mass 0.1, 3 levels (1024 -> 256 -> 64), 8 dof", lambda: capi.KCycle(gpu, 1024, 0.1, g, n_refine=2, block=4, coarse_dof=8, seed=3, staggered=True, inner_iters=100, coarsest_iters=400),
    dict(tol=1e-10, max_iter=600))
g = latutil.synthetic_gauge(4096, 4096, 6.0, 1337, slab=True)
run("n22: Wilson 4096^2, mass -0.05, one adaptive set-up round, 3 levels", lambda: capi.KCycle(gpu, 4096, -0.05, g, n_refine=2, block=4, coarse_dof=8, seed=3, adaptive_setups=1, inner_iters=100, coarsest_iters=400),
    dict(tol=1e-10, restart=16, max_iter=100))
