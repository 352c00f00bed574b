"""TEST INFRASTRUCTURE: why bench.py and the 128^2 / 256^2 parity tests run the n13 K-cycle at mass -0.05.

n13's usage string suggests -0.075 (tests/n13_wilson_kcycle/wilson_kcycle.cpp:40,81: "eigenvalues go negative around
-0.075").  On the reference's own thermalised configs l128t128b60 and l256t256b60 that mass is beyond critical: the CPU
reference (oracle/_ref, unmodified headers) does not converge in 100 outer iterations.  This script runs the oracle at
both masses and prints what it finds; its output is committed as profiles/r03_oracle_mass_m0075_nonconvergence.log.

  python tools/oracle_mass_probe.py [L ...]        (default: 128)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import capi      # noqa: E402
import latutil   # noqa: E402

be = capi.Backend("ref")
for L in [int(a) for a in sys.argv[1:]] or [128]:
    g = latutil.load_gauge(L)
    for mass in (-0.05, -0.075):
        t0 = time.time()
        kc = capi.KCycle(be, L, mass, g, n_refine=1, inner_iters=100, coarsest_iters=400)
        out = kc.solve(tol=1e-10, restart=32, max_iter=100)
        print("oracle n13 K-cycle  l%dt%db60  mass %+.3f : success %s  outer iterations %d  |r| %.3e  explicit relres %.3e  "
              "setup %.1f s  solve %.1f s  (wall %.1f s)" % (L, L, mass, out["success"], out["iter"], out["resSq"] ** 0.5,
                                                          out["check_relres"], out["setup_seconds"], out["seconds"], time.time() - t0), flush=True)
        kc.free()
