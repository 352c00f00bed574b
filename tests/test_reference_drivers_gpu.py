"""The reference's OWN test drivers, unmodified, on the B200.

oracle/Makefile compiles tests/n11_wilson_test/wilson_test.cpp and tests/n13_wilson_kcycle/wilson_kcycle.cpp from where they lie
under /root/reference twice: against the reference headers + the quantum-linalg shim (the oracle, CPU) and against the product
(include/qmg + libqmg_b200.so).  Both binaries travel under oracle/_ref/drivers/.  Here each pair runs on the same inputs (the
shipped U(1) configurations, re-written in the reference's text format from the committed fixtures; both seed
std::mt19937(1337)) and what they PRINT is compared: iteration counts and explicit residuals.

n11 indexes a vector from host code (wilson_test.cpp:170), so the product run uses QMG_MANAGED=1."""
import os
import re
import subprocess

import numpy as np
import pytest

import latutil

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "oracle", "_ref", "drivers")


@pytest.fixture(scope="module")
def rundir(tmp_path_factory):
    for n in ("n11_ref", "n11_b200", "n13_ref", "n13_b200"):
        if not os.path.exists(os.path.join(DRV, n)):
            pytest.skip("oracle/_ref/drivers not built (make -C oracle in the build container)")
    base = tmp_path_factory.mktemp("refdrivers")
    cfg = base / "common_cfgs_u1"
    cfg.mkdir()
    for L in (32, 64):
        ph = np.load(os.path.join(latutil.GOLDEN, "l%dt%db60_phases.npy" % (L, L)))
        np.savetxt(str(cfg / ("l%dt%db60_heatbath.dat" % (L, L))), ph, fmt="%.20f")
    run = base / "run"
    run.mkdir()
    return str(run)


def run_driver(name, rundir, args=(), managed=False):
    env = dict(os.environ)
    env.pop("QMG_LOOPBACK", None)
    if managed:
        env["QMG_MANAGED"] = "1"
    r = subprocess.run([os.path.join(DRV, name)] + list(args), cwd=rundir, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def test_n13_wilson_kcycle_driver(qmg_gpu, rundir):
    """./wilson_kcycle 64 -0.05 6.0 2: 3-level K-cycle on l64t64b60 (wilson_kcycle.cpp:459-471 prints the outer iteration
    count and the explicit check)."""
    outs = {}
    for name in ("n13_ref", "n13_b200"):
        text = run_driver(name, rundir, ["64", "-0.05", "6.0", "2"])
        m = re.search(r"Multigrid converged in (\d+) iterations with alleged tolerance ([0-9.]+e[+-]?[0-9]+)", text)
        c = re.search(r"Check tolerance ([0-9.eE+-]+)", text)
        assert m and c, text[-2000:]
        outs[name] = (int(m.group(1)), float(m.group(2)), float(c.group(1)))
    (ir, tr, cr), (ig, tg, cg) = outs["n13_ref"], outs["n13_b200"]
    assert abs(ir - ig) <= 1, outs
    assert cr < 1e-9 and cg < 1e-9 and abs(tg - cg) < 1e-12


def test_n11_wilson_test_driver(qmg_gpu, rundir):
    """The solver survey on l32t32b60 (point source): every solver line the two builds print carries the same iteration count
    (+-1; +-2 for the long BiCGstab runs) and an explicit residual below tolerance.  TFQMR exists only in the product (the
    oracle's shim declares it), so it is held to its own explicit residual."""
    pat = re.compile(r"\[QMG-TEST-([A-Z0-9()-]+)\]: (Potential Error! )?Algorithm (.+?) took (\d+) iterations to reach a tolerance of ([0-9.eE+-]+)")
    err = re.compile(r"\[QMG-TEST-([A-Z0-9()-]+)\]?: The relative error is ([0-9.eE+-]+)")
    outs = {}
    for name, managed in (("n11_ref", False), ("n11_b200", True)):
        text = run_driver(name, rundir, managed=managed)
        outs[name] = ({m.group(1): (int(m.group(4)), m.group(2) is None) for m in pat.finditer(text)},
                      {m.group(1): float(m.group(2)) for m in err.finditer(text)})
    (itr, er), (itg, eg) = outs["n11_ref"], outs["n11_b200"]
    assert set(itg) == set(itr) and len(itg) >= 7, (itr, itg)
    for key, (n_g, ok_g) in itg.items():
        assert ok_g and eg[key] < 1e-9, (key, n_g, eg[key])
        if key == "TFQMR":
            continue
        n_r, ok_r = itr[key]
        assert ok_r and abs(n_r - n_g) <= (2 if "BICGSTAB" in key else 1), (key, n_r, n_g)
