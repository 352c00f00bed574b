"""Regenerates the fixtures in this directory.  Run in the build container only
(needs /root/reference and oracle/_ref/libqmg_ref.so):

    python tests/golden/make_golden.py

1. l{L}t{L}b60_phases.npy: the reference's thermalised U(1) phase configs
   (/root/reference/tests/common_cfgs_u1/*.dat, text, x outer / y / mu inner) as float64 arrays.
2. golden_outputs.npz: outputs of the reference's own code (oracle/_ref = unmodified reference
   headers) on seeded inputs, so `-m gpu` runs on a box without /root/reference and without a
   rebuilt oracle can still check against the reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import capi  # noqa: E402
import latutil  # noqa: E402

REF_CFG = "/root/reference/tests/common_cfgs_u1"


def main():
    for L in (32, 64, 128, 256):
        ph = np.loadtxt(os.path.join(REF_CFG, "l%dt%db60_heatbath.dat" % (L, L)), dtype=np.float64)
        assert ph.size == 2 * L * L
        np.save(os.path.join(HERE, "l%dt%db60_phases.npy" % (L, L)), ph)

    be = capi.Backend("ref")
    out = {}
    meta = {}
    # n11: Wilson 64^2, beta 6.0, mass -0.055 (tests/n11_wilson_test/wilson_test.cpp:43), apply to a point source and a gaussian
    L = 64
    lat = be.lattice(L, L, 2)
    gauge = latutil.load_gauge(L)
    w = lat.wilson(-0.055, gauge)
    rhs = latutil.gaussian_cv(lat.size_cv, 11)
    out["n11_wilson64_gauss_rhs_seed"] = np.array([11])
    out["n11_wilson64_gauss_out"] = w.apply(rhs, 0)
    pt = np.zeros(lat.size_cv, np.complex128)
    pt[int(latutil.site_index(32, 32, L, L)) * 2] = 1.0
    out["n11_wilson64_point_out"] = w.apply(pt, 0)
    w.build(dagger=True, rbjacobi=True, rbj_dagger=True)
    for t, name in ((1, "dagger"), (2, "rbjacobi"), (6, "rbj_dagger"), (5, "mdagm")):
        out["n11_wilson64_%s_out" % name] = w.apply(rhs, t)
    x, info = w.solve(3, pt, type=0, max_iter=4000, tol=1e-8, iparam=16)
    meta["n11_gcr16_point"] = info
    w.free()
    # staggered 32^2 and laplace 32^2
    L = 32
    lat1 = be.lattice(L, L, 1)
    gauge = latutil.load_gauge(L)
    rhs1 = latutil.gaussian_cv(lat1.size_cv, 4)
    st = lat1.staggered(0.1, gauge)
    out["n04_stag32_out"] = st.apply(rhs1, 0)
    st.free()
    lp = lat1.laplace(0.01, gauge)
    out["n03_laplace32_out"] = lp.apply(rhs1, 0)
    lp.free()
    np.savez_compressed(os.path.join(HERE, "golden_outputs.npz"), **out)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", sorted(out), meta)


if __name__ == "__main__":
    main()
