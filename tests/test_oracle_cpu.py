"""CPU suite, part 1: pin the ORACLE (oracle/_ref = the reference's unmodified headers + the clean-room quantum-linalg
shim) against everything the reference itself offers as a known answer -- its test programs print-only, so these are
the invariants they print plus the committed golden outputs.  No GPU, no product code."""
import json
import os

import numpy as np
import pytest

import capi
import latutil

pytestmark = pytest.mark.skipif(not capi.have_ref(), reason="oracle/_ref/libqmg_ref.so not built (needs /root/reference; run __graft_entry__.build())")


@pytest.fixture(scope="module")
def ref():
    return capi.Backend("ref")


def test_lattice_layout(ref):
    """lattice.h:75-81,199-205 against the numpy restatement used by every fixture."""
    for X, Y in ((6, 4), (8, 8), (2, 2), (16, 4)):
        lat = ref.lattice(X, Y, 3)
        xs, ys = latutil.site_coords(X, Y)
        for i in range(X * Y):
            assert lat.index_to_coord(i) == (xs[i], ys[i])
            assert lat.coord_to_index(int(xs[i]), int(ys[i])) == i
        assert (lat.volume, lat.size_cv, lat.size_cm, lat.size_gauge, lat.size_hopping) == (X * Y, 3 * X * Y, 9 * X * Y, 18 * X * Y, 36 * X * Y)


def test_n00_cshift_known_answer(ref):
    """tests/n00_cshift/cshift_2d_test.cpp: values = site number; FROM_XP1 gives out(x,y) = in(x+1,y) etc., periodic."""
    X, Y = 6, 4
    xs, ys = latutil.site_coords(X, Y)
    for nc in (1, 2):
        lat = ref.lattice(X, Y, nc)
        v = np.repeat((ys * X + xs).astype(np.complex128), nc)
        for cdir, (dx, dy) in ((2, (1, 0)), (3, (0, 1)), (4, (-1, 0)), (5, (0, -1))):
            want = np.repeat((((ys + dy) % Y) * X + (xs + dx) % X).astype(np.complex128), nc)
            assert np.array_equal(lat.cshift(v, cdir, 3, nc), want)


def test_free_field_stencils(ref):
    """Unit gauge field: Laplace = 4 + m^2 on site, -1 to each neighbour (tests/n02 intent, gaugedlaplace.h:45-68);
    constant vectors are eigenvectors of the free Laplace (m^2) and of the free Wilson operator (mass)."""
    L = 8
    unit = np.ones(2 * L * L, np.complex128)
    lat = ref.lattice(L, L, 1)
    op = lat.laplace(0.3, unit)
    assert np.allclose(op.get("clover"), 4.0) and np.allclose(op.get("hopping"), -1.0)
    assert np.allclose(op.apply(np.ones(L * L, np.complex128), 0), 0.3)
    delta = np.zeros(L * L, np.complex128)
    delta[latutil.site_index(3, 3, L, L)] = 1.0
    out = op.apply(delta, 0)
    assert np.isclose(out[latutil.site_index(3, 3, L, L)], 4.3)
    for x, y in ((4, 3), (2, 3), (3, 4), (3, 2)):
        assert np.isclose(out[latutil.site_index(x, y, L, L)], -1.0)
    assert np.isclose(np.abs(out).sum(), 8.3)
    lw = ref.lattice(L, L, 2)
    w = lw.wilson(0.25, unit)
    const = np.ones(2 * L * L, np.complex128)
    assert np.allclose(w.apply(const, 0), 0.25 * const)     # 2w - 4 (w/2) + (spin parts cancel) + mass
    op.free(); w.free()


def test_golden_outputs(ref):
    """The committed fixtures (tests/golden/make_golden.py) are what the freshly built oracle still produces."""
    gold = np.load(os.path.join(latutil.GOLDEN, "golden_outputs.npz"))
    L = 64
    lat = ref.lattice(L, L, 2)
    w = lat.wilson(-0.055, latutil.load_gauge(L))
    rhs = latutil.gaussian_cv(lat.size_cv, int(gold["n11_wilson64_gauss_rhs_seed"][0]))
    assert np.array_equal(w.apply(rhs, 0), gold["n11_wilson64_gauss_out"])
    w.build(dagger=True, rbjacobi=True, rbj_dagger=True)
    for t, name in ((1, "dagger"), (2, "rbjacobi"), (6, "rbj_dagger"), (5, "mdagm")):
        assert latutil.rel_l2(w.apply(rhs, t), gold["n11_wilson64_%s_out" % name]) < 1e-14
    pt = np.zeros(lat.size_cv, np.complex128)
    pt[int(latutil.site_index(32, 32, L, L)) * 2] = 1.0
    x, info = w.solve(3, pt, type=0, max_iter=4000, tol=1e-8, iparam=16)
    meta = json.load(open(os.path.join(latutil.GOLDEN, "golden_meta.json")))
    assert info["iter"] == meta["n11_gcr16_point"]["iter"] and info["success"]
    # n11's own check: explicit residual with the operator (wilson_test.cpp:199-204)
    assert np.linalg.norm(w.apply(x, 0) - pt) < 1.5e-8
    w.free()


def test_dagger_is_adjoint(ref):
    """tests/n17, n21: <y|M x> = <M^dag y|x> for the stored dagger and rbj-dagger stencils."""
    L = 16
    g = latutil.phases_to_gauge(np.random.default_rng(1).normal(0, 0.5, size=2 * L * L), L, L)
    w = ref.lattice(L, L, 2).wilson(0.1 + 0.05j, g)
    w.build(dagger=True, rbjacobi=True, rbj_dagger=True)
    x, y = latutil.gaussian_cv(2 * L * L, 1), latutil.gaussian_cv(2 * L * L, 2)
    assert abs(np.vdot(y, w.apply(x, 0)) - np.vdot(w.apply(y, 1), x)) < 1e-11
    assert abs(np.vdot(y, w.apply(x, 2)) - np.vdot(w.apply(y, 6), x)) < 1e-11
    # right block Jacobi: M B^-1 with B^-1 = rbjacobi_cinv applied by reconstruct
    z = w.reconstruct(x, x, 2)
    assert latutil.rel_l2(w.apply(z, 0), w.apply(x, 2)) < 1e-12
    w.free()


@pytest.mark.parametrize("case", [(4, 4, 2, 1, 1, 6), (16, 16, 2, 4, 4, 8), (8, 8, 1, 4, 4, 2)])
def test_n05_n06_transfer_identities(ref, case):
    """tests/n05_prolong_restrict_test (:85-103): (1 - P P^dag) v_i = 0 on the null vectors, (1 - P^dag P) v_c = 0;
    tests/n06_transfer_decomp: Sigma^dag Sigma = block Gram matrix."""
    Xf, Yf, ncf, Xc, Yc, ncc = case
    fl, cl = ref.lattice(Xf, Yf, ncf), ref.lattice(Xc, Yc, ncc)
    nv = np.stack([latutil.gaussian_cv(fl.size_cv, 60 + v) for v in range(ncc)])
    tr = capi.Transfer(fl, cl, nv, block_ortho=True, save_decomp=True)
    for v in nv:
        assert latutil.rel_l2(tr.prolong(tr.restrict(v)), v) < 1e-12
    cv = latutil.gaussian_cv(cl.size_cv, 3)
    assert latutil.rel_l2(tr.restrict(tr.prolong(cv)), cv) < 1e-12
    if Xc * Yc == 1:
        chol = tr.cholesky().reshape(ncc, ncc)
        assert np.allclose(chol.conj().T @ chol, nv.conj() @ nv.T, rtol=1e-10, atol=1e-10)
        assert np.allclose(np.tril(chol, -1), 0)
    tr.free()


def test_n08_coarse_equals_RAP(ref):
    """tests/n08_distance1_build_test/build_test.cpp:117-155: explicit coarse stencil == R A P over successive coarsenings."""
    L = 16
    g = latutil.phases_to_gauge(np.random.default_rng(2).normal(0, 0.5, size=2 * L * L), L, L)
    fl = ref.lattice(L, L, 1)
    op = fl.laplace(0.1, g)
    cur, ncf = L, 1
    for step in range(3):
        cur //= 2
        cl = ref.lattice(cur, cur, 2)
        nv = np.stack([latutil.gaussian_cv(fl.size_cv, 10 * step + v) for v in range(2)])
        tr = capi.Transfer(fl, cl, nv)
        co = tr.coarse_operator(op, is_chiral=False)
        x = latutil.gaussian_cv(cl.size_cv, 99)
        assert latutil.rel_l2(co.apply(x, 0), tr.restrict(op.apply(tr.prolong(x), 0))) < 1e-12
        fl, op = cl, co


def test_n13_kcycle_converges(ref):
    """tests/n13_wilson_kcycle semantics on l64t64b60: the K-cycle preconditioned VPGCR(32) reaches 1e-10 and the
    explicit residual (:464-471) agrees; far fewer outer iterations than plain GCR(32)."""
    L = 64
    g = latutil.load_gauge(L)
    kc = capi.KCycle(ref, L, -0.075, g, n_refine=1)
    x, out = kc.solve(b=latutil.gaussian_cv(2 * L * L, 13), want_x=True)
    assert out["success"] and out["iter"] < 40 and out["check_relres"] < 2e-10
    t0, t1 = kc.tracker(0), kc.tracker(1)
    assert t0["presmooth"] == 5 * out["iter"] and t0["postsmooth"] == 4 * out["iter"]   # MR(2): 4 ops + residual
    assert t1["iters"] > 0 and t1["krylov"] >= t1["iters"]
    kc.free()


def test_synthetic_gauge_matches_heatbath_plaquette():
    """The large-lattice synthetic field has the plaquette of the reference's thermalised beta = 6 configs."""
    L = 64
    want = np.exp(-1.0 / 12.0)
    assert abs(latutil.average_plaquette(latutil.load_gauge(L), L, L) - want) < 0.01
    assert abs(latutil.average_plaquette(latutil.synthetic_gauge(L, L, 6.0, 5), L, L) - want) < 0.01


def test_u1_known_answers(ref, tmp_path):
    """tests/n01_u1_test: a unit field has plaquette 1 and topology 0; the plaquette is gauge invariant; a written
    field reads back; the shipped thermalised beta = 6 configuration sits at <plaq> = exp(-1/12); a charge-1 instanton
    on a smooth field shifts the topology by one; the non-compact action of a pure gauge field vanishes."""
    L = 16
    lat = ref.lattice(L, L, 1)
    unit = lat.u1_create(0)
    pl, q = lat.u1_observables(unit)
    assert abs(pl - 1.0) < 1e-15 and abs(q) < 1e-15
    hot = lat.u1_create(1, seed=3)
    tr = lat.u1_create(3, seed=4)
    pl0, q0 = lat.u1_observables(hot)
    pl1, q1 = lat.u1_observables(lat.u1_gauge_trans(hot, tr))
    assert abs(pl0 - pl1) < 1e-13 and abs(q0 - q1) < 1e-9
    path = str(tmp_path / "cfg16_hot.dat")
    lat.u1_file(1, path, gauge=hot)
    back = lat.u1_file(0, path)
    assert np.max(np.abs(back - hot)) < 1e-14
    smooth = lat.u1_create(2, beta=60.0, seed=5)
    q_before = lat.u1_observables(smooth)[1]
    q_after = lat.u1_observables(lat.u1_instanton(smooth, 1.0, L // 2, L // 2))[1]
    assert abs((q_after - q_before) - 1.0) < 0.05 or abs((q_after - q_before) + 1.0) < 0.05
    g64 = latutil.load_gauge(64)
    pl64, _ = ref.lattice(64, 64, 1).u1_observables(g64)
    assert abs(pl64.real - np.exp(-1.0 / 12.0)) < 0.01
    # pure gauge phases A_mu(x) = a(x + mu) - a(x): every plaquette angle is zero
    a = np.random.default_rng(2).normal(size=(L, L))
    xs, ys = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    idx = latutil.site_index(xs, ys, L, L)
    ph = np.zeros(2 * L * L)
    ph[idx.ravel()] = (np.roll(a, -1, axis=0) - a).ravel()
    ph[L * L + idx.ravel()] = (np.roll(a, -1, axis=1) - a).ravel()
    assert lat.u1_action(ph, 6.0) < 1e-24
