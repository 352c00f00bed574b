"""Parity of the B200 host classes (include/qmg: the reference's class API over device memory) against the
oracle (the reference's unmodified headers, oracle/_ref).  The SAME driver text (quantum-mg_b200/host/qmg_capi_body.h)
is compiled against both, so each test is "same calls, two back ends".
Tolerances: 1e-12 relative L2 for stencil applies (BASELINE.json north_star), 1e-10/1e-11 where an nc x nc inverse
or a Gram-Schmidt sits in between; solver iteration counts +-1."""
import numpy as np
import pytest

import capi
import latutil

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ref():
    if not capi.have_ref():
        pytest.skip("oracle/_ref/libqmg_ref.so not built")
    return capi.Backend("ref")


@pytest.fixture(scope="module")
def gpu(qmg_gpu):
    return capi.Backend("gpu")


def both(ref, gpu, fn):
    return fn(ref), fn(gpu)


def test_cshift_n00(ref, gpu):
    """tests/n00_cshift: site-number lattice, every direction and source parity, nc = 1, 2."""
    X, Y = 6, 4
    xs, ys = latutil.site_coords(X, Y)
    for nc in (1, 2):
        v = np.repeat((ys * X + xs).astype(np.complex128), nc)
        for cdir in (2, 3, 4, 5):
            for eo in (1, 2, 3):
                a, b = both(ref, gpu, lambda be: be.lattice(X, Y, nc).cshift(v, cdir, eo, nc, lhs=-np.ones_like(v)))
                assert np.array_equal(a, b), (nc, cdir, eo)


OPS = {
    "wilson": lambda lat, g: lat.wilson(-0.055, g),
    "staggered": lambda lat, g: lat.staggered(0.1, g),
    "laplace": lambda lat, g: lat.laplace(0.01, g),
    "dwf4": lambda lat, g: lat.dwf(0.05, g, 4, -1.0),
}
NC = {"wilson": 2, "staggered": 1, "laplace": 1, "dwf4": 8}


@pytest.mark.parametrize("kind", ["wilson", "staggered", "laplace", "dwf4"])
def test_operator_fill_apply_variants(ref, gpu, kind):
    """Operator construction from U(1) links, all nine QMGStencilType applies (stencil_2d.h:2418), every accumulate piece,
    prepare_M / reconstruct_M round trip."""
    L = 32
    g = latutil.load_gauge(L)
    lr, lg = ref.lattice(L, L, NC[kind]), gpu.lattice(L, L, NC[kind])
    a, b = OPS[kind](lr, g), OPS[kind](lg, g)
    for name in ("clover", "hopping"):
        x, y = a.get(name), b.get(name)
        assert (x is None) == (y is None)
        if x is not None:
            assert latutil.rel_l2(y, x) < TOL
    assert a.shifts() == b.shifts()
    rhs = latutil.gaussian_cv(lr.size_cv, 1)
    acc = latutil.gaussian_cv(lr.size_cv, 2)
    for piece in (0, 1, 2, 3, 4, 6, 9, 10):
        assert latutil.rel_l2(b.apply_piece(piece, rhs, lhs=acc), a.apply_piece(piece, rhs, lhs=acc)) < TOL, piece
    for piece in (5, 7, 8):
        for mu in range(4):
            assert latutil.rel_l2(b.apply_piece(piece, rhs, dir=mu, lhs=acc), a.apply_piece(piece, rhs, dir=mu, lhs=acc)) < TOL, (piece, mu)
    a.build(dagger=True, rbjacobi=(kind != "staggered" or True), rbj_dagger=True)
    b.build(dagger=True, rbjacobi=True, rbj_dagger=True)
    assert a.built() == b.built()
    for name in capi.Stencil.NAMES[2:]:
        x, y = a.get(name), b.get(name)
        assert (x is None) == (y is None), name
        if x is not None:
            assert latutil.rel_l2(y, x) < 1e-11, name
    for t in range(9):
        want, got = a.apply(rhs, t), b.apply(rhs, t)
        assert latutil.rel_l2(got, want) < 1e-11, t
        # accumulate flavour apply_M(lhs, rhs, type)
        assert latutil.rel_l2(b.apply_piece(12, rhs, dir=t, lhs=acc), a.apply_piece(12, rhs, dir=t, lhs=acc)) < 1e-11, t
        pa, pb = a.prepare(rhs, t), b.prepare(rhs, t)
        assert latutil.rel_l2(pb, pa) < 1e-11, t
        ra, rb = a.reconstruct(rhs, acc, t), b.reconstruct(rhs, acc, t)
        assert latutil.rel_l2(rb, ra) < 1e-11, t
    assert latutil.rel_l2(b.apply_piece(11, rhs, lhs=acc), a.apply_piece(11, rhs, lhs=acc)) < 1e-11
    a.free(); b.free()


@pytest.mark.parametrize("kind", ["wilson", "staggered", "dwf4"])
def test_chirality(ref, gpu, kind):
    """gamma5 / sigma1 / chiral projections / apply_sigma (wilson.h:74-148, staggered.h:140-186, dwf.h:104-147)."""
    L = 16
    g = latutil.phases_to_gauge(np.random.default_rng(3).normal(0, 0.4, size=L * L * 2), L, L)
    a, b = OPS[kind](ref.lattice(L, L, NC[kind]), g), OPS[kind](gpu.lattice(L, L, NC[kind]), g)
    v, w = latutil.gaussian_cv(L * L * NC[kind], 1), latutil.gaussian_cv(L * L * NC[kind], 2)
    for op in range(0, 13):
        if op in (13,):
            continue
        ra, rb = a.chiral(op, v, w), b.chiral(op, v, w)
        for x, y in zip(ra, rb):
            if x is not None:
                assert np.allclose(y, x, rtol=0, atol=1e-14), (kind, op)
    a.free(); b.free()


def test_n18_noise_on_clover(ref, gpu):
    """tests/n18_rbjacobi_stencil_test/rbjacobi_stencil_test.cpp:133-231: noise is added to the PUBLIC clover pointer,
    then GCR runs on the original, right-block-Jacobi and Schur systems and the solution is reconstructed."""
    L = 32
    g = latutil.load_gauge(L)
    noise = 0.1 * latutil.gaussian_cv(L * L * 4, 18)
    b_src = latutil.gaussian_cv(L * L * 2, 19)
    res = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        op = be.lattice(L, L, 2).wilson(0.1, g)
        op.add_to("clover", noise)
        op.build(rbjacobi=True)
        out = {}
        for t, n in ((0, None), (2, None), (3, L * L)):
            bp = op.prepare(b_src, t)
            y, info = op.solve(2, bp, type=t, n=n, max_iter=500, tol=1e-9)
            x = op.reconstruct(y, b_src, t)
            check = np.linalg.norm(op.apply(x, 0) - b_src) / np.linalg.norm(b_src)
            out[t] = (x, info, check)
        res[name] = out
        op.free()
    for t in (0, 2, 3):
        xr, ir, cr = res["ref"][t]
        xg, ig, cg = res["gpu"][t]
        assert ir["success"] and ig["success"]
        assert abs(ir["iter"] - ig["iter"]) <= 1, (t, ir, ig)
        assert cg < 5e-9 and cr < 5e-9
        assert latutil.rel_l2(xg, xr) < 1e-7


@pytest.mark.parametrize("solver,kind,type,kw", [
    (0, "laplace", 0, dict(tol=1e-8)),                               # n03: CG on the Laplace operator
    (2, "wilson", 0, dict(tol=1e-8, max_iter=400)),                  # n11: GCR
    (3, "wilson", 0, dict(tol=1e-8, iparam=16, max_iter=3000)),      # n11: GCR(16)
    (4, "wilson", 0, dict(tol=1e-3, dparam=0.85, max_iter=200)),     # MR(omega) as used by the smoother
    (5, "wilson", 0, dict(tol=5e-5, iparam=6, max_iter=500)),        # n13: BiCGstab-6 null-vector solve
    (6, "wilson", 0, dict(tol=1e-10, dparam=0.33, iparam=250, max_iter=10)),  # n22: Richardson relaxation
    (0, "wilson", 5, dict(tol=1e-8, max_iter=4000)),                 # n17: CGNR on M^dag M
    (1, "wilson", 4, dict(tol=1e-8, iparam=64, max_iter=4000)),      # n17: restarted CGNE on M M^dag
])
def test_solvers_iteration_parity(ref, gpu, solver, kind, type, kw):
    L = 32
    g = latutil.load_gauge(L)
    b_src = np.zeros(L * L * NC[kind], np.complex128)
    b_src[latutil.site_index(L // 2, L // 2, L, L) * NC[kind]] = 1.0
    x0 = latutil.gaussian_cv(b_src.size, 5)
    out = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        op = OPS[kind](be.lattice(L, L, NC[kind]), g) if kind != "wilson" else be.lattice(L, L, 2).wilson(0.05, g)
        op.build(dagger=True)
        bp = op.prepare(b_src, type)
        out[name] = op.solve(solver, bp, type=type, x0=x0, **kw)
        op.free()
    (xr, ir), (xg, ig) = out["ref"], out["gpu"]
    assert ir["success"] == ig["success"]
    assert abs(ir["iter"] - ig["iter"]) <= 1, (ir, ig)
    assert abs(ir["ops"] - ig["ops"]) <= 2
    assert latutil.rel_l2(xg, xr) < 1e-6
    assert abs(np.sqrt(ig["resSq"]) - np.sqrt(ir["resSq"])) <= 1e-3 * np.sqrt(ir["resSq"]) + 1e-14


@pytest.mark.parametrize("Lb", [1, 2, 4, 6, 7, 8])
def test_bicgstab_l_fused_sweep_matches_call_by_call(gpu, Lb):
    """BiCGstab(L) through the fused kernels (qmg_bicgstab_replay / _mgs / _finish: 148 instead of 280 vector passes per L = 6
    sweep) against the call-by-call sequence of the same solver: same iteration and operator counts, solutions equal to
    rounding (the replay and finish passes repeat the element arithmetic exactly; the Gram-Schmidt sums are added in another
    order).  L = 8 exceeds qmg_bicgstab_max_l() and must take the call-by-call path on its own."""
    import ctypes as C
    import qmg
    lib = qmg.lib()
    L = 32
    g = latutil.load_gauge(L)
    b_src = latutil.gaussian_cv(L * L * 2, 31)
    x0 = latutil.gaussian_cv(b_src.size, 5)
    out = {}
    was = lib.qmg_get_bicgstab_fused()
    try:
        for fused in (1, 0):
            qmg.check(lib.qmg_set_bicgstab_fused(C.c_int(fused)))
            op = gpu.lattice(L, L, 2).wilson(0.05, g)
            out[fused] = op.solve(5, b_src, type=0, x0=x0, tol=5e-7, iparam=Lb, max_iter=600)
            op.free()
    finally:
        qmg.check(lib.qmg_set_bicgstab_fused(C.c_int(was)))
    (xf, inf_f), (xc, inf_c) = out[1], out[0]
    assert inf_f["success"] and inf_c["success"]
    assert inf_f["iter"] == inf_c["iter"] and inf_f["ops"] == inf_c["ops"], (inf_f, inf_c)
    assert latutil.rel_l2(xf, xc) < 1e-9
    if Lb > lib.qmg_bicgstab_max_l():
        assert np.array_equal(xf, xc)


TRANSFER_CASES = [(16, 16, 2, 4, 4, 8), (8, 8, 8, 2, 2, 8), (8, 8, 1, 4, 4, 2), (4, 4, 2, 1, 1, 6)]


@pytest.mark.parametrize("case", TRANSFER_CASES)
def test_transfer_class(ref, gpu, case):
    """TransferMG (transfer.h:118-179): block-orthonormalised copies, P, R, Cholesky factor; n05 / n06 identities."""
    Xf, Yf, ncf, Xc, Yc, ncc = case
    nv = np.stack([latutil.gaussian_cv(Xf * Yf * ncf, 70 + v) for v in range(ncc)])
    cv, fv = latutil.gaussian_cv(Xc * Yc * ncc, 1), latutil.gaussian_cv(Xf * Yf * ncf, 2)
    out = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        fl, cl = be.lattice(Xf, Yf, ncf), be.lattice(Xc, Yc, ncc)
        tr = capi.Transfer(fl, cl, nv, block_ortho=True, save_decomp=True, doubling=1)
        out[name] = (tr.nullvecs(), tr.cholesky(), tr.prolong(cv, fv), tr.restrict(fv, cv), tr.props())
        tr.free()
    for i in range(4):
        assert latutil.rel_l2(out["gpu"][i], out["ref"][i]) < 1e-11, i
    assert out["gpu"][4] == out["ref"][4]
    # Sigma^dag Sigma = block Gram matrix of the ORIGINAL vectors (tests/n06_transfer_decomp)
    chol = out["gpu"][1].reshape(Xc * Yc, ncc, ncc)
    if Xc * Yc == 1:
        gram = nv.conj() @ nv.T
        assert np.allclose(chol[0].conj().T @ chol[0], gram, rtol=1e-10, atol=1e-10)


def test_transfer_asymmetric(ref, gpu):
    """Separate restrict vectors with block bi-orthonormalisation and saved L, U (transfer.h:185-225, 610-769; n05 :121-139)."""
    Xf, Yf, ncf, Xc, Yc, ncc = 8, 8, 2, 2, 2, 4
    pv = np.stack([latutil.gaussian_cv(Xf * Yf * ncf, 80 + v) for v in range(ncc)])
    rv = pv + 0.3 * np.stack([latutil.gaussian_cv(Xf * Yf * ncf, 90 + v) for v in range(ncc)])
    cv = latutil.gaussian_cv(Xc * Yc * ncc, 1)
    out = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        fl, cl = be.lattice(Xf, Yf, ncf), be.lattice(Xc, Yc, ncc)
        tr = capi.Transfer(fl, cl, pv, block_ortho=True, save_decomp=True, restrict_vecs=rv)
        L_, U_ = tr.LU()
        out[name] = (tr.nullvecs(0), tr.nullvecs(1), L_, U_, tr.restrict(tr.prolong(cv)), tr.props())
        tr.free()
    for i in range(4):
        assert latutil.rel_l2(out["gpu"][i], out["ref"][i]) < 1e-10, i
    assert latutil.rel_l2(out["gpu"][4], cv) < 1e-11      # (1 - R P) v_c = 0
    assert out["gpu"][5] == out["ref"][5]


@pytest.mark.parametrize("use_rbj,extra", [(False, 0), (False, 5), (True, 2)])
def test_coarse_operator_class(ref, gpu, use_rbj, extra):
    """CoarseOperator2D built through the class API (coarse.h:90-471), incl. building from the rbjacobi fine stencil
    and the extra dagger / rbjacobi / rbj-dagger sets; then R A P == A_c on a random vector (tests/n08)."""
    L, Lc = 16, 4
    g = latutil.phases_to_gauge(np.random.default_rng(8).normal(0, 0.4, size=L * L * 2), L, L)
    nv = np.stack([latutil.gaussian_cv(L * L * 2, 30 + v) for v in range(8)])
    x = latutil.gaussian_cv(Lc * Lc * 8, 3)
    out = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        fl, cl = be.lattice(L, L, 2), be.lattice(Lc, Lc, 8)
        op = fl.wilson(0.02, g)
        if use_rbj:
            op.build(rbjacobi=True)
        tr = capi.Transfer(fl, cl, nv, doubling=1)
        co = tr.coarse_operator(op, is_chiral=True, use_rbjacobi=use_rbj, build_extra=extra)
        arrays = [co.get(n) for n in capi.Stencil.NAMES]
        rap = tr.restrict(op.apply(tr.prolong(x), 2 if use_rbj else 0))
        out[name] = (arrays, co.apply(x, 0), rap, co.shifts(), co.built(), co.chiral(8, x, x))
        co.free(); tr.free(); op.free()
    for n, xa, ya in zip(capi.Stencil.NAMES, out["ref"][0], out["gpu"][0]):
        assert (xa is None) == (ya is None), n
        if xa is not None:
            assert np.linalg.norm(ya - xa) < 1e-10 * max(1.0, np.linalg.norm(xa)), n
    assert latutil.rel_l2(out["gpu"][1], out["ref"][1]) < 1e-11
    # Galerkin identity: the explicit coarse stencil (+ inherited shift) equals R A P
    assert latutil.rel_l2(out["gpu"][1], out["gpu"][2]) < 1e-11
    assert out["gpu"][3] == out["ref"][3] and out["gpu"][4] == out["ref"][4]
    for a_, b_ in zip(out["ref"][5], out["gpu"][5]):
        assert np.allclose(a_, b_, atol=1e-14)


@pytest.mark.parametrize("asym", [False, True])
def test_coarse_apply_sigma(ref, gpu, asym):
    """CoarseOperator2D::apply_sigma with the four QMGSigmaTypeCoarse flavours (coarse.h:19-25, 661-894): sigma_1 seen
    through the Cholesky factor (R = P^dag) or the L, U factors (R != P^dag) the transfer saved; the RBJ flavours also use
    B and B^-dag of the coarse operator.  Without saved factors the call is an error that leaves the output alone."""
    L, Lc, ncc = 16, 4, 8
    g = latutil.phases_to_gauge(np.random.default_rng(8).normal(0, 0.4, size=L * L * 2), L, L)
    pv = np.stack([latutil.gaussian_cv(L * L * 2, 30 + v) for v in range(ncc)])
    rv = pv + 0.3 * np.stack([latutil.gaussian_cv(L * L * 2, 60 + v) for v in range(ncc)]) if asym else None
    x = latutil.gaussian_cv(Lc * Lc * ncc, 3)
    out = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        fl, cl = be.lattice(L, L, 2), be.lattice(Lc, Lc, ncc)
        op = fl.wilson(0.02, g)
        tr = capi.Transfer(fl, cl, pv, block_ortho=True, save_decomp=True, doubling=2, restrict_vecs=rv)
        co = tr.coarse_operator(op, is_chiral=True, use_rbjacobi=False, build_extra=5)
        res = [co.coarse_sigma(t, x) for t in (6, 7, 8, 9)]
        tr0 = capi.Transfer(fl, cl, pv, block_ortho=True, save_decomp=False, doubling=2)
        co0 = tr0.coarse_operator(op, is_chiral=True)
        res.append(co0.coarse_sigma(6, x))
        out[name] = res
        co.free(); co0.free(); tr.free(); tr0.free(); op.free()
    for i, (a_, b_) in enumerate(zip(out["ref"], out["gpu"])):
        assert latutil.rel_l2(b_, a_) < 1e-10, i
    assert not np.any(out["gpu"][4])
    if not asym:
        assert np.array_equal(out["gpu"][0], out["gpu"][1])      # sigma_1^L == sigma_1^R when R = P^dag


def _n13_nullvecs(be, lat, op, coarse_dof, seed):
    """Null vectors as tests/n13_wilson_kcycle/wilson_kcycle.cpp:338-385 generates them (through backend `be`)."""
    rng = np.random.default_rng(seed)
    n = lat.size_cv
    nv = np.zeros((coarse_dof, n), np.complex128)
    for j in range(coarse_dof // 2):
        eta = rng.normal(size=n) + 1j * rng.normal(size=n)
        for k in range(j):
            eta -= np.vdot(nv[k], eta) / np.vdot(nv[k], nv[k]) * nv[k]
        e, _ = op.solve(5, -op.apply(eta, 0), max_iter=500, tol=5e-5, iparam=6)
        nv[j] = e + eta
        for k in range(j):
            nv[j] -= np.vdot(nv[k], nv[j]) / np.vdot(nv[k], nv[k]) * nv[k]
    for j in range(coarse_dof // 2):
        up, down = op.chiral(8, nv[j], nv[j])
        nv[j], nv[j + coarse_dof // 2] = up / np.linalg.norm(up), down / np.linalg.norm(down)
    return nv


def _build_mg(be, L, g, mass, nvs, level_app=0, coarsest_app=0, build_from=0, build_extra=0, block=4, coarse_dof=8):
    lat0 = be.lattice(L, L, 2)
    op = lat0.wilson(mass, g)
    if level_app != 0 or coarsest_app != 0:
        op.build(rbjacobi=True)
    mg = capi.Multigrid(lat0, op, coarsest_type=coarsest_app, coarsest_tol=0.2, coarsest_iters=1000, coarsest_restart=32)
    lats, cur = [lat0], L
    for nv in nvs:
        cur //= block
        lc = be.lattice(cur, cur, coarse_dof)
        tr = capi.Transfer(lats[-1], lc, nv, block_ortho=True, save_decomp=False, doubling=1)
        mg.push_level(lc, tr, fine_stencil_app=level_app, inner_tol=0.2, inner_iters=1000, inner_restart=32, pre_iters=2, post_iters=2,
                      build_stencil=True, is_chiral=True, build_from=build_from, build_extra=build_extra, nvecs=nv)
        lats.append(lc)
    return mg, op, lats


@pytest.mark.parametrize("L,levels", [(64, 2), (64, 3)])
def test_n13_kcycle_parity(ref, gpu, L, levels):
    """tests/n13_wilson_kcycle: identical raw null vectors fed to both back ends; outer VPGCR(32) to 1e-10.
    Outer iteration count +-1, per-level operator counts close, same solution (north_star parity gates)."""
    g = latutil.load_gauge(L)
    mass = -0.075
    # raw null vectors level by level from the ORACLE, then reused verbatim on the GPU
    nvs = []
    mg_r, op_r, lats_r = _build_mg(ref, L, g, mass, [])
    for lev in range(levels - 1):
        st = mg_r.stencil(lev)
        nv = _n13_nullvecs(ref, lats_r[lev], st, 8, seed=100 + lev)
        nvs.append(nv)
        mg_r.free()
        mg_r, op_r, lats_r = _build_mg(ref, L, g, mass, nvs)
    mg_g, op_g, lats_g = _build_mg(gpu, L, g, mass, nvs)
    assert mg_g.num_levels() == mg_r.num_levels() == levels
    # coarse operators agree level by level
    for lev in range(1, levels):
        for name in ("clover", "hopping"):
            xa, ya = mg_r.stencil(lev).get(name), mg_g.stencil(lev).get(name)
            assert np.linalg.norm(ya - xa) < 1e-10 * np.linalg.norm(xa), (lev, name)
    b = latutil.gaussian_cv(L * L * 2, 13)
    # one preconditioner application
    zr, zg = mg_r.precond(b), mg_g.precond(b)
    assert latutil.rel_l2(zg, zr) < 1e-6
    mg_r.reset_tracker(); mg_g.reset_tracker()
    xr, ir = mg_r.solve(b, tol=1e-10, restart=32)
    xg, ig = mg_g.solve(b, tol=1e-10, restart=32)
    assert ir["success"] and ig["success"]
    assert abs(ir["iter"] - ig["iter"]) <= 1, (ir, ig)
    assert np.sqrt(ig["resSq"]) / np.linalg.norm(b) < 1e-10
    assert latutil.rel_l2(xg, xr) < 1e-8
    check = np.linalg.norm(op_g.apply(xg, 0) - b) / np.linalg.norm(b)
    assert check < 2e-10
    for lev in range(levels):
        tr_, tg_ = mg_r.tracker(lev), mg_g.tracker(lev)
        assert abs(tr_["total"] - tg_["total"]) <= 0.1 * tr_["total"] + 8, (lev, tr_, tg_)
    assert mg_g.be.fn("mg_storage_counts")(mg_g.h, 0, (capi.C.c_int * 2)()) == 0
    mg_r.free(); mg_g.free()


def test_n13_kcycle_parity_256(ref, gpu):
    """BASELINE config 2 at its stated size: tests/n13_wilson_kcycle on the reference's own l256t256b60 config
    (wilson_kcycle.cpp:148-194), 2 levels, 4x4 blocks, 8 coarse dof, MR(2,2), coarsest GCR(32) tol 0.2, outer VPGCR(32) to
    1e-10, whole flow native on both back ends (kcycle_new draws the same mt19937 stream).  Mass -0.05: at the usage string's
    -0.075 the CPU reference itself stalls on this config (profiles/r03_oracle_mass_m0075_nonconvergence.log).
    Gates: iteration count +-1, explicit residual, solution to 1e-8."""
    L = 256
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 256)
    res = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        kc = capi.KCycle(be, L, -0.05, g, n_refine=1)
        x, info = kc.solve(b=b, tol=1e-10, restart=32, want_x=True)
        res[name] = (x, info, [kc.tracker(l) for l in range(2)])
        kc.free()
    (xr, ir, tr_), (xg, ig, tg_) = res["ref"], res["gpu"]
    assert ir["success"] and ig["success"], (ir, ig)
    assert abs(ir["iter"] - ig["iter"]) <= 1, (ir, ig)
    assert ig["check_relres"] < 2e-10 and ir["check_relres"] < 2e-10
    assert latutil.rel_l2(xg, xr) < 1e-8
    for lev in range(2):
        assert abs(tr_[lev]["total"] - tg_[lev]["total"]) <= 0.1 * tr_[lev]["total"] + 8, (lev, tr_[lev], tg_[lev])


@pytest.mark.parametrize("level_app,coarsest_app", [(0, 0), (2, 2), (3, 3)])
def test_fused_kcycle_is_bit_identical(gpu, level_app, coarsest_app):
    """The fused K-cycle (zero-start smoothers and coarse solves without A.0, no unread true-residual applies, one-pass
    residual / restrict / prolong-correct, lhs += z3 folded into the smoother's last step) against the reference's
    sweep-for-sweep sequence on the same hierarchy: same iteration counts, same reference operator counts, the SAME BITS in
    one preconditioner application and in the solution -- and a third fewer operator launches on every level.  On top, the
    default mode hands the pre-smoother's residual over (not bit-identical: checked to rounding)."""
    L = 64
    g = latutil.load_gauge(L)
    kc = capi.KCycle(gpu, L, -0.05, g, n_refine=2, level_app=level_app, coarsest_app=coarsest_app)
    mg = capi.Multigrid.__new__(capi.Multigrid)
    mg.be, mg.h = gpu, kc._mg
    b = latutil.gaussian_cv(L * L * 2, 77)
    out = {}
    for fused in (0, 2, 1):      # 2: fused with the explicit pre-smoother residual; 1 (default): the smoother's own residual handed over
        kc.set_fused(fused)
        z = mg.precond(b)
        mg.reset_tracker()
        x, info = kc.solve(b=b, tol=1e-10, restart=32, want_x=True, outer_type=level_app)
        out[fused] = (z, x, info, [kc.tracker(l) for l in range(3)], [kc.executed(l) for l in range(3)])
    # the hand-over is the one shortcut that changes bits (the recurrence residual against rhs - A z): same counts, same
    # solution to rounding, one apply less per K-cycle application on the levels that smooth with plain MR
    (z1, x1, i1, t1, e1), (zh, xh, ih, th, eh) = out[2], out[1]
    assert ih["success"] and ih["iter"] == i1["iter"] and th == t1, (ih, i1, th, t1)
    assert latutil.rel_l2(zh, z1) < 1e-12 and latutil.rel_l2(xh, x1) < 1e-9
    for lev in range(2):
        assert eh[lev] <= e1[lev], (lev, eh, e1)
    if level_app == 0:
        assert eh[0] < e1[0] and eh[1] < e1[1], (eh, e1)
    (z0, x0, i0, t0, e0) = out[0]
    assert i0["success"] and i1["success"]
    assert np.array_equal(z0, z1)
    assert np.array_equal(x0, x1)
    assert i0["iter"] == i1["iter"] and t0 == t1, (i0, i1, t0, t1)
    # unfused: everything the reference counts is launched (+ the post-smooth residual apply it does not count)
    for lev in range(3):
        assert e0[lev] >= t0[lev]["total"] - t0[lev]["nullvec"], (lev, e0, t0)
    # fused: levels 0 and 1 launch 3 + 3 applies per K-cycle application instead of 5 + 5
    for lev in range(2):
        assert e1[lev] < 0.7 * e0[lev], (lev, e0, e1)
    assert e1[2] < e0[2]
    kc.free()


def test_wilson_matrix_free_operator_and_kcycle(ref, gpu):
    """B200 extension on the host classes: Wilson2D::enable_matrix_free_apply checks the stored blocks, then every whole-operator
    apply (apply_M, the residual epilogue of the K-cycle, the solvers' callbacks) reads the links -- same bits, so a K-cycle whose
    fine operator applies matrix-free from the first null-vector solve on reproduces the stored-block hierarchy, iteration counts
    and solution bit for bit (and the oracle's iteration count); an operator whose clover was edited (the n18 mutation) is refused."""
    L = 64
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 5)
    lat = gpu.lattice(L, L, 2)
    op = lat.wilson(-0.03, g)
    want = [op.apply(b, t) for t in (0, 2)]
    assert op.matrix_free(g) == 1
    got = [op.apply(b, t) for t in (0, 2)]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert op.matrix_free(on=2) == 0 and op.matrix_free(on=3) == 1
    op.add_to("clover", 1e-3 * latutil.gaussian_cv(4 * L * L, 2))
    assert op.matrix_free(on=0) == 0 and op.matrix_free(g) == 0       # blocks no longer those of g: refused
    op.free()

    kr = capi.KCycle(ref, L, -0.03, g, n_refine=2, seed=5)
    ir = kr.solve(b, tol=1e-10)
    kr.free()
    res = {}
    for mf in (0, 1):
        was = gpu.fn("kcycle_setup_matrix_free")(mf)
        try:
            kc = capi.KCycle(gpu, L, -0.03, g, n_refine=2, seed=5)
        finally:
            gpu.fn("kcycle_setup_matrix_free")(was)
        assert kc.matrix_free(True) == mf
        x, info = kc.solve(b, tol=1e-10, want_x=True)
        res[mf] = (x, info, [kc.tracker(l) for l in range(3)])
        if mf:
            assert kc.matrix_free(False) == 0
            x2, info2 = kc.solve(b, tol=1e-10, want_x=True)
            assert np.array_equal(x2, x) and info2["iter"] == info["iter"]
        kc.free()
    (x0, i0, t0), (x1, i1, t1) = res[0], res[1]
    assert i0["null_ops"] == i1["null_ops"] and i0["iter"] == i1["iter"] and t0 == t1
    assert np.array_equal(x0, x1)
    assert abs(i1["iter"] - ir["iter"]) <= 1


def test_n19_schur_kcycle_parity(ref, gpu):
    """tests/n19_wilson_kcycle_precond: every level solved as the Schur system of the right-block-Jacobi operator,
    coarse stencils built from the rbjacobi fine stencil, outer tolerance 1e-8."""
    L = 64
    g = latutil.load_gauge(L)
    mass = -0.07
    SCHUR = 3
    mg_r, op_r, lats_r = _build_mg(ref, L, g, mass, [], level_app=SCHUR, coarsest_app=SCHUR)
    nv = _n13_nullvecs(ref, lats_r[0], op_r, 8, seed=19)
    mg_r.free()
    out = {}
    b = latutil.gaussian_cv(L * L * 2, 19)
    for name, be in (("ref", ref), ("gpu", gpu)):
        mg, op, lats = _build_mg(be, L, g, mass, [nv], level_app=SCHUR, coarsest_app=SCHUR, build_from=1, build_extra=2)
        bp = op.prepare(b, SCHUR)
        y, info = mg.solve(bp, outer_type=SCHUR, tol=1e-8, restart=32)
        x = op.reconstruct(y, b, SCHUR)
        out[name] = (x, info, np.linalg.norm(op.apply(x, 0) - b) / np.linalg.norm(b))
        mg.free()
    (xr, ir, cr), (xg, ig, cg) = out["ref"], out["gpu"]
    assert ir["success"] and ig["success"]
    assert abs(ir["iter"] - ig["iter"]) <= 1, (ir, ig)
    assert cg < 5e-8
    assert latutil.rel_l2(xg, xr) < 1e-6


def test_native_kcycle_driver(ref, gpu):
    """kcycle_new / kcycle_solve: the n13 flow with everything generated on the device (same mt19937 draws as the oracle).
    Null vectors come out of 500-iteration BiCGstab-L solves, so only iteration counts (+-2) and residuals are compared."""
    L = 64
    g = latutil.load_gauge(L)
    res = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        kc = capi.KCycle(be, L, -0.075, g, n_refine=2)
        res[name] = kc.solve(tol=1e-10)
        kc.free()
    assert res["gpu"]["success"] and res["ref"]["success"]
    assert abs(res["gpu"]["iter"] - res["ref"]["iter"]) <= 2, res
    assert res["gpu"]["check_relres"] < 2e-10


def test_n16_measurement_loop(ref, gpu):
    """tests/n16_wilson_kcycle_heatbath: heatbath updates -> compact links -> Wilson2D::update_links on the SAME operator ->
    fresh 3-level hierarchy -> two point-source solves -> |prop|^2 per time slice, folded.  The gauge chain is produced
    once (by the GPU heatbath) and fed to both back ends; correlators must agree to the solver tolerance."""
    L, beta, mass = 32, 6.0, -0.01
    lg1 = gpu.lattice(L, L, 1)
    phases = lg1.u1_heatbath(np.zeros(2 * L * L), beta, 300, 1)
    kcs = {name: capi.KCycle(be, L, mass, lg1.u1_polar(phases), n_refine=2, block=4, coarse_dof=8, seed=3) for name, be in (("ref", ref), ("gpu", gpu))}
    for step in range(2):
        phases = lg1.u1_heatbath(phases, beta, 20, 10 + step)
        g = lg1.u1_polar(phases)
        plaq = lg1.u1_observables(g)[0].real
        assert 0.85 < plaq < 0.97
        res = {}
        for name, kc in kcs.items():
            kc.update_links(g)
            res[name] = kc.pion(0, 0, tol=1e-10)
        (pr, ir), (pg, ig) = res["ref"], res["gpu"]
        assert ir["success"] and ig["success"]
        assert abs(ir["iters"] - ig["iters"]) <= 2          # two solves, +-1 each
        assert np.allclose(pg, pr, rtol=1e-6, atol=0)
        assert np.allclose(pg[1:L // 2], pg[:L // 2:-1], rtol=1e-12)     # folded
        assert pg[0] > pg[L // 4] > pg[L // 2] > 0                        # decays away from the source
    for kc in kcs.values():
        kc.free()


def test_staggered_kcycle_parity(ref, gpu):
    """BASELINE config 3 at a size the oracle finishes in seconds: 3-level K-cycle on the staggered operator (nc = 1; four
    BiCGstab-L null vectors doubled by the even / odd projection, staggered.h:176-181; 4x4 blocks twice) on the shipped
    64^2 beta = 6 configuration.  The un-preconditioned staggered operator is a hard case for this cycle (about 160 outer
    iterations at m = 0.1), which makes it a long lock-step comparison of the two back ends."""
    L, mass = 64, 0.1
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L, 5)
    res = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        kc = capi.KCycle(be, L, mass, g, n_refine=2, block=4, coarse_dof=8, seed=3, staggered=True)
        x, info = kc.solve(b, tol=1e-10, want_x=True)
        res[name] = (x, info, [kc.tracker(l)["total"] for l in range(3)])
        kc.free()
    (xr, ir, opr), (xg, ig, opg) = res["ref"], res["gpu"]
    assert ir["success"] and ig["success"]
    assert abs(ir["iter"] - ig["iter"]) <= 2
    assert ig["check_relres"] < 2e-10
    assert all(abs(p - q) <= 0.03 * q + 2 for p, q in zip(opg, opr))
    assert latutil.rel_l2(xg, xr) < 1e-8


def test_staggered_kcycle_full_size(gpu):
    """BASELINE config 3 at full size (1024^2, 3 levels: 1024 -> 256 -> 64) through size-independent properties: the solve
    converges, the explicit residual of the ORIGINAL system meets the tolerance, and the iteration count stays in the
    range the 64^2 run sets (the cycle's convergence rate does not depend on the volume)."""
    L, mass = 1024, 0.1
    g = latutil.synthetic_gauge(L, L, 6.0, 21)
    kc = capi.KCycle(gpu, L, mass, g, n_refine=2, block=4, coarse_dof=8, seed=3, staggered=True, inner_iters=100, coarsest_iters=400)
    out = kc.solve(tol=1e-10, max_iter=600)
    ops = [kc.tracker(l)["total"] for l in range(3)]
    kc.free()
    assert out["success"] and out["check_relres"] < 2e-10
    assert 50 < out["iter"] < 400, (out, ops)


@pytest.mark.parametrize("n_setup", [0, 1])
def test_n22_adaptive_setup_parity(ref, gpu, n_setup):
    """tests/n22_wilson_kcycle_adaptive: test vectors relaxed by Richardson(10, 0.33), then refined by 10 iterations of
    flexible GCR preconditioned by the current K-cycle, update_level + rebuild of the levels below (go_coarser / go_finer),
    outer VPGCR(64).  Same driver text on both back ends; the refinement must pay off identically (132 -> 21 iterations)."""
    L = 64
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 5)
    res = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        kc = capi.KCycle(be, L, -0.05, g, n_refine=2, block=4, coarse_dof=8, seed=3, adaptive_setups=n_setup)
        x, info = kc.solve(b, tol=1e-10, restart=64, want_x=True)
        res[name] = (x, info, [kc.tracker(l)["total"] for l in range(3)])
        kc.free()
    (xr, ir, opr), (xg, ig, opg) = res["ref"], res["gpu"]
    assert ir["success"] and ig["success"]
    assert abs(ir["iter"] - ig["iter"]) <= 1 + ir["iter"] // 50
    assert ir["null_ops"] == ig["null_ops"]
    assert latutil.rel_l2(xg, xr) < 1e-8
    if n_setup == 1:
        assert ig["iter"] < 40


def test_n22_adaptive_full_size(gpu):
    """BASELINE config 4 at full size (4096^2, 3 levels, one adaptive set-up round) through size-independent properties:
    convergence to 1e-10 on the explicit residual of the original system in about as many iterations as at 64^2 / 128^2."""
    L = 4096
    g = latutil.synthetic_gauge(L, L, 6.0, 1337, slab=True)
    kc = capi.KCycle(gpu, L, -0.05, g, n_refine=2, block=4, coarse_dof=8, seed=3, adaptive_setups=1, inner_iters=100, coarsest_iters=400)
    del g
    out = kc.solve(tol=1e-10, restart=16, max_iter=100)
    ops = [kc.tracker(l)["total"] for l in range(3)]
    kc.free()
    import qmg
    qmg.check(qmg.lib().qmg_trim())
    assert out["success"] and out["check_relres"] < 2e-10, (out, ops)
    assert out["iter"] <= 40, (out, ops)


def test_kcycle_with_link_compressed_applies(ref, gpu):
    """B200 extension on the whole hierarchy: every level of the Wilson K-cycle passes the gamma5-hermiticity check, the
    solve takes the same iterations and lands on the same solution (and the oracle's iteration count); an operator whose
    blocks were edited (the n18 mutation) is refused."""
    L = 64
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 5)
    kr = capi.KCycle(ref, L, -0.03, g, n_refine=2, seed=5)
    ir = kr.solve(b, tol=1e-10)
    kr.free()
    kc = capi.KCycle(gpu, L, -0.03, g, n_refine=2, seed=5)
    x0, i0 = kc.solve(b, tol=1e-10, want_x=True)
    assert kc.gamma5_hermitian(True) == 3
    x1, i1 = kc.solve(b, tol=1e-10, want_x=True)
    assert kc.gamma5_hermitian(False) == 0
    x2, i2 = kc.solve(b, tol=1e-10, want_x=True)
    kc.free()
    assert i0["iter"] == i1["iter"] == i2["iter"] and abs(i1["iter"] - ir["iter"]) <= 1
    assert latutil.rel_l2(x1, x0) < 1e-9 and np.array_equal(x2, x0)
    lat = gpu.lattice(L, L, 2)
    op = lat.wilson(-0.03, g)
    assert op.gamma5_hermitian(True) == 1
    op.add_to("hopping", 1e-3 * latutil.gaussian_cv(16 * L * L, 2))
    assert op.gamma5_hermitian(True) == 0
    op.free()


def test_kcycle_shapes_the_reference_uses(ref, gpu):
    """Hierarchy shapes beyond the square 3-level case: four levels down to a SINGLE coarsest site (n16's 64 -> 16 -> 4 -> 1,
    wilson_kcycle_heatbath.cpp:116: the coarsest operator is clover-only and its lattice has volume 1) and a non-square
    64 x 32 lattice (what a y-slab is).  Iteration counts +-1, per-level operator counts within 3 %, same solution."""
    import shard
    L = 64
    g = latutil.load_gauge(L)
    V = L * L
    sl = shard.Slab(L, L, 2, 0)
    cases = [("4 levels to one site", dict(n_refine=3), g, L, latutil.gaussian_cv(V * 2, 5)),
             ("64 x 32", dict(n_refine=2, Y=32), np.concatenate([sl.take(g[:V], 1), sl.take(g[V:], 1)]), 32, latutil.gaussian_cv(L * 32 * 2, 6))]
    for name, kw, gauge, Y, b in cases:
        res = {}
        for bn, be in (("ref", ref), ("gpu", gpu)):
            kc = capi.KCycle(be, L, -0.01, gauge, block=4, coarse_dof=8, seed=3, **kw)
            x, info = kc.solve(b, tol=1e-10, want_x=True)
            res[bn] = (x, info, [kc.tracker(l)["total"] for l in range(kw["n_refine"] + 1)])
            kc.free()
        (xr, ir, opr), (xg, ig, opg) = res["ref"], res["gpu"]
        assert ir["success"] and ig["success"], name
        assert abs(ir["iter"] - ig["iter"]) <= 1, name
        assert all(abs(p - q) <= 0.03 * q + 2 for p, q in zip(opg, opr)), (name, opg, opr)
        assert latutil.rel_l2(xg, xr) < 1e-8, name


def test_unbuilt_variants_and_bad_requests_behave_like_the_reference(ref, gpu, capfd):
    """Error behaviour of the class API (SURVEY.md 8b "Errors"): applying a link set that was never built prints a
    [QMG-WARNING] and leaves the (zeroed) output alone, on both back ends alike; nothing throws, nothing aborts."""
    L = 16
    g = latutil.phases_to_gauge(np.random.default_rng(3).normal(0, 0.3, size=L * L * 2), L, L)
    rhs = latutil.gaussian_cv(L * L * 2, 1)
    outs = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        op = be.lattice(L, L, 2).wilson(0.1, g)
        assert op.built() == 0
        outs[name] = [op.apply(rhs, t) for t in (1, 2, 3, 4, 5, 6, 7, 8)]      # every variant needs a build that did not happen
        op.free()
    capfd.readouterr()          # the warnings themselves sit in the libraries' stdio buffers; what is compared is the effect
    for a_, b_ in zip(outs["ref"], outs["gpu"]):
        assert np.array_equal(a_, b_)


def test_lanczos_eigensolver_and_coarsest_deflation(ref, gpu):
    """the eigensolver behind StatefulMultigridMG::deflate_coarsest (the reference calls ARPACK, absent
    here and in the oracle, so this is checked against numpy): extreme eigenpairs of M^dag M on a small Wilson lattice
    against a dense diagonalisation, residuals |A v - lambda v|, refusal of a non-Hermitian operator; then a 3-level K-cycle
    whose coarsest solve is CG on the normal equations, with and without a deflation space: same answer, fewer coarsest
    iterations."""
    L = 8
    g = latutil.phases_to_gauge(np.random.default_rng(5).normal(0, 0.4, size=L * L * 2), L, L)
    lat = gpu.lattice(L, L, 2)
    op = lat.wilson(0.05, g)
    op.build(dagger=True)
    n = lat.size_cv
    cols = np.stack([op.apply(np.eye(n, dtype=np.complex128)[j], 5) for j in range(n)], axis=1)      # M^dag M, column by column
    assert np.allclose(cols, cols.conj().T, atol=1e-12)
    want = np.linalg.eigvalsh(cols)
    ok, ev, vec = op.eigs(5, 6, tol=1e-9, want_vectors=True)
    assert ok and np.allclose(ev, want[:6], rtol=1e-7)
    for lam, v in zip(ev, vec):
        assert np.linalg.norm(cols @ v - lam * v) < 1e-6 * np.linalg.norm(v) * want[-1]
    ok, evh, _ = op.eigs(5, 4, high=True, tol=1e-9)
    assert ok and np.allclose(np.sort(evh), want[-4:], rtol=1e-7)
    ok, _, _ = op.eigs(0, 4)                 # the Wilson operator itself is not Hermitian: refused
    assert not ok
    op.free()

    L = 64
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 5)
    kc = capi.KCycle(gpu, L, -0.06, g, n_refine=2, seed=5, coarsest_app=5, coarsest_tol=0.05)     # coarsest: CG on M^dag M
    x0, i0 = kc.solve(b, tol=1e-10, want_x=True)
    plain = kc.tracker(2)["total"]
    ev = kc.deflate_coarsest(16)
    assert ev.size == 16 and np.all(ev > 0) and np.all(np.diff(ev) >= -1e-12)
    x1, i1 = kc.solve(b, tol=1e-10, want_x=True)
    deflated = kc.tracker(2)["total"]
    kc.free()
    assert i0["success"] and i1["success"] and abs(i0["iter"] - i1["iter"]) <= 2
    assert latutil.rel_l2(x1, x0) < 1e-8
    assert deflated < 0.8 * plain, (plain, deflated)


def test_n11_solver_survey_extras(ref, gpu):
    """tests/n11_wilson_test also calls minv_vector_bicgstab and minv_vector_tfqmr (wilson_test.cpp:185, :241).  BiCGstab is
    BiCGstab(1) on both back ends (iteration parity); TFQMR exists only as a declaration in the oracle's shim, so the device
    version is held to its own contract: it converges and the reported residual is the true one."""
    L = 32
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 4)
    res = {}
    for name, be in (("ref", ref), ("gpu", gpu)):
        op = be.lattice(L, L, 2).wilson(0.1, g)
        res[name] = op.solve(7, b, max_iter=2000, tol=1e-9)
        if name == "gpu":
            x, info = op.solve(8, b, max_iter=2000, tol=1e-9)
            assert info["success"] and 0 < info["iter"] < 400
            assert latutil.rel_l2(op.apply(x, 0), b) < 2e-9
            assert abs(np.sqrt(info["resSq"]) / np.linalg.norm(b) - latutil.rel_l2(op.apply(x, 0), b)) < 1e-12
        op.free()
    (xr, ir), (xg, ig) = res["ref"], res["gpu"]
    assert ir["success"] and ig["success"] and abs(ir["iter"] - ig["iter"]) <= 2
    assert latutil.rel_l2(xg, xr) < 1e-7
